#!/bin/bash
# round-2 GPU pass: first run of k_advect5 (parity tests, C768 bench against k_advect4, tracer-group sweep)
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|^E  |rc=" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -15
for tg in 9 5 3; do
  FV3T_ADV_TG=$tg timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_a5_tg$tg.json 2> gpurun_out/bench_a5_tg$tg.err
  echo "tg=$tg rc=$?"; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_a5_tg$tg.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
except Exception as e: print("ERR",e); print(open("gpurun_out/bench_a5_tg$tg.err").read()[-1500:])
P
done
FV3T_ADV5=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_a4.json 2> gpurun_out/bench_a4.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_a4.json").read().strip().splitlines()[-1])
print("adv4", d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
P
