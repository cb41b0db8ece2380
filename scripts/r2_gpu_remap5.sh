#!/bin/bash
# two-walk remap (k_remap5): parity (remap / step tests, the five BASELINE configs) and the quick C768 bench, both kernel pairs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "remap or step or map" ) > gpurun_out/pytest_remap.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_remap.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_remap.log | cut -c1-300 | tail -8
( timeout 900 python -m pytest tests/test_gpu_configs.py -m gpu -q ) > gpurun_out/pytest_configs.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_configs.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_configs.log | cut -c1-300 | tail -8
for v in 1 0; do
FV3T_REMAP5=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/bench_r5_$v.json 2> gpurun_out/bench_r5_$v.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_r5_$v.json").read().strip().splitlines()[-1])
    print("REMAP5=$v", d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_r5_$v.err").read()[-2000:])
P
done
FV3T_REMAP_MINB=5 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/bench_r5_mb5.json 2> gpurun_out/bench_r5_mb5.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_r5_mb5.json").read().strip().splitlines()[-1])
print("MINB=5", d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
P
