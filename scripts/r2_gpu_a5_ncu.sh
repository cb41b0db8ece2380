#!/bin/bash
# round-2: full ncu capture of k_advect5 (C768, 16 levels: 1344 CTAs = 9 waves) + parity tests with the micro-optimised element functions
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_a5.json 2> gpurun_out/bench_a5.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_a5.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
P
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect5 -s 2 -c 1 -o gpurun_out/prof_advect5_c768l16 -f \
  python bench.py --npz 16 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv5.log 2>&1
tail -3 gpurun_out/ncu_adv5.log
ls -la gpurun_out/*.ncu-rep
