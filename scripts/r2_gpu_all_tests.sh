#!/bin/bash
# round-2: the whole GPU suite (parity matrix + BASELINE-config sizes) and a device-resident bench line
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q -s --durations=8 ) > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_all.log
grep -E "FAILED|passed|failed|rc=|^E  |config [0-9]|real" gpurun_out/pytest_gpu_all.log | cut -c1-300 | tail -30
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity $BENCH_ARGS > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_quick.err").read()[-2000:])
P
