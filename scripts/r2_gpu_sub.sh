#!/bin/bash
# sub-tile contexts: parity with whole tiles; regression of the whole-tile suites that share the changed kernels
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_subdomain.py -m gpu -q -x ) > gpurun_out/pytest_sub.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_sub.log
grep -E "FAILED|passed|failed|rc=|Error|assert " gpurun_out/pytest_sub.log | cut -c1-300 | tail -12
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_c_client.py -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], {k:v["avg_ms"] if isinstance(v,dict) else v for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("ERR", e); print(open("gpurun_out/bench_quick.err").read()[-2000:])
P
