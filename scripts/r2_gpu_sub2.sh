#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_subdomain.py tests/test_gpu_parity.py tests/test_c_client.py -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|rc=" gpurun_out/pytest_gpu.log | cut -c1-300 | tail -8
for h in 10 8; do timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity --hord $h 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hord', $h, round(d['ms_per_step'],2), {k:(round(v['avg_ms'],2) if isinstance(v,dict) else round(v,2)) for k,v in d['roofline']['kernels'].items()})"; done
