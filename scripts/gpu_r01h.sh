#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200 | tail -30
FV3T_ADV_NT=64 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c768_h.json 2> gpurun_out/bench_c768_h.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c768_h.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernels'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_remap3 -s 1 -c 1 -o gpurun_out/prof_remap3_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rm3.log 2>&1
