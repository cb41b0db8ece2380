#!/bin/bash
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q -k "tracer_2d" ) > gpurun_out/pytest_gpu.log 2>&1; grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | tail -5
for u in 1 2 3; do
FV3T_ADV_UNROLL=$u timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c768_u$u.json 2> gpurun_out/bench_c768_u$u.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c768_u$u.json').read().strip().splitlines()[-1]); print('unroll', $u, d['ms_per_step'], d['roofline']['kernels'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect3 -s 3 -c 1 -o gpurun_out/prof_advect3_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv3.log 2>&1
