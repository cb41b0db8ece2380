#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['roofline']['kernels'])"; tail -2 gpurun_out/b.err; }
run FV3T_ADV_RING=1
run FV3T_ADV_RING=1 FV3T_ADV_NT=128
run FV3T_ADV_RING=0
timeout 600 python -m pytest tests -m gpu -q -k "tracer_2d or config1 or extremes or resident or conservation" 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect4 -s 3 -c 1 -o gpurun_out/prof_advect4_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv4.log 2>&1
