#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernels'])"; tail -2 gpurun_out/b.err
timeout 600 python -m pytest tests -m gpu -q -k "remap or resident or extremes or config1 or conservation" 2>&1 | tail -2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_remap3 -s 1 -c 1 -o gpurun_out/prof_remap3_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rm3.log 2>&1
