#!/bin/bash
# round-2: full ncu capture of the advection kernel (C768, 16 levels: 1344 CTAs = 9 waves)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect5 -s 2 -c 1 -o gpurun_out/prof_advect5_${TAG:-cur} -f \
  python bench.py --npz 16 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv5.log 2>&1
tail -2 gpurun_out/ncu_adv5.log
