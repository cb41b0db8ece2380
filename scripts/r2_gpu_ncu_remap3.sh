#!/bin/bash
# full ncu capture of k_remap3 (C384 L127 x9 fp64) for the per-line / per-instruction picture
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_remap3 -s 2 -c 1 -o gpurun_out/prof_remap3_${TAG:-r02b} -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_remap3.log 2>&1
tail -2 gpurun_out/ncu_remap3.log
