#!/bin/bash
# final 1-GPU evidence: default bench line, launch list with DRAM traffic of the C768 step
mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_advect5|k_remap3|k_prep5|k_remap_coef3|k_cmax|k_halo_fill|k_pad" -s 16 -c 8 --csv \
  --log-file gpurun_out/launches_c768.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_launch.log 2>&1
grep -c "k_" gpurun_out/launches_c768.csv
