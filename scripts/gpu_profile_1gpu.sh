#!/bin/bash
# round-1 GPU pass I: default bench (C768, e2e + cpu baseline), reference arm, launch list and DRAM traffic of the C768 kernels,
# full ncu captures of the fast kernels at C384
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/box.txt 2>&1
nproc >> gpurun_out/box.txt; lscpu | grep "Model name" >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
( timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' ) > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 2500 gpurun_out/bench_default.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 600 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_advect4|k_remap3|k_prep3|k_remap_coef3|k_cmax|k_halo_fill" -s 18 -c 12 --csv \
  --log-file gpurun_out/launches_c768.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect4 -s 3 -c 1 -o gpurun_out/prof_advect4_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_remap3 -s 1 -c 1 -o gpurun_out/prof_remap3_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rm3.log 2>&1
ls -la gpurun_out | head -40
