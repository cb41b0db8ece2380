#!/bin/bash
# round-1 GPU pass B: parity tests, full default bench (C768, e2e + cpu baseline), reference arm, launch list,
# DRAM traffic of the C768 kernels, full ncu captures at C384
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/box.txt 2>&1
nproc >> gpurun_out/box.txt; lscpu | grep "Model name" >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 1500 gpurun_out/bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_c384.csv \
  python bench.py --n 384 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_advect2|k_remap2" -s 6 -c 2 --csv \
  --log-file gpurun_out/traffic_c768.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_traffic.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect2 -s 3 -c 1 -o gpurun_out/prof_advect2_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_remap2 -s 3 -c 1 -o gpurun_out/prof_remap2_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_rm.log 2>&1
ls -la gpurun_out
