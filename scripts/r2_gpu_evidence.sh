#!/bin/bash
# round-2 evidence pass: default bench (C768, e2e + cpu baseline + parity), reference arm, launch list with DRAM traffic of the C768 step,
# fp32 / 30-tracer / hord-10 bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/box.txt 2>&1
nproc >> gpurun_out/box.txt; lscpu | grep "Model name" >> gpurun_out/box.txt; free -g >> gpurun_out/box.txt
( time timeout 1500 python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 4000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 900 gpurun_out/bench_reference.json
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"k_advect5|k_remap3|k_prep5|k_remap_coef3|k_cmax|k_halo_fill|k_pad" -s 16 -c 8 --csv \
  --log-file gpurun_out/launches_c768.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-parity > gpurun_out/ncu_launch.log 2>&1
grep -c "k_" gpurun_out/launches_c768.csv
for extra in "--dtype float32 --ncell 384" "--nq 30 --ncell 384" "--hord 10" "--dtype float32"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-parity $extra 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$extra', round(d['ms_per_step'],2), d['value'], {k:(round(v['avg_ms'],2) if isinstance(v,dict) else round(v,2)) for k,v in d['roofline']['kernels'].items()})" | tee -a gpurun_out/bench_extra.txt
done
