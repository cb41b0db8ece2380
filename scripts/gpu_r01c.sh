#!/bin/bash
# round-1 GPU pass C: fast advection kernel (k_advect3): parity, C768 bench, full ncu capture at C384
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for nt in 0 96 128 224; do
FV3T_ADV_NT=$nt timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c768_nt$nt.json 2> gpurun_out/bench_c768_nt$nt.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c768_nt$nt.json').read().strip().splitlines()[-1]); print($nt, d['ms_per_step'], d['roofline']['kernels'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect3 -s 3 -c 1 -o gpurun_out/prof_advect3_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv3.log 2>&1
ls -la gpurun_out | head -30
