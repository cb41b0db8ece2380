#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed" gpurun_out/pytest_gpu.log | cut -c1-200 | tail -30
for mb in 1 2; do for nt in 128 160; do
FV3T_ADV_MINB=$mb FV3T_ADV_NT=$nt timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c768_mb${mb}_nt$nt.json 2> gpurun_out/bench_c768_mb${mb}_nt$nt.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c768_mb${mb}_nt$nt.json').read().strip().splitlines()[-1]); print($mb, $nt, d['ms_per_step'], d['roofline']['kernels']['k_advect']['avg_ms'])"
done; done
for sm in 110000 74000 56000 44000; do
FV3T_REMAP_SMEM=$sm timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c768_rs$sm.json 2> gpurun_out/bench_c768_rs$sm.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c768_rs$sm.json').read().strip().splitlines()[-1]); print('remap smem', $sm, d['ms_per_step'], d['roofline']['kernels']['k_remap']['avg_ms'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_advect3 -s 3 -c 1 -o gpurun_out/prof_advect3_c384 -f \
  python bench.py --n 384 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_adv3.log 2>&1
