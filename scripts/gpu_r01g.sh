#!/bin/bash
mkdir -p gpurun_out
for mb in 2 1; do for nt in 32 64 96; do
FV3T_ADV_MINB=$mb FV3T_ADV_NT=$nt timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_c768_mb${mb}_nt$nt.json 2> gpurun_out/bench_c768_mb${mb}_nt$nt.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_c768_mb${mb}_nt$nt.json').read().strip().splitlines()[-1]); print($mb, $nt, d['ms_per_step'], d['roofline']['kernels']['k_advect']['avg_ms'])"
done; done
