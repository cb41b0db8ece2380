#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['roofline']['kernels']['k_advect3']['avg_ms'])"; }
run FV3T_ADV_UNROLL=1
run FV3T_ADV_UNROLL=2
run FV3T_ADV_MINB=3
run FV3T_ADV_MINB=3 FV3T_ADV_NT=128
run FV3T_ADV_UNROLL=2 FV3T_ADV_NT=128
