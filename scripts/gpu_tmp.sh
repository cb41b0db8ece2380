#!/bin/bash
mkdir -p gpurun_out
run() { env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['roofline']['kernels']['k_advect3']['avg_ms'])"; tail -2 gpurun_out/b.err; }
run FV3T_ADV_MINB=8
run FV3T_ADV_MINB=10
run FV3T_ADV_MINB=12
run FV3T_ADV_MINB=10 FV3T_ADV_NT=32
