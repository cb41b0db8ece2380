#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['gpu_launches'], d['e2e'], d['roofline']['kernels'])"; tail -3 gpurun_out/b.err
