#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/b.json 2> gpurun_out/b.err; python -c "import json,sys; d=json.loads(open('gpurun_out/b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernels'])"; tail -2 gpurun_out/b.err
timeout 600 python -m pytest tests -m gpu -q -k "tracer_2d or config1 or extremes or tracer_step" 2>&1 | tail -2
