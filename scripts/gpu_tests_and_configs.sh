#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "FAILED|passed|failed|^E  " gpurun_out/pytest_gpu.log | cut -c1-300 | tail -12
( timeout 600 python bench.py --n 384 --dtype float32 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench_c384_f32.json 2> gpurun_out/bench_c384_f32.err; tail -c 900 gpurun_out/bench_c384_f32.json; echo
( timeout 600 python bench.py --n 384 --nq 30 --steps 3 --warmup 3 --no-cpu --no-e2e ) > gpurun_out/bench_c384_nq30.json 2> gpurun_out/bench_c384_nq30.err; tail -c 900 gpurun_out/bench_c384_nq30.json; echo
( timeout 600 python bench.py --n 384 --dtype float32 --hord 13 --steps 3 --warmup 3 --no-cpu --no-e2e ) > gpurun_out/bench_c384_f32_h13.json 2> gpurun_out/bench_c384_f32_h13.err; tail -c 600 gpurun_out/bench_c384_f32_h13.json; echo
( timeout 600 python bench.py --n 384 --dtype float32 --courant 2.6 --steps 3 --warmup 3 --no-cpu --no-e2e ) > gpurun_out/bench_c384_f32_nsplt.json 2> gpurun_out/bench_c384_f32_nsplt.err; tail -c 600 gpurun_out/bench_c384_f32_nsplt.json
