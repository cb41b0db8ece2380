#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python bench.py --n 384 --dtype float32 --steps 5 --warmup 3 --no-cpu ) > gpurun_out/bench_c384_f32.json 2> gpurun_out/bench_c384_f32.err; tail -c 300 gpurun_out/bench_c384_f32.json; echo
( timeout 600 python bench.py --n 384 --nq 30 --steps 3 --warmup 3 --no-cpu --no-e2e ) > gpurun_out/bench_c384_nq30.json 2> gpurun_out/bench_c384_nq30.err; tail -c 300 gpurun_out/bench_c384_nq30.json; echo
( timeout 600 python bench.py --n 384 --dtype float32 --hord 13 --steps 3 --warmup 3 --no-cpu --no-e2e ) > gpurun_out/bench_c384_f32_h13.json 2> gpurun_out/bench_c384_f32_h13.err; tail -c 200 gpurun_out/bench_c384_f32_h13.json; echo
