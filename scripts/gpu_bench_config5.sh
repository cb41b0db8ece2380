#!/bin/bash
# BASELINE config 5 in miniature: C384 L127, 30 tracers, fp64, ONE problem sharded by tracer group (15+15) over 2 GPUs, winds / fluxes / delp replicated
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29547"
( timeout 100 $TR bench.py --gpus 2 --ncell 384 --nq 30 --steps 3 --warmup 3 --shard group ) > gpurun_out/bench_n2_c384_nq30_group.json 2> gpurun_out/bench_n2_c384_nq30_group.err; tail -c 900 gpurun_out/bench_n2_c384_nq30_group.json; tail -3 gpurun_out/bench_n2_c384_nq30_group.err
