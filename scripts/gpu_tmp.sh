#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29549"
( timeout 80 $TR bench.py --gpus 2 --ncell 96 --npz 16 --steps 3 --warmup 3 --no-cpu ) > gpurun_out/b2.json 2> gpurun_out/b2.err; tail -c 500 gpurun_out/b2.json; tail -2 gpurun_out/b2.err
