#!/bin/bash
# round-1 GPU pass R (8 GPUs): tracer-group (weak) and face x tracer (strong) bench at N=8, sharded-path check on 2 of them
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/box8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
( timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e ) > gpurun_out/bench_n8_tracer.json 2> gpurun_out/bench_n8_tracer.err; tail -c 700 gpurun_out/bench_n8_tracer.json; echo
( timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 --shard face ) > gpurun_out/bench_n8_face.json 2> gpurun_out/bench_n8_face.err; tail -c 1300 gpurun_out/bench_n8_face.json; tail -3 gpurun_out/bench_n8_face.err
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543"
( timeout 600 $TR4 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e ) > gpurun_out/bench_n4_tracer.json 2> gpurun_out/bench_n4_tracer.err; tail -c 300 gpurun_out/bench_n4_tracer.json; echo
