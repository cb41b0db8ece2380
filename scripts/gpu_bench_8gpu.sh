#!/bin/bash
# 8-GPU pass: tracer-group (weak, default flags as the driver runs it) and face x tracer (strong) bench at N=8, weak at N=4
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/box8.txt; free -g >> gpurun_out/box8.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
( time timeout 900 $TR bench.py --gpus 8 --steps 3 --warmup 3 ) > gpurun_out/bench_n8_tracer.json 2> gpurun_out/bench_n8_tracer.err; tail -c 900 gpurun_out/bench_n8_tracer.json; echo; tail -4 gpurun_out/bench_n8_tracer.err
( timeout 600 $TR bench.py --gpus 8 --steps 3 --warmup 3 --shard face ) > gpurun_out/bench_n8_face.json 2> gpurun_out/bench_n8_face.err; tail -c 400 gpurun_out/bench_n8_face.json; echo
TR4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29543"
( timeout 600 $TR4 bench.py --gpus 4 --steps 3 --warmup 3 --no-e2e ) > gpurun_out/bench_n4_tracer.json 2> gpurun_out/bench_n4_tracer.err; tail -c 300 gpurun_out/bench_n4_tracer.json; echo
