#!/bin/bash
# round-1 GPU pass J (2 GPUs): sharded-path correctness, tracer-group (weak) and face-sharded (strong) bench at N=2
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/box2.txt
( timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/pytest_mgpu.log 2>&1; tail -3 gpurun_out/pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
( timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e ) > gpurun_out/bench_n2_tracer.json 2> gpurun_out/bench_n2_tracer.err; tail -c 1200 gpurun_out/bench_n2_tracer.json
( timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 --shard face ) > gpurun_out/bench_n2_face.json 2> gpurun_out/bench_n2_face.err; tail -c 1500 gpurun_out/bench_n2_face.json; tail -5 gpurun_out/bench_n2_face.err
