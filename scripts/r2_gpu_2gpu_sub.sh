#!/bin/bash
# 2 GPUs: multi-GPU parity (face strips and sub-domain mosaic), then ONE C768 problem: 24 sub-domains vs 2 face groups
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/pytest_multi.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_multi.log
grep -E "FAILED|passed|failed|rc=|ok=" gpurun_out/pytest_multi.log | cut -c1-300 | tail -8
for sh in sub face; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --shard $sh --e2e-steps 1 > gpurun_out/bench_2gpu_$sh.json 2> gpurun_out/bench_2gpu_$sh.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_2gpu_$sh.json").read().strip().splitlines()[-1])
    print("$sh", d["ms_per_step"], d["e2e"]["ms_per_step"] if d.get("e2e") else None, d["roofline"]["rank0_ms_per_step"], d["config"]["halo_bytes_sent_per_rank_and_substep"])
except Exception as e:
    print("ERR $sh", e); print(open("gpurun_out/bench_2gpu_$sh.err").read()[-2500:])
P
done
