#!/bin/bash
# N GPUs (NG env): ONE C768 problem, 24 sub-domains (--shard sub) vs face x tracer groups (--shard face)
mkdir -p gpurun_out
NG=${NG:-8}
for sh in ${SHARDS:-sub face}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $NG --steps 5 --warmup 3 --shard $sh --e2e-steps 1 > gpurun_out/bench_${NG}gpu_$sh.json 2> gpurun_out/bench_${NG}gpu_$sh.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${NG}gpu_$sh.json").read().strip().splitlines()[-1])
    print("$sh", d["n_gpus"], d["ms_per_step"], d["e2e"]["ms_per_step"] if d.get("e2e") else None, d["roofline"]["rank0_ms_per_step"], d["config"]["halo_bytes_sent_per_rank_and_substep"], d["config"].get("halo_update_ms"), d["config"].get("cmax_reduction_ms"))
except Exception as e:
    print("ERR $sh", e); print(open("gpurun_out/bench_${NG}gpu_$sh.err").read()[-2500:])
P
done
