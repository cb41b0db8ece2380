#!/bin/bash
# the driver's own command line at N GPUs (NG env), default flags; optional second run with SHARD2
mkdir -p gpurun_out
NG=${NG:-8}
run() {
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $NG --steps 5 --warmup 3 $2 > gpurun_out/bench_${NG}gpu_$1.json 2> gpurun_out/bench_${NG}gpu_$1.err
python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${NG}gpu_$1.json").read().strip().splitlines()[-1])
    c=d["config"]
    print("$1", d["n_gpus"], round(d["ms_per_step"],2), round(d["e2e"]["ms_per_step"],1) if d.get("e2e") else None, d["roofline"]["rank0_ms_per_step"], c.get("halo_update_ms"), c.get("cmax_reduction_ms"), c.get("kernel_ms_per_step_by_rank"), c["parallelism"][:60]); print(c.get("rank0_phases"))
except Exception as e:
    print("ERR $1", e); print(open("gpurun_out/bench_${NG}gpu_$1.err").read()[-2500:])
P
}
run default ""
if [ -n "$SHARD2" ]; then run $SHARD2 "--shard $SHARD2"; fi
