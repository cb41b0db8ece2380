"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py does that when the box has them): every rank computes its
(faces, tracers) block of ONE global tracer_2d + remap problem through ShardedTracerStep (NCCL strip exchange + cmax
all-reduce) and compares it with the same block of a single-context run of the whole mosaic on its own GPU.  Same kernels,
same inputs -> the blocks must be bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fv3atm_b200 import partition, synthetic as sy  # noqa: E402
from fv3atm_b200.tracer import TracerContext  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    n, npz, nq, hord = 24, 8, 9, 8
    case = sy.make_case(n, npz, nq, dtype=np.float64, courant=1.8)
    kord = np.full(nq, 9, dtype=np.int32)
    # reference: the whole mosaic in one context on this GPU
    full = TracerContext(n + 1, npz, nq, case.metrics(), dtype=np.float64, device=lr)
    for f in ("q", "dp1", "mfx", "mfy", "cx", "cy", "pe"):
        full.upload(f, getattr(case, f), nq)
    full.set_vertical(case.ak, case.bk, case.ptop)
    nsplt_ref = full.tracer_2d_resident(nq, hord)
    full.remap_tracers_resident(nq, kord, fill=True)
    qref = np.empty_like(case.q)
    full.download("q", qref, nq)
    full.close()
    # sharded
    F, G = partition.choose_layout(world, nq, prefer="face")
    layout = partition.Layout(world, F, G, nq)
    tiles = layout.tiles(rank)
    q0, cnt = layout.tracers(rank)
    tl = [t - 1 for t in tiles]
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = TracerContext(n + 1, npz, cnt, case.metrics(), dtype=np.float64, tiles=tiles, device=lr, stream=stream.cuda_stream)
        ctx.upload("q", np.ascontiguousarray(case.q[tl][:, q0:q0 + cnt]), cnt)
        for f in ("dp1", "mfx", "mfy", "cx", "cy", "pe"):
            ctx.upload(f, np.ascontiguousarray(getattr(case, f)[tl]), cnt)
        ctx.set_vertical(case.ak, case.bk, case.ptop)
        step = partition.ShardedTracerStep(ctx, layout, rank, lr)
        nsplt = step.tracer_2d(hord)
        step.remap(kord[q0:q0 + cnt], fill=True)
        q = np.empty((len(tl), cnt) + case.q.shape[2:], dtype=np.float64)
        ctx.download("q", q, cnt)
        ctx.close()
    sl = slice(3, -3)
    ok = nsplt == nsplt_ref and nsplt >= 2 and np.array_equal(q[..., sl, sl], qref[tl][:, q0:q0 + cnt][..., sl, sl])
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{lr}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MGPU_CHECK world={world} layout={F}x{G} nsplt={nsplt} ok={bool(flag.item())}")
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
