"""Run under torchrun on >= 2 GPUs (tests/test_gpu_multi.py): ONE global tracer_2d + remap problem decomposed into 2 x 2
sub-domains per tile (24 sub-domains), every rank holding its share in sub-tile contexts; halos --
diagonal blocks included -- travel as packed gather lists over NCCL, cmax by all-reduce(max).  Every rank compares its
sub-domains with a single-context whole-tile run on its own GPU: bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fv3atm_b200 import synthetic as sy  # noqa: E402
from fv3atm_b200.subdomain import SubMosaic, SubMosaicStep  # noqa: E402
from fv3atm_b200.tracer import TracerContext  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    n, npz, nq, hord = 24, 8, 9, 10
    case = sy.make_case(n, npz, nq, dtype=np.float64, courant=1.8)
    kord = np.full(nq, 9, dtype=np.int32)
    full = TracerContext(n + 1, npz, nq, case.metrics(), dtype=np.float64, device=lr)
    for f in ("q", "dp1", "mfx", "mfy", "cx", "cy", "pe"):
        full.upload(f, getattr(case, f), nq)
    full.set_vertical(case.ak, case.bk, case.ptop)
    nsplt_ref = full.tracer_2d_resident(nq, hord)
    full.remap_tracers_resident(nq, kord, fill=True)
    qref = np.empty_like(case.q)
    full.download("q", qref, nq)
    full.close()
    mo = SubMosaic(n, 2)
    run = SubMosaicStep(mo, rank, world, lr, npz, nq, np.float64, case.metrics())
    run.upload_case(case)
    nsplt = run.tracer_2d(hord)
    run.remap(kord)
    got = run.download("q", np.full_like(case.q, np.nan))
    run.close()
    sl = slice(3, -3)
    mine = ~np.isnan(got[..., sl, sl])
    ok = nsplt == nsplt_ref and nsplt >= 2 and mine.sum() * world == mine.size and np.array_equal(got[..., sl, sl][mine], qref[..., sl, sl][mine])
    flag = torch.tensor([1 if ok else 0], device=f"cuda:{lr}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"MGPU_SUB_CHECK world={world} subdomains={len(mo)} nsplt={nsplt} ok={bool(flag.item())}")
    dist.destroy_process_group()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
