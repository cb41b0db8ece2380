"""Multi-GPU partitioning of the tracer-transport path: one process per GPU, torch.distributed for the plumbing.

Two axes, freely combined as world = F x G:
  * G tracer groups -- tracers never interact in tracer_2d (the iq loop body, fv_tracer2d.F90:527-544) nor in
    mapn_tracer / fillz, so a rank that owns all faces of its tracers needs NO data-path collective; winds, mass
    fluxes and delp are replicated.
  * F face groups   -- the six cubed-sphere faces are split over F in {1, 2, 3, 6} ranks.  Each sub-step then has one
    real exchange: the 3-cell edge strips of q across the face-group boundaries (the reference's
    start/complete_group_halo_update through mpp_domains, fv_tracer2d.F90:499,561; contact table
    tools/fv_mp_mod.F90:581-629), plus one all-reduce(MAX) of cmax(1:npz) per call (fv_tracer2d.F90:433).  The strips
    are packed on the device already rotated into the receiver's index order (fv3t_*_halo_pack) and travel as ONE
    grouped NCCL send/recv batch per sub-step (torch.distributed.batch_isend_irecv -> ncclGroupStart/End) over NVLink.
The vertical remap is column-local: no communication in either axis.

Everything here is host logic + plumbing; the arithmetic lives in libfv3tracer.so.  The same exchange schedule runs
on CPU tensors with the gloo backend (tests/test_partition_gloo.py) against a numpy strip packer built from the same
contact table."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import cubed_sphere as cs

FACE_GROUPS = {
    1: [(1, 2, 3, 4, 5, 6)],
    # tiles 1,2,3 share the corner they all meet at: 3 of the 12 contacts stay inside each half
    2: [(1, 2, 3), (4, 5, 6)],
    3: [(1, 2), (3, 4), (5, 6)],
    6: [(1,), (2,), (3,), (4,), (5,), (6,)],
}


def tracer_groups(nq: int, ngroups: int):
    """Contiguous tracer ranges [(first, count)], sizes differing by at most one (30 -> 4+4+4+4+4+4+3+3)."""
    if ngroups < 1 or ngroups > nq:
        raise ValueError(f"cannot split {nq} tracers into {ngroups} groups")
    base, extra = divmod(nq, ngroups)
    out, first = [], 0
    for g in range(ngroups):
        cnt = base + (1 if g < extra else 0)
        out.append((first, cnt))
        first += cnt
    return out


def choose_layout(world: int, nq: int, prefer: str = "face"):
    """(F, G) with F * G == world, F in {1,2,3,6}, G <= nq.  prefer='face' maximises F, 'tracer' maximises G."""
    cands = [(f, world // f) for f in (1, 2, 3, 6) if world % f == 0 and world // f <= nq]
    if not cands:
        raise ValueError(f"no face x tracer layout for world={world}, nq={nq}")
    return max(cands, key=lambda fg: fg[0]) if prefer == "face" else min(cands, key=lambda fg: fg[0])


@dataclass(frozen=True)
class Layout:
    """rank = g * F + f : face group f (fastest) of tracer group g."""
    world: int
    F: int
    G: int
    nq: int

    def coords(self, rank: int):
        return rank % self.F, rank // self.F  # (f, g)

    def rank_of(self, f: int, g: int) -> int:
        return g * self.F + f

    def tiles(self, rank: int):
        return FACE_GROUPS[self.F][self.coords(rank)[0]]

    def tracers(self, rank: int):
        return tracer_groups(self.nq, self.G)[self.coords(rank)[1]]

    def owner_of_tile(self, tile: int, g: int) -> int:
        for f, ts in enumerate(FACE_GROUPS[self.F]):
            if tile in ts:
                return self.rank_of(f, g)
        raise ValueError(tile)

    def face_peers(self, rank: int):
        """ranks holding the other face groups of this rank's tracer group (the cmax / halo communicator)"""
        _, g = self.coords(rank)
        return [self.rank_of(f, g) for f in range(self.F)]


@dataclass(frozen=True)
class StripMsg:
    """One directed edge strip: produced by halo_pack(local_tile, edge) on `src_rank`, consumed by
    halo_unpack(dst_local_tile, dst_edge) on `dst_rank`."""
    src_rank: int
    src_local_tile: int
    src_edge: int
    dst_rank: int
    dst_local_tile: int
    dst_edge: int
    tag: int


def strip_schedule(layout: Layout, rank: int):
    """(sends, recvs) of `rank` for one sub-step, both sorted by a global tag (= 4 * receiver tile + receiver edge) so
    that every pair of ranks posts matching operations in the same order (required by NCCL and by gloo)."""
    maps = cs.edge_maps(8)  # topology does not depend on n
    f, g = layout.coords(rank)
    mine = layout.tiles(rank)
    sends, recvs = [], []
    for lt, tile in enumerate(mine):
        for e in range(4):
            m = maps[tile - 1][e]
            nbr = m.nbr_tile + 1
            if nbr in mine:
                continue  # both sides resident: fv3t_*_halo_local
            peer = layout.owner_of_tile(nbr, g)
            peer_lt = layout.tiles(peer).index(nbr)
            # my cells feed the halo of (nbr, nbr_edge); my halo beyond e is fed by nbr's cells
            sends.append(StripMsg(rank, lt, e, peer, peer_lt, m.nbr_edge, 4 * (nbr - 1) + m.nbr_edge))
            recvs.append(StripMsg(peer, peer_lt, m.nbr_edge, rank, lt, e, 4 * (tile - 1) + e))
    sends.sort(key=lambda s: s.tag)
    recvs.sort(key=lambda s: s.tag)
    return sends, recvs


class StripExchanger:
    """Runs strip_schedule with torch.distributed point-to-point ops.  `pack(local_tile, edge, buf)` fills a flat
    tensor, `unpack(local_tile, edge, buf)` scatters one; `alloc(nelem)` returns a flat tensor on the right device."""

    def __init__(self, layout: Layout, rank: int, alloc, strip_elems: int, group=None):
        import torch.distributed as dist
        self.dist = dist
        self.layout, self.rank, self.group = layout, rank, group
        self.sends, self.recvs = strip_schedule(layout, rank)
        self.sbuf = [alloc(strip_elems) for _ in self.sends]
        self.rbuf = [alloc(strip_elems) for _ in self.recvs]

    def exchange(self, pack, unpack, nelem: int | None = None):
        dist = self.dist
        if not self.sends:
            return
        ops = []
        for s, b in zip(self.sends, self.sbuf):
            v = b if nelem is None else b[:nelem]
            pack(s.src_local_tile, s.src_edge, v)
        # interleave by tag so both ends of every pair enqueue in one global order
        todo = sorted([(s.tag, 0, i) for i, s in enumerate(self.sends)] + [(r.tag, 1, i) for i, r in enumerate(self.recvs)])
        for _, kind, i in todo:
            if kind == 0:
                v = self.sbuf[i] if nelem is None else self.sbuf[i][:nelem]
                ops.append(dist.P2POp(dist.isend, v, self.sends[i].dst_rank, group=self.group))
            else:
                v = self.rbuf[i] if nelem is None else self.rbuf[i][:nelem]
                ops.append(dist.P2POp(dist.irecv, v, self.recvs[i].src_rank, group=self.group))
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        for r, b in zip(self.recvs, self.rbuf):
            v = b if nelem is None else b[:nelem]
            unpack(r.dst_local_tile, r.dst_edge, v)


# ---- numpy strip packer (the CPU model of fv3t_*_halo_pack / halo_unpack used by the gloo tests) -------------------
def _halo_cells(edge: int, n: int):
    """canonical order of the halo strip beyond `edge`: depth m = 1..3 outer, position s = 1..n inner (1-based i, j)"""
    m, s = np.meshgrid(np.arange(1, cs.NG + 1), np.arange(1, n + 1), indexing="ij")
    m, s = m.ravel(), s.ravel()
    if edge == cs.W:
        return 1 - m, s
    if edge == cs.E:
        return n + m, s
    if edge == cs.S:
        return s, 1 - m
    return s, n + m


def np_pack(q_tile: np.ndarray, tile: int, edge: int, n: int) -> np.ndarray:
    """cells of `tile` (1-based) that the neighbour across `edge` needs, in the neighbour's canonical halo order.
    q_tile: [..., n+6, n+6]; returns [..., 3n]."""
    maps = cs.edge_maps(n)
    m = maps[tile - 1][edge]
    back = maps[m.nbr_tile][m.nbr_edge]
    bi, bj = _halo_cells(m.nbr_edge, n)
    si, sj = back.map_cells(bi, bj)
    return q_tile[..., sj + cs.NG - 1, si + cs.NG - 1]


def np_unpack(q_tile: np.ndarray, edge: int, n: int, strip: np.ndarray):
    i, j = _halo_cells(edge, n)
    q_tile[..., j + cs.NG - 1, i + cs.NG - 1] = strip


# ---- the face-sharded / hybrid driver on GPUs --------------------------------------------------------------------------
class ShardedTracerStep:
    """tracer_2d + tracer remap for this rank's (faces, tracers) block of one global problem.

    Mirrors fv3t_*_tracer_2d_resident but with the two collective sites of the reference made explicit: the
    mp_reduce_max of cmax (fv_tracer2d.F90:433) and the q halo update per sub-step (:499)."""

    def __init__(self, ctx, layout: Layout, rank: int, device: int):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.ctx, self.layout, self.rank = ctx, layout, rank
        self.dev = torch.device(f"cuda:{device}")
        self.tdt = torch.float64 if ctx.dtype == np.float64 else torch.float32
        peers = layout.face_peers(rank)
        self.group = None
        if layout.F > 1 and layout.G > 1:
            # one communicator per tracer group; new_group must be called by every rank for every group
            for g in range(layout.G):
                ranks = [layout.rank_of(f, g) for f in range(layout.F)]
                grp = dist.new_group(ranks)
                if rank in ranks:
                    self.group = grp
        self.peers = peers
        self.nq_local = layout.tracers(rank)[1]
        self.xch = None
        if layout.F > 1:
            nelem = 3 * ctx.n * ctx.npz * self.nq_local
            self.xch = StripExchanger(layout, rank, lambda ne: torch.empty(ne, dtype=self.tdt, device=self.dev), nelem, group=self.group)
        self.cmax_dev = torch.empty(ctx.npz, dtype=self.tdt, device=self.dev)

    def tracer_2d(self, hord: int, q_split: int = 0, lim_fac: float = 1.0) -> int:
        ctx, torch, dist = self.ctx, self.torch, self.dist
        nq = self.nq_local
        cmax = ctx.tracer_2d_begin(nq, q_split)
        if self.layout.F > 1 and q_split == 0:
            self.cmax_dev.copy_(torch.from_numpy(cmax))
            dist.all_reduce(self.cmax_dev, op=dist.ReduceOp.MAX, group=self.group)
            cmax = self.cmax_dev.cpu().numpy()
        nsplt = ctx.tracer_2d_set_cmax(cmax, q_split)
        for it in range(1, nsplt + 1):
            ctx.halo_local(it)
            if self.xch is not None:
                self.xch.exchange(lambda lt, e, b: ctx.halo_pack(it, lt, e, b.data_ptr()),
                                  lambda lt, e, b: ctx.halo_unpack(it, lt, e, b.data_ptr()))
            ctx.tracer_2d_substep(it, hord, lim_fac)
        ctx.tracer_2d_finish()
        return nsplt

    def remap(self, kord, fill=True):
        self.ctx.remap_tracers_resident(self.nq_local, kord, fill)


def bench_face_sharded(args, rank: int, world: int, local_rank: int, prefer: str = "face") -> int:
    """bench.py --shard face: ONE global problem (6 faces x nq tracers) split over the ranks as F face groups x G tracer
    groups (strong scaling).  Per sub-step one grouped NCCL send/recv of the packed edge strips between face groups and,
    per call, one all-reduce(MAX) of cmax; everything (kernels, pack/unpack, NCCL) is ordered on one CUDA stream."""
    import json
    import torch
    import torch.distributed as dist
    from . import synthetic_device as sd
    from .tracer import TracerContext

    n, npz, nq = args.n, args.npz, args.nq
    F, G = choose_layout(world, nq, prefer=prefer)
    layout = Layout(world, F, G, nq)
    tiles = layout.tiles(rank)
    q_first, nq_local = layout.tracers(rank)
    dev = torch.device(f"cuda:{local_rank}")
    stream = torch.cuda.Stream(device=dev)
    grid = cs.make_grid(n)
    w = 8 if args.dtype == "float64" else 4
    with torch.cuda.stream(stream):
        ctx = TracerContext(n + 1, npz, nq_local, grid.astype(args.dtype), dtype=args.dtype, tiles=tiles, device=local_rank,
                            stream=stream.cuda_stream)
        sd.fill_context(ctx, grid, nq_local, courant=args.courant, seed=20260101, device=local_rank, q_first=q_first)
        step = ShardedTracerStep(ctx, layout, rank, local_rank)
        kord = np.full(nq_local, args.kord, dtype=np.int32)

        def one():
            ns = step.tracer_2d(args.hord)
            step.remap(kord, fill=True)
            return ns

        # (the sampler initialises NVML: before the warm-up, so that rank 0 does not enter the timed region late)
        sampler = args.clock_sampler(local_rank) if rank == 0 and getattr(args, "clock_sampler", None) else None
        nsplt = 1
        for _ in range(max(args.warmup, 3)):
            nsplt = one()
        stream.synchronize()
        dist.barrier()
        if sampler:
            sampler.start()
        l0 = ctx.kernel_launches()
        ctx.timer_start()
        for _ in range(args.steps):
            one()
        ms = ctx.timer_stop_ms()
        launches = ctx.kernel_launches() - l0
        stream.synchronize()
        dist.barrier()
        clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    updates = 6 * n * n * npz * nq
    cells_rank = len(tiles) * n * n * npz

    # ---- per-kernel roofline of rank 0's share (separate pass with per-kernel event timing, not part of the timed region)
    roof = None
    with torch.cuda.stream(stream):
        ctx.profile_enable(True)
        for _ in range(2):
            one()
        adv_ms, adv_n = ctx.profile_get("advect")
        rm_ms, rm_n = ctx.profile_get("remap")
        halo_ms, _ = ctx.profile_get("halo")
        oth = sum(ctx.profile_get(k)[0] for k in ("cmax", "scale"))
        ctx.profile_enable(False)
        stream.synchronize()
    dist.barrier()
    if rank == 0:
        import json as _json
        import os as _os
        peak, which = 6650.0, "fallback"
        try:
            peak = float(_json.load(open(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
            which = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
        adv_bytes = cells_rank * (2 * w * nq_local + 5 * w)
        rm_bytes = cells_rank * (2 * w * nq_local + 2 * w)
        adv_avg, rm_avg = adv_ms / max(adv_n, 1), rm_ms / max(rm_n, 1)
        dom_adv = adv_ms >= rm_ms
        a_bytes, a_ms = (adv_bytes, adv_avg) if dom_adv else (rm_bytes, rm_avg)
        ach = a_bytes / (a_ms * 1e-3) / 1e9
        B = 2 * w * 2 + w * 7 / nq_local
        roof = {"bound": "hbm", "kernel": ("k_advect5" if nq_local >= 4 else "k_advect4") if dom_adv else "k_remap3", "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": None, "peak_source": which, "bytes_per_launch": a_bytes, "avg_launch_ms": a_ms,
                "rank0_ms_per_step": {"advect": adv_ms / 2, "remap": rm_ms / 2, "halo pack/unpack kernels": halo_ms / 2, "prep/coef/cmax": oth / 2},
                "step_frac_of_roofline_per_gpu": (updates / world / (ms_max / args.steps * 1e-3)) * B / (peak * 1e9)}

    # ---- end to end: every rank's share of the inputs starts in pinned host memory and its results end there
    e2e = None
    if not getattr(args, "no_e2e", False):
        from .devarray import field_shape
        fields_in = ["q", "dp1", "mfx", "mfy", "cx", "cy", "pe"]
        host = {f: torch.empty(field_shape(ctx, f, nq_local), dtype=step.tdt, pin_memory=True) for f in fields_in + ["delp"]}
        with torch.cuda.stream(stream):
            sd.fill_context(ctx, grid, nq_local, courant=args.courant, seed=20260101, device=local_rank, q_first=q_first)
            for f in fields_in:
                ctx.download_ptr(f, host[f].data_ptr(), nq_local)
            stream.synchronize()

            def e2e_step():
                for f in fields_in:
                    ctx.upload_ptr(f, host[f].data_ptr(), nq_local)
                ctx.set_vertical(ak_h, bk_h, ptop_h)
                one()
                ctx.download_ptr("q", host["q"].data_ptr(), nq_local)
                ctx.download_ptr("delp", host["delp"].data_ptr(), nq_local)

            ak_h, bk_h, ptop_h = sd.hybrid(npz)
            e2e_step()
            stream.synchronize()
            dist.barrier()
            import time as _time
            t0 = _time.perf_counter()
            for _ in range(args.e2e_steps):
                e2e_step()
            stream.synchronize()
            dist.barrier()
            wall = (_time.perf_counter() - t0) * 1e3
        et = torch.tensor([wall], dtype=torch.float64, device=dev)
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
        h2d = torch.tensor([sum(host[f].numel() * w for f in fields_in)], dtype=torch.float64, device=dev)
        d2h = torch.tensor([(host["q"].numel() + host["delp"].numel()) * w], dtype=torch.float64, device=dev)
        dist.all_reduce(h2d)
        dist.all_reduce(d2h)
        e2e = {"value": updates * args.e2e_steps / (float(et.item()) * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": int(h2d.item()),
               "d2h_bytes_per_step": int(d2h.item()), "steps": args.e2e_steps, "ms_per_step": float(et.item()) / args.e2e_steps}
        del host
    if rank == 0:
        strip_bytes = 3 * n * npz * nq_local * w
        sends, _ = strip_schedule(layout, rank)
        line = {"metric": "tracer_cell_updates_per_s", "value": updates * args.steps / (ms_max * 1e-3), "unit": "cell-updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if w == 8 else "f32",
                "data": "synthetic",
                "config": {"workload": f"C{n} L{npz}, {nq} tracers in total, {args.dtype}, hord_tr={args.hord}, kord_tr={args.kord}, fill, "
                                       f"tracer_2d + tracer remap",
                           "parallelism": f"ONE problem split over {F} face groups x {G} tracer groups; rank 0: tiles {list(tiles)}, {nq_local} tracers",
                           "halo": (f"{len(sends)} NCCL send/recv strips of {strip_bytes} B per rank and sub-step ({len(sends) * strip_bytes} B sent per rank) "
                                    f"+ all-reduce(max) of cmax" if F > 1 else
                                    "none: tracer groups only, winds / mass fluxes / delp replicated (no data-path collective)"),
                           "halo_bytes_sent_per_rank_and_substep": len(sends) * strip_bytes if F > 1 else 0,
                           "nsplt": int(nsplt), "updates_per_step": updates,
                           "l2": "inputs (GBs per rank) far exceed the 126 MB L2; no flush needed"},
                "gpu_launches": int(launches), "e2e": e2e, "roofline": roof, "cpu_baseline": None, "clocks": clocks}
        print(json.dumps(line))
    ctx.close()
    dist.destroy_process_group()
    return 0
