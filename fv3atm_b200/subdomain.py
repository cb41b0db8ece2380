"""Sub-tile decomposition of the cubed sphere: ``layout = L x L`` square sub-domains per tile, as a reference rank sees them when
the model runs with ``layout > 1,1`` (``fv_grid_bounds_type``: ``is..ie`` a proper sub-range of ``1..npx-1``, fv_arrays.F90:1178-1186;
``mpp_define_domains`` with the cubed-sphere mosaic, fv_mp_mod.F90:581-629).

What a sub-domain needs that a whole tile does not:
  * its halo comes from up to eight neighbours: the four sides AND the four diagonal neighbours where the corner of the sub-domain
    lies inside a tile or on a tile edge (``mpp_update_domains`` fills those); only at a TRUE cube corner does ``copy_corners``
    (tp_core.F90:253-330) rebuild the 3 x 3 block, and only those sub-domains carry the corner flags (fv_arrays.F90:181);
  * the tile-edge formulas of ``xppm`` / ``yppm`` apply on the sides that lie on a tile edge only.
This module slices whole-tile arrays into sub-domain arrays and builds the halo exchange as flat gather lists
``dst (sub-domain, plane offset) <- src (sub-domain, plane offset)``; the kernels get the edge / corner flags (``A5Sub`` in
csrc/fv3t_advect5.cuh).  Pure numpy: shared by the multi-GPU driver (partition.py), the CPU tests and the host-simulated kernels."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import cubed_sphere as cs

NG = cs.NG
SW, SE, NE, NW = 1, 2, 4, 8


@dataclass(frozen=True)
class Sub:
    tile: int   # 0-based
    bi: int     # block column / row inside the tile
    bj: int
    i0: int     # global cell index = local cell index + i0
    j0: int


class SubMosaic:
    def __init__(self, n: int, L: int):
        if n % L:
            raise ValueError(f"layout {L}x{L} does not divide n = {n}")
        self.n, self.L, self.m = n, L, n // L
        if self.m < 2 * NG + 2:
            raise ValueError("sub-domains narrower than 8 cells are not supported (the edge stencils of the two sides would overlap)")
        self.subs = [Sub(t, bi, bj, bi * self.m, bj * self.m) for t in range(6) for bj in range(L) for bi in range(L)]
        self._table = None

    def __len__(self):
        return len(self.subs)

    def index(self, tile: int, bi: int, bj: int) -> int:
        return (tile * self.L + bj) * self.L + bi

    # ---- kernel flags: (no_w, no_e, no_s, no_n, cmask) -------------------------------------------------------------------
    def flags(self, s: int):
        sb, L = self.subs[s], self.L
        w, e, so, no = sb.bi == 0, sb.bi == L - 1, sb.bj == 0, sb.bj == L - 1
        cmask = (SW if w and so else 0) | (SE if e and so else 0) | (NE if e and no else 0) | (NW if w and no else 0)
        return (int(not w), int(not e), int(not so), int(not no), cmask)

    # ---- slicing whole-tile arrays [6, ..., rows, cols] ---------------------------------------------------------------------
    def _cut(self, a, s, rows, cols, r_extra, c_extra):
        sb, m = self.subs[s], self.m
        return np.ascontiguousarray(a[sb.tile][..., sb.j0:sb.j0 + m + r_extra, sb.i0:sb.i0 + m + c_extra])

    def cells(self, a, s):      # (isd:ied, jsd:jed): q, dp1, area, rarea, dxa, dya, sin_sg [6, ..., nd, nd]
        return self._cut(a, s, None, None, 2 * NG, 2 * NG)

    def xface(self, a, s):      # (is:ie+1, jsd:jed): cx [6, ..., nd, n+1]
        return self._cut(a, s, None, None, 2 * NG, 1)

    def yface(self, a, s):      # (isd:ied, js:je+1): cy [6, ..., n+1, nd]
        return self._cut(a, s, None, None, 1, 2 * NG)

    def mfx(self, a, s):        # (is:ie+1, js:je) [6, ..., n, n+1]
        return self._cut(a, s, None, None, 0, 1)

    def mfy(self, a, s):        # (is:ie, js:je+1) [6, ..., n+1, n]
        return self._cut(a, s, None, None, 1, 0)

    def dx(self, a, s):         # (isd:ied, jsd:jed+1) [6, nd+1, nd]
        return self._cut(a, s, None, None, 2 * NG + 1, 2 * NG)

    def dy(self, a, s):         # (isd:ied+1, jsd:jed) [6, nd, nd+1]
        return self._cut(a, s, None, None, 2 * NG, 2 * NG + 1)

    def pe(self, a, s):         # (is-1:ie+1, km+1, js-1:je+1) stored [6, n+2 (j), km+1, n+2 (i)]
        sb, m = self.subs[s], self.m
        return np.ascontiguousarray(a[sb.tile][sb.j0:sb.j0 + m + 2, :, sb.i0:sb.i0 + m + 2])

    def metrics(self, g: dict, s: int) -> dict:
        """Sub-domain slices of the whole-tile metric dictionary of ``SyntheticCase.metrics()``."""
        out = {k: self.cells(g[k], s) for k in ("area", "rarea", "dxa", "dya", "sin_sg")}
        out["dx"] = self.dx(g["dx"], s)
        out["dy"] = self.dy(g["dy"], s)
        return out

    def put_interior(self, whole, s, local):
        """whole[tile, ..., interior of sub-domain s] <- local[..., interior]"""
        sb, m = self.subs[s], self.m
        whole[sb.tile][..., sb.j0 + NG:sb.j0 + NG + m, sb.i0 + NG:sb.i0 + NG + m] = local[..., NG:NG + m, NG:NG + m]

    # ---- halo exchange ------------------------------------------------------------------------------------------------------
    def halo_table(self):
        """Gather lists of the scalar halo update of every sub-domain: int64 arrays (dst_sub, dst_off, src_sub, src_off), offsets
        into the (m+6) x (m+6) planes.  Covers the side halos and the diagonal blocks that are not true cube corners."""
        if self._table is not None:
            return self._table
        n, m, L = self.n, self.m, self.L
        nd, md = n + 2 * NG, m + 2 * NG
        # source of every cell of the halo-extended whole tiles, as a flat index into the [6, nd, nd] stack (-1: true corner block)
        G = np.full((6, nd, nd), -1, dtype=np.int64)
        own = np.arange(6 * nd * nd, dtype=np.int64).reshape(6, nd, nd)
        G[:, NG:NG + n, NG:NG + n] = own[:, NG:NG + n, NG:NG + n]
        dt, dj, di, st, sj, si = cs.halo_index_table(n)
        G[dt, dj, di] = (st.astype(np.int64) * nd + sj) * nd + si
        ds, do, ss, so = [], [], [], []
        aj, ai = np.meshgrid(np.arange(md), np.arange(md), indexing="ij")
        halo = ~((aj >= NG) & (aj < NG + m) & (ai >= NG) & (ai < NG + m))
        for s, sb in enumerate(self.subs):
            g = G[sb.tile, sb.j0:sb.j0 + md, sb.i0:sb.i0 + md]
            sel = halo & (g >= 0)
            src = g[sel]
            t2, r2 = np.divmod(src, nd * nd)
            j2, i2 = np.divmod(r2, nd)          # array coordinates in the whole source tile: interior cells
            bj2, bi2 = (j2 - NG) // m, (i2 - NG) // m
            ds.append(np.full(src.size, s, dtype=np.int64))
            do.append((aj[sel] * md + ai[sel]).astype(np.int64))
            ss.append((t2 * L + bj2) * L + bi2)
            so.append((j2 - bj2 * m) * md + (i2 - bi2 * m))
        self._table = tuple(np.concatenate(a) for a in (ds, do, ss, so))
        return self._table

    def pair_lists(self):
        """The exchange grouped by (dst sub-domain, src sub-domain): {(d, s): (dst_off int32[], src_off int32[])}."""
        ds, do, ss, so = self.halo_table()
        out = {}
        key = ds * len(self) + ss
        order = np.argsort(key, kind="stable")
        ks, starts = np.unique(key[order], return_index=True)
        ends = list(starts[1:]) + [key.size]
        for k, a, b in zip(ks, starts, ends):
            idx = order[a:b]
            out[(int(k // len(self)), int(k % len(self)))] = (do[idx].astype(np.int32), so[idx].astype(np.int32))
        return out

    def fill_halos(self, stack: np.ndarray) -> np.ndarray:
        """numpy reference of the exchange: stack [nsub, ..., m+6, m+6] in place."""
        ds, do, ss, so = self.halo_table()
        md = self.m + 2 * NG
        flat = stack.reshape(stack.shape[0], -1, md * md)
        flat[ds, :, do] = flat[ss, :, so]
        return stack


# ---- distribution over ranks and contexts ------------------------------------------------------------------------------------
def assign(nsub: int, world: int, max_per_ctx: int = 6):
    """owner[s] = (rank, context index on that rank, local tile in that context): contiguous blocks of sub-domains per rank (so that
    the sub-domains of a face stay together), split into contexts of at most six resident sub-domains (fv3t_dims.tile_id[6])."""
    if nsub % world:
        raise ValueError(f"{nsub} sub-domains do not divide over {world} ranks")
    per = nsub // world
    owner = []
    for s in range(nsub):
        r, k = divmod(s, per)
        owner.append((r, k // max_per_ctx, k % max_per_ctx))
    return owner


class ExchangePlan:
    """The halo update of a sub-domain mosaic from the point of view of one rank, in three classes:
      local[c]   = (dst_flat, src_flat): both sub-domains resident in context c -> fv3t_halo_local_table / fv3t_*_halo_local
      copies     = [(src_ctx, src_lt, src_offs, dst_ctx, dst_lt, dst_offs)]: same rank, different contexts
      sends[p] / recvs[p] = [(ctx, lt, offs)] in one global order (sorted by (dst sub, src sub)): one message per peer rank p,
                   the pieces concatenated, each piece `planes x len(offs)` values."""

    def __init__(self, mo: SubMosaic, owner, rank: int):
        self.rank = rank
        md = mo.m + 2 * NG
        plane = md * md
        self.local, self.copies, self.sends, self.recvs = {}, [], {}, {}
        loc = {}
        for (d, s), (doff, soff) in sorted(mo.pair_lists().items()):
            rd, cd, ld = owner[d]
            rs, cs_, ls = owner[s]
            if rd == rank and rs == rank:
                if cd == cs_:
                    a = loc.setdefault(cd, ([], []))
                    a[0].append(doff.astype(np.int64) + ld * plane)
                    a[1].append(soff.astype(np.int64) + ls * plane)
                else:
                    self.copies.append((cs_, ls, soff, cd, ld, doff))
            elif rs == rank:
                self.sends.setdefault(rd, []).append((cs_, ls, soff))
            elif rd == rank:
                self.recvs.setdefault(rs, []).append((cd, ld, doff))
        for c, (a, b) in loc.items():
            self.local[c] = (np.concatenate(a).astype(np.int32), np.concatenate(b).astype(np.int32))
        # consecutive pieces of a message that live in the same context become ONE flat list (local_tile = -1): one gather / scatter
        # launch per context and peer instead of one per pair of sub-domains
        def merge(pieces):
            out = []
            for c, lt, offs in pieces:
                flat = offs.astype(np.int64) + lt * plane
                if out and out[-1][0] == c:
                    out[-1][1].append(flat)
                else:
                    out.append([c, [flat]])
            return [(c, -1, np.concatenate(fl).astype(np.int32)) for c, fl in out]
        self.sends = {p: merge(v) for p, v in self.sends.items()}
        self.recvs = {p: merge(v) for p, v in self.recvs.items()}

    def message_cells(self, peer: int, sending: bool) -> int:
        return sum(len(o) for _, _, o in (self.sends if sending else self.recvs).get(peer, []))


def run_exchange(plan: ExchangePlan, planes: int, copy_buf, sbuf: dict, rbuf: dict, gather, scatter, group=None):
    """The cross-context and cross-rank part of one halo update.  gather(ctx, local_tile, offsets, buf, at, stride) writes
    buf[pl * stride + at + e] for every plane pl and list entry e, scatter is its inverse; a message is plane-major over ALL its
    cells (stride = cells of the message), so its layout does not depend on how either side groups the pieces into launches.
    The buffers are flat torch tensors (CUDA with NCCL, CPU with gloo)."""
    for (cs_, ls, soff, cd, ld, doff) in plan.copies:
        gather(cs_, ls, soff, copy_buf, 0, len(soff))
        scatter(cd, ld, doff, copy_buf, 0, len(soff))
    if not plan.sends and not plan.recvs:
        return
    import torch.distributed as dist
    for p, pieces in plan.sends.items():
        at, total = 0, plan.message_cells(p, True)
        for (c, lt, offs) in pieces:
            gather(c, lt, offs, sbuf[p], at, total)
            at += len(offs)
    ops = []
    for p in sorted(set(plan.sends) | set(plan.recvs)):
        for kind in ((0, 1) if plan.rank < p else (1, 0)):  # both ends of a pair enqueue in one order
            if kind == 0 and p in sbuf:
                ops.append(dist.P2POp(dist.isend, sbuf[p], p, group=group))
            if kind == 1 and p in rbuf:
                ops.append(dist.P2POp(dist.irecv, rbuf[p], p, group=group))
    for w in dist.batch_isend_irecv(ops):
        w.wait()
    for p, pieces in plan.recvs.items():
        at, total = 0, plan.message_cells(p, False)
        for (c, lt, offs) in pieces:
            scatter(c, lt, offs, rbuf[p], at, total)
            at += len(offs)


class SubMosaicStep:
    """tracer_2d + tracer remap of ONE global problem decomposed into L x L sub-domains per tile, this rank's share resident in
    one or more sub-tile contexts (fv3t_dims.sub_layout).  The two collective sites of the reference are explicit: the
    mp_reduce_max of cmax (fv_tracer2d.F90:433) and the q halo update per sub-step (:499) -- here gather lists that also carry
    the diagonal blocks, like mpp_update_domains does for a rank that owns part of a tile."""

    def __init__(self, mo: SubMosaic, rank: int, world: int, device: int, npz: int, nq: int, dtype, metrics: dict, group=None):
        import torch
        from .tracer import TracerContext
        self.torch = torch
        self.mo, self.rank, self.world, self.group = mo, rank, world, group
        self.npz, self.nq = npz, nq
        self.dev = torch.device(f"cuda:{device}")
        self.tdt = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        self.owner = assign(len(mo), world)
        self.mine = [s for s in range(len(mo)) if self.owner[s][0] == rank]
        nctx = 1 + max(self.owner[s][1] for s in self.mine)
        self.ctx_subs = [[s for s in self.mine if self.owner[s][1] == c] for c in range(nctx)]
        # ONE stream for every context of the rank and for the collectives: the cross-context copies and the packed messages are
        # ordered by the stream alone (a context without a stream handle creates a private one)
        self.stream = torch.cuda.Stream(self.dev)
        stream = self.stream.cuda_stream
        self.ctxs = []
        for subs in self.ctx_subs:
            g = {k: np.stack([mo.metrics(metrics, s)[k] for s in subs]) for k in ("area", "rarea", "dx", "dy", "dxa", "dya", "sin_sg")}
            self.ctxs.append(TracerContext(mo.m + 1, npz, nq, g, dtype=dtype, tiles=[mo.subs[s].tile + 1 for s in subs], device=device,
                                           stream=stream, sub_layout=mo.L, sub_blocks=[(mo.subs[s].bi, mo.subs[s].bj) for s in subs]))
        self.plan = ExchangePlan(mo, self.owner, rank)
        for c, (dst, src) in self.plan.local.items():
            self.ctxs[c].halo_local_table(dst, src)
        planes = npz * nq
        mk = lambda ne: torch.empty(max(ne, 1), dtype=self.tdt, device=self.dev)
        self._lists = {}
        self.copy_buf = mk(planes * max([len(o) for _, _, o, _, _, _ in self.plan.copies] + [0]))
        self.sbuf = {p: mk(planes * self.plan.message_cells(p, True)) for p in self.plan.sends}
        self.rbuf = {p: mk(planes * self.plan.message_cells(p, False)) for p in self.plan.recvs}
        self.cmax_dev = torch.empty(npz, dtype=self.tdt, device=self.dev)
        self.halo_bytes = sum(b.numel() for b in self.sbuf.values()) * self.cmax_dev.element_size()

    def _list(self, c, offs):
        key = (c, offs.ctypes.data)
        if key not in self._lists:
            self._lists[key] = (self.ctxs[c].halo_list_create(offs), offs)  # keep offs alive: its address is the key
        return self._lists[key][0]

    def upload(self, field: str, whole: np.ndarray, cut):
        """whole-tile array -> the resident sub-domain slices (cut = the SubMosaic slicer of that field)"""
        for c, subs in enumerate(self.ctx_subs):
            a = np.ascontiguousarray(np.stack([cut(whole, s) for s in subs]))
            self.ctxs[c].upload(field, a, nq=self.nq if field == "q" else None)

    def upload_case(self, case):
        mo = self.mo
        self.upload("q", case.q, mo.cells)
        self.upload("dp1", case.dp1, mo.cells)
        self.upload("cx", case.cx, mo.xface)
        self.upload("cy", case.cy, mo.yface)
        self.upload("mfx", case.mfx, mo.mfx)
        self.upload("mfy", case.mfy, mo.mfy)
        self.upload("pe", case.pe, mo.pe)
        for ctx in self.ctxs:
            ctx.set_vertical(case.ak, case.bk, case.ptop)

    def download(self, field: str, whole: np.ndarray):
        """interiors of the resident sub-domains -> the whole-tile array (tile axis first)"""
        md = self.mo.m + 2 * NG
        for c, subs in enumerate(self.ctx_subs):
            shape = (len(subs),) + whole.shape[1:-2] + (md, md)
            a = np.zeros(shape, dtype=whole.dtype)
            self.ctxs[c].download(field, a, nq=self.nq if field == "q" else None)
            for k, s in enumerate(subs):
                self.mo.put_interior(whole, s, a[k])
        return whole

    def exchange(self, it: int):
        with self.torch.cuda.stream(self.stream):
            self._exchange(it)

    def _exchange(self, it: int):
        for ctx in self.ctxs:
            ctx.halo_local(it)
        esz = self.cmax_dev.element_size()
        gather = lambda c, lt, offs, buf, at, stride: self.ctxs[c].halo_gather(it, lt, self._list(c, offs), buf.data_ptr() + at * esz, stride)
        scatter = lambda c, lt, offs, buf, at, stride: self.ctxs[c].halo_scatter(it, lt, self._list(c, offs), buf.data_ptr() + at * esz, stride)
        run_exchange(self.plan, self.npz * self.nq, self.copy_buf, self.sbuf, self.rbuf, gather, scatter, self.group)

    def _mark(self, name):
        """phase marks on the stream (bench diagnostics: SubMosaicStep.trace = [] switches them on)"""
        tr = getattr(self, "trace", None)
        if tr is not None:
            import time
            e = self.torch.cuda.Event(enable_timing=True)
            e.record(self.stream)
            tr.append((name, e, time.perf_counter()))

    def tracer_2d(self, hord: int, q_split: int = 0, lim_fac: float = 1.0) -> int:
        torch = self.torch
        self._mark("start")
        # The halo update of the first sub-step needs neither cmax nor ksplt (every level takes part, and outside a tracer_2d call
        # every level lives in the current buffer): it is queued FIRST, behind whatever the stream is still running, so that its
        # launches and the NCCL transfer overlap the host round trips of the cmax reduction instead of following them.
        # (Not on the very first call: tracer_2d_begin is what tells a context how many tracers are resident.)
        early = getattr(self, "_primed", False)
        if early:
            self.exchange(1)
        self._mark("halo1")
        cm = None
        for ctx in self.ctxs:
            c = ctx.tracer_2d_begin(self.nq, q_split)
            cm = c if cm is None else np.maximum(cm, c)
        self._primed = True
        if not early:
            self.exchange(1)
        if self.world > 1 and q_split == 0:
            import torch.distributed as dist
            with torch.cuda.stream(self.stream):
                self.cmax_dev.copy_(torch.from_numpy(np.ascontiguousarray(cm)))
                dist.all_reduce(self.cmax_dev, op=dist.ReduceOp.MAX, group=self.group)
                cm = self.cmax_dev.cpu().numpy()
        self._mark("cmax")
        nsplt = 0
        for ctx in self.ctxs:
            nsplt = ctx.tracer_2d_set_cmax(cm, q_split)
        for it in range(1, nsplt + 1):
            if it > 1:
                self.exchange(it)
            for ctx in self.ctxs:
                ctx.tracer_2d_substep(it, hord, lim_fac)
        for ctx in self.ctxs:
            ctx.tracer_2d_finish()
        self._mark("advect")
        return nsplt

    def remap(self, kord, fill=True):
        for ctx in self.ctxs:
            ctx.remap_tracers_resident(self.nq, kord, fill)
        self._mark("remap")

    def close(self):
        for ctx in self.ctxs:
            ctx.close()


# ---- bench.py --shard sub ---------------------------------------------------------------------------------------------------------
def fill_submosaic(run: SubMosaicStep, grid, args, device: int):
    """Synthetic inputs of the resident sub-domains, generated on the device: each tile this rank touches is generated once in a
    temporary whole-tile context (the generator of bench.py's single-GPU workload) and its windows are copied into the sub-tile
    contexts."""
    import torch
    from . import synthetic_device as sd
    from .devarray import field_view
    from .tracer import TracerContext
    mo, n, m = run.mo, run.mo.n, run.mo.m
    win = {"q": (6, 6), "dp1": (6, 6), "cx": (6, 1), "cy": (1, 6), "mfx": (0, 1), "mfy": (1, 0)}
    ak = bk = ptop = None
    with torch.cuda.stream(run.stream):
        for tile in sorted({mo.subs[s].tile for s in run.mine}):
            w = TracerContext(n + 1, run.npz, run.nq, grid.astype(args.dtype), dtype=args.dtype, tiles=[tile + 1], device=device,
                              stream=run.stream.cuda_stream)
            ak, bk, ptop = sd.fill_context(w, grid, run.nq, courant=args.courant, seed=20260101, device=device)
            for c, subs in enumerate(run.ctx_subs):
                for lt, s in enumerate(subs):
                    sb = mo.subs[s]
                    if sb.tile != tile:
                        continue
                    for f, (re, ce) in win.items():
                        nq = run.nq if f == "q" else None
                        field_view(run.ctxs[c], f, nq, device)[lt].copy_(
                            field_view(w, f, nq, device)[0][..., sb.j0:sb.j0 + m + re, sb.i0:sb.i0 + m + ce])
                    field_view(run.ctxs[c], "pe", None, device)[lt].copy_(
                        field_view(w, "pe", None, device)[0][sb.j0:sb.j0 + m + 2, :, sb.i0:sb.i0 + m + 2])
            run.stream.synchronize()
            w.close()
        for ctx in run.ctxs:
            ctx.set_vertical(ak, bk, ptop)
    return ak, bk, ptop


def bench_submosaic(args, rank: int, world: int, local_rank: int) -> int:
    """bench.py --shard sub: ONE global problem decomposed into L x L sub-domains per tile (24 for L = 2: SURVEY.md section 8e), the
    sub-domains of a rank resident in sub-tile contexts; per sub-step one packed NCCL message per peer pair (side halos and diagonal
    blocks), per call one all-reduce(max) of cmax; kernels, packing and collectives ordered on one CUDA stream per rank."""
    import json
    import os
    import time
    import torch
    import torch.distributed as dist

    n, npz, nq, L = args.n, args.npz, args.nq, args.sub_layout
    dev = torch.device(f"cuda:{local_rank}")
    grid = cs.make_grid(n)
    w = 8 if args.dtype == "float64" else 4
    mo = SubMosaic(n, L)
    run = SubMosaicStep(mo, rank, world, local_rank, npz, nq, args.dtype, grid.astype(args.dtype))
    ak, bk, ptop = fill_submosaic(run, grid, args, local_rank)
    kord = np.full(nq, args.kord, dtype=np.int32)

    def one():
        ns = run.tracer_2d(args.hord)
        run.remap(kord, fill=True)
        return ns

    def barrier():
        run.stream.synchronize()
        if world > 1:
            dist.barrier()

    # (the sampler initialises NVML: before the warm-up, so that rank 0 does not enter the timed region late)
    sampler = args.clock_sampler(local_rank) if rank == 0 and getattr(args, "clock_sampler", None) else None
    nsplt = 1
    for _ in range(max(args.warmup, 3)):
        nsplt = one()
    barrier()
    if sampler:
        sampler.start()
    l0 = sum(c.kernel_launches() for c in run.ctxs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(run.stream)
    for _ in range(args.steps):
        one()
    e1.record(run.stream)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sum(c.kernel_launches() for c in run.ctxs) - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    updates = 6 * n * n * npz * nq
    cells_rank = len(run.mine) * mo.m * mo.m * npz

    # where the step time goes besides the kernels: the halo update alone and the cmax reduction alone (events on the stream)
    def timed(fn, reps=10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        a.record(run.stream)
        for _ in range(reps):
            fn()
        b.record(run.stream)
        barrier()
        v = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    def cmax_only():
        cm = None
        for ctx in run.ctxs:
            c = ctx.tracer_2d_begin(nq, 0)
            cm = c if cm is None else np.maximum(cm, c)
        if world > 1:
            with torch.cuda.stream(run.stream):
                run.cmax_dev.copy_(torch.from_numpy(np.ascontiguousarray(cm)))
                dist.all_reduce(run.cmax_dev, op=dist.ReduceOp.MAX)
                run.cmax_dev.cpu()

    # phase marks of three traced steps: GPU time between the marks (events on the stream) and host time between them
    phases = None
    run.trace = []
    for _ in range(3):
        one()
    barrier()
    tr, run.trace = run.trace, None
    if rank == 0:
        acc = {}
        for (n0, ev0, h0), (n1, ev1, h1) in zip(tr, tr[1:]):
            a = acc.setdefault(f"{n0}->{n1}", [0.0, 0.0, 0])
            a[0] += ev0.elapsed_time(ev1)
            a[1] += (h1 - h0) * 1e3
            a[2] += 1
        phases = {k: {"gpu_ms": round(v[0] / v[2], 3), "host_ms": round(v[1] / v[2], 3)} for k, v in acc.items()}
    exchange_ms = timed(lambda: run.exchange(1))
    cmax_ms = timed(cmax_only)
    for c in run.ctxs:
        c.profile_enable(True)
    for _ in range(2):
        one()
    prof = {k: sum(c.profile_get(k)[0] for c in run.ctxs) / 2 for k in ("advect", "remap", "halo", "cmax", "scale")}
    adv_n = sum(c.profile_get("advect")[1] for c in run.ctxs)
    for c in run.ctxs:
        c.profile_enable(False)
    barrier()
    # kernel time per step of EVERY rank (the ranks meet twice per step, so the slowest one sets the pace)
    ksum = torch.tensor([sum(prof.values())], dtype=torch.float64, device=dev)
    kall = [torch.zeros_like(ksum) for _ in range(world)]
    if world > 1:
        dist.all_gather(kall, ksum)
    else:
        kall = [ksum]
    kernel_ms_by_rank = [round(float(v.item()), 3) for v in kall]
    roof = None
    if rank == 0:
        peak, which = 6650.0, "fallback"
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
            which = "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
        adv_bytes = cells_rank * (2 * w * nq + 5 * w)
        ach = adv_bytes / (prof["advect"] * 1e-3) / 1e9
        B = 2 * w * 2 + w * 7 / nq
        roof = {"bound": "hbm", "kernel": "k_advect5", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                "peak_source": which, "bytes_per_launch": adv_bytes / max(adv_n // 2, 1) * 1.0, "avg_launch_ms": prof["advect"] / max(adv_n // 2, 1),
                "rank0_ms_per_step": {"advect": prof["advect"], "remap": prof["remap"], "halo gather/scatter kernels": prof["halo"],
                                      "prep/coef/cmax": prof["cmax"] + prof["scale"]},
                "step_frac_of_roofline_per_gpu": (updates / world / (ms_max / args.steps * 1e-3)) * B / (peak * 1e9)}

    e2e = None
    if not getattr(args, "no_e2e", False):
        from .devarray import field_shape
        fields_in = ["q", "dp1", "mfx", "mfy", "cx", "cy", "pe"]
        tdt = run.tdt
        host = [{f: torch.empty(field_shape(c, f, nq), dtype=tdt, pin_memory=True) for f in fields_in + ["delp"]} for c in run.ctxs]
        fill_submosaic(run, grid, args, local_rank)
        for c, h in zip(run.ctxs, host):
            for f in fields_in:
                c.download_ptr(f, h[f].data_ptr(), nq)
        run.stream.synchronize()

        def e2e_step():
            for c, h in zip(run.ctxs, host):
                for f in fields_in:
                    c.upload_ptr(f, h[f].data_ptr(), nq)
                c.set_vertical(ak, bk, ptop)
            one()
            for c, h in zip(run.ctxs, host):
                c.download_ptr("q", h["q"].data_ptr(), nq)
                c.download_ptr("delp", h["delp"].data_ptr(), nq)

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        et = torch.tensor([wall, sum(h[f].numel() * w for h in host for f in fields_in),
                           sum((h["q"].numel() + h["delp"].numel()) * w for h in host)], dtype=torch.float64, device=dev)
        mx = et.clone()
        if world > 1:
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            dist.all_reduce(et)
        e2e = {"value": updates * args.e2e_steps / (float(mx[0].item()) * 1e-3), "unit": "cell-updates/s", "h2d_bytes_per_step": int(et[1].item()),
               "d2h_bytes_per_step": int(et[2].item()), "steps": args.e2e_steps, "ms_per_step": float(mx[0].item()) / args.e2e_steps}
        del host
    if rank == 0:
        line = {"metric": "tracer_cell_updates_per_s", "value": updates * args.steps / (ms_max * 1e-3), "unit": "cell-updates/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64" if w == 8 else "f32", "data": "synthetic",
                "config": {"workload": f"C{n} L{npz}, {nq} tracers in total, {args.dtype}, hord_tr={args.hord}, kord_tr={args.kord}, fill, "
                                       f"tracer_2d + tracer remap",
                           "parallelism": f"ONE problem as {len(mo)} sub-domains (layout {L}x{L} per tile, C{mo.m} each), {len(run.mine)} per rank in "
                                          f"{len(run.ctxs)} sub-tile context(s), all {nq} tracers on every rank",
                           "halo": f"per sub-step one packed NCCL send/recv per peer pair ({len(run.plan.sends)} peers of rank 0, {run.halo_bytes} B sent by "
                                   f"rank 0: side halos + diagonal blocks as gather lists) + all-reduce(max) of cmax",
                           "halo_bytes_sent_per_rank_and_substep": run.halo_bytes, "halo_update_ms": exchange_ms, "cmax_reduction_ms": cmax_ms,
                           "kernel_ms_per_step_by_rank": kernel_ms_by_rank, "rank0_phases": phases,
                           "nsplt": int(nsplt), "updates_per_step": updates,
                           "l2": "inputs (GBs per rank) far exceed the 126 MB L2; no flush needed"},
                "gpu_launches": int(launches), "e2e": e2e, "roofline": roof, "cpu_baseline": None, "clocks": clocks}
        print(json.dumps(line))
    run.close()
    if world > 1:
        dist.destroy_process_group()
    return 0
