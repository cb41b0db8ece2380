"""Host logic of the multi-GPU path on CPU: layouts, the strip schedule, and the real exchange code
(partition.StripExchanger) run by 2 and 3 processes over the gloo backend against the numpy halo fill of the
whole mosaic (cubed_sphere.fill_edge_halos = the exchange mpp_update_domains performs, fv_tracer2d.F90:499)."""
import os
import socket

import numpy as np
import pytest

from fv3atm_b200 import cubed_sphere as cs
from fv3atm_b200 import partition as pt


def test_tracer_groups():
    assert pt.tracer_groups(30, 8) == [(0, 4), (4, 4), (8, 4), (12, 4), (16, 4), (20, 4), (24, 3), (27, 3)]
    assert pt.tracer_groups(9, 1) == [(0, 9)]
    assert [c for _, c in pt.tracer_groups(9, 4)] == [3, 2, 2, 2]
    with pytest.raises(ValueError):
        pt.tracer_groups(3, 4)


def test_choose_layout():
    assert pt.choose_layout(8, 9) == (2, 4)
    assert pt.choose_layout(4, 9) == (2, 2)
    assert pt.choose_layout(2, 9) == (2, 1)
    assert pt.choose_layout(6, 9) == (6, 1)
    assert pt.choose_layout(8, 30, prefer="tracer") == (1, 8)
    with pytest.raises(ValueError):
        pt.choose_layout(7, 3)


@pytest.mark.parametrize("F,G", [(1, 1), (2, 1), (3, 1), (6, 1), (2, 4), (3, 2), (1, 8)])
def test_strip_schedule_is_consistent(F, G):
    """Every send has exactly one matching recv on the peer, in the same per-pair order; every directed tile edge
    is covered exactly once by either a local fill or a message."""
    lay = pt.Layout(F * G, F, G, 9)
    sched = {r: pt.strip_schedule(lay, r) for r in range(lay.world)}
    maps = cs.edge_maps(8)
    for r in range(lay.world):
        sends, recvs = sched[r]
        assert len(sends) == len(recvs)
        for peer in range(lay.world):
            mine = [(s.tag, s.src_local_tile, s.src_edge, s.dst_local_tile, s.dst_edge) for s in sends if s.dst_rank == peer]
            theirs = [(s.tag, s.src_local_tile, s.src_edge, s.dst_local_tile, s.dst_edge) for s in sched[peer][1] if s.src_rank == r]
            assert mine == theirs
            if mine:
                assert lay.coords(peer)[1] == lay.coords(r)[1], "strips never cross tracer groups"
        tiles = lay.tiles(r)
        remote = {(tiles[m.dst_local_tile], m.dst_edge) for m in recvs}
        local = {(t, e) for t in tiles for e in range(4) if maps[t - 1][e].nbr_tile + 1 in tiles}
        assert remote.isdisjoint(local)
        assert remote | local == {(t, e) for t in tiles for e in range(4)}
    if F == 1:
        assert all(not sched[r][0] for r in sched)


def test_np_pack_unpack_equals_halo_fill():
    n = 12
    rng = np.random.default_rng(3)
    q = rng.standard_normal((6, 2, n + 6, n + 6))
    ref = cs.fill_edge_halos(q.copy(), n)
    got = q.copy()
    maps = cs.edge_maps(n)
    for t in range(1, 7):
        for e in range(4):
            m = maps[t - 1][e]
            strip = pt.np_pack(q[t - 1], t, e, n)                     # from tile t towards its neighbour across e
            pt.np_unpack(got[m.nbr_tile], m.nbr_edge, n, strip)
    assert np.array_equal(got, ref)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, F, G, n, nq, npz, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = pt.Layout(world, F, G, nq)
        tiles = lay.tiles(rank)
        q0, cnt = lay.tracers(rank)
        rng = np.random.default_rng(11)
        full = rng.standard_normal((6, nq, npz, n + 6, n + 6))      # same on every rank
        mine = np.ascontiguousarray(full[[t - 1 for t in tiles]][:, q0:q0 + cnt])
        # local (both tiles resident) fills, the model of fv3t_*_halo_local
        maps = cs.edge_maps(n)
        src = mine.copy()
        for lt, t in enumerate(tiles):
            for e in range(4):
                m = maps[t - 1][e]
                if m.nbr_tile + 1 in tiles:
                    strip = pt.np_pack(src[lt], t, e, n)
                    pt.np_unpack(mine[tiles.index(m.nbr_tile + 1)], m.nbr_edge, n, strip)
        group = None
        if F > 1 and G > 1:
            for g in range(G):
                ranks = [lay.rank_of(f, g) for f in range(F)]
                grp = dist.new_group(ranks)
                if rank in ranks:
                    group = grp
        nelem = 3 * n * npz * cnt
        xch = pt.StripExchanger(lay, rank, lambda ne: torch.empty(ne, dtype=torch.float64), nelem, group=group)

        def pack(lt, e, buf):
            buf.copy_(torch.from_numpy(np.ascontiguousarray(pt.np_pack(src[lt], tiles[lt], e, n)).ravel()))

        def unpack(lt, e, buf):
            pt.np_unpack(mine[lt], e, n, buf.numpy().reshape(cnt, npz, 3 * n))

        xch.exchange(pack, unpack)
        # cmax all-reduce (fv_tracer2d.F90:433) over the face peers
        cm = torch.tensor([float(rank + 1), 0.5], dtype=torch.float64)
        if F > 1:
            dist.all_reduce(cm, op=dist.ReduceOp.MAX, group=group)
        ref = cs.fill_edge_halos(full.copy(), n)[[t - 1 for t in tiles]][:, q0:q0 + cnt]
        ok = np.array_equal(mine, ref) and cm[0].item() == float(max(lay.face_peers(rank)) + 1)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,F,G", [(2, 2, 1), (2, 1, 2), (3, 3, 1), (4, 2, 2)])
def test_strip_exchange_gloo(tmp_path, world, F, G):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, F, G, 10, 4, 3, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").read_text() == "1", f"rank {r}"


# ---- sub-tile decomposition: the exchange plan of fv3atm_b200/subdomain.py run by several processes over gloo ----------------
def test_submosaic_assignment_and_plan_consistency():
    from fv3atm_b200.subdomain import ExchangePlan, SubMosaic, assign
    mo = SubMosaic(24, 2)
    total = mo.halo_table()[0].size
    for world in (1, 2, 3, 4, 6, 8, 12, 24):
        owner = assign(len(mo), world)
        assert max(c for _, c, _ in owner) == (len(mo) // world - 1) // 6 and max(lt for _, _, lt in owner) <= 5
        plans = [ExchangePlan(mo, owner, r) for r in range(world)]
        cells = 0
        for r, pl in enumerate(plans):
            cells += sum(len(d) for d, _ in pl.local.values()) + sum(len(o) for *_, o in pl.copies)
            for p in pl.sends:
                assert pl.message_cells(p, True) == plans[p].message_cells(r, False)    # (pieces are merged per context: totals match)
                cells += pl.message_cells(p, True)
        assert cells == total          # every halo cell is filled exactly once, by exactly one of the three mechanisms
    with pytest.raises(ValueError):
        assign(len(mo), 5)
    with pytest.raises(ValueError):
        SubMosaic(24, 5)
    with pytest.raises(ValueError):
        SubMosaic(24, 4)               # 6-cell sub-domains: narrower than the two edge stencils


def _sub_worker(rank, world, port, n, L, planes, out_dir):
    import torch
    import torch.distributed as dist
    from fv3atm_b200.subdomain import ExchangePlan, SubMosaic, assign, run_exchange
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mo = SubMosaic(n, L)
        md = mo.m + 6
        owner = assign(len(mo), world)
        plan = ExchangePlan(mo, owner, rank)
        rng = np.random.default_rng(5)
        whole = rng.standard_normal((6, planes, n + 6, n + 6))          # same on every rank
        mine = [s for s in range(len(mo)) if owner[s][0] == rank]
        nctx = 1 + max(owner[s][1] for s in mine)
        ctx = [np.stack([mo.cells(whole, s) for s in mine if owner[s][1] == c]) for c in range(nctx)]   # [nt_c, planes, md, md]
        interior = np.zeros((md, md), bool)
        interior[3:-3, 3:-3] = True
        for a in ctx:
            a[:, :, ~interior] = np.nan
        for c, (dst, src) in plan.local.items():                         # the model of fv3t_*_halo_local with the host's table
            flat = ctx[c].transpose(1, 0, 2, 3).reshape(planes, -1)
            flat[:, dst] = flat[:, src]
            ctx[c][:] = flat.reshape(planes, ctx[c].shape[0], md, md).transpose(1, 0, 2, 3)

        def cells(lt, offs):   # lt = -1: flat offsets local_tile * plane + offset (fv3t_*_halo_gather's rule)
            return (np.full(len(offs), lt), offs) if lt >= 0 else np.divmod(offs, md * md)

        def gather(c, lt, offs, buf, at, stride):
            tt, r = cells(lt, offs)
            buf.numpy()[:planes * stride].reshape(planes, stride)[:, at:at + len(offs)] = ctx[c].reshape(-1, planes, md * md)[tt, :, r].T

        def scatter(c, lt, offs, buf, at, stride):
            tt, r = cells(lt, offs)
            ctx[c].reshape(-1, planes, md * md)[tt, :, r] = buf.numpy()[:planes * stride].reshape(planes, stride)[:, at:at + len(offs)].T

        mk = lambda ne: torch.empty(max(ne, 1), dtype=torch.float64)
        copy_buf = mk(planes * max([len(o) for *_, o in plan.copies] + [0]))
        sbuf = {p: mk(planes * plan.message_cells(p, True)) for p in plan.sends}
        rbuf = {p: mk(planes * plan.message_cells(p, False)) for p in plan.recvs}
        run_exchange(plan, planes, copy_buf, sbuf, rbuf, gather, scatter)
        ref = cs.fill_edge_halos(whole.copy(), n)
        ref[:, :, :3, :3] = ref[:, :, :3, -3:] = ref[:, :, -3:, :3] = ref[:, :, -3:, -3:] = np.nan   # true corner blocks: copy_corners' job
        ok = True
        for s in mine:
            got, want = ctx[owner[s][1]][owner[s][2]], mo.cells(ref, s)
            ok = ok and np.array_equal(got, want, equal_nan=True)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,L", [(2, 2), (3, 2), (2, 3)])
def test_submosaic_exchange_gloo(tmp_path, world, L):
    """24 (54) sub-domains over 2 or 3 processes: local tables, cross-context copies and one packed message per peer pair
    reproduce the whole-mosaic halo fill on every sub-domain halo cell, diagonal blocks included."""
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_sub_worker, args=(world, port, 24, L, 3, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert (tmp_path / f"ok{r}").read_text() == "1", f"rank {r}"
