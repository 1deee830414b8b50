"""N > 1 host logic on CPU: world_size-2/3 gloo runs of the slab assembly (migration + halo) and the field
halo exchange, with a numpy stand-in for the engine's neighbour search (no GPU, no oracle)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


CX, CY, CZ = 9, 3, 3          # global grid, cell length 1


STRAY_CAP = 96


def _worker(rank, world, port, steps, q, bullets=False):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pkgload
    pkgload.load()
    from cpp_fluid_particles_b200 import slabs
    rng = np.random.default_rng(123)                      # same stream on every rank: global knowledge
    n = 4000
    pos = np.stack([rng.uniform(0.6, CX - 0.6, n), rng.uniform(0, CY, n), rng.uniform(0, CZ, n)], 1).astype(np.float32)
    ids = np.arange(n, dtype=np.float32)
    plane = pos[:, 0].astype(np.int64)
    cuts = slabs.choose_cuts(plane, CX, world)
    assert cuts[0] == 0 and cuts[-1] == CX and all(b > a for a, b in zip(cuts, cuts[1:]))
    x0, x1 = cuts[rank], cuts[rank + 1]
    w = x1 - x0
    ex = slabs.SlabExchange(rank, world, "cpu")
    # deliberately sloppy initial partition (like the host-side approximation of the device hash): shift the cut
    # by up to one plane for some particles
    fuzzy = plane + (rng.integers(0, 20, n) == 0) * rng.integers(-1, 2, n)
    fuzzy = np.clip(fuzzy, 0, CX - 1)
    fuzzy = np.where(np.abs(fuzzy - plane) <= 1, fuzzy, plane)
    mine = (fuzzy >= x0) & (fuzzy < x1)
    cap = n
    P, V, ID = torch.zeros((cap, 3)), torch.zeros((cap, 3)), torch.zeros(cap)
    alt = [torch.zeros((cap, 3)), torch.zeros((cap, 3)), torch.zeros(cap)]
    n_local = int(mine.sum())
    P[:n_local] = torch.from_numpy(pos[mine]); ID[:n_local] = torch.from_numpy(ids[mine])
    state = {}

    def keys_of(p):
        c = np.floor(p.numpy()).astype(np.int64)
        lx = c[:, 0] - (x0 - 1)
        ok = (lx >= 0) & (lx < w + 2) & (c[:, 1] >= 0) & (c[:, 1] < CY) & (c[:, 2] >= 0) & (c[:, 2] < CZ)
        k = (lx * CY + c[:, 1]) * CZ + c[:, 2]
        return np.where(ok, k, (w + 2) * CY * CZ)

    def search(m):                                        # stand-in for sphk_neighbor_search: stable sort by key
        k = keys_of(P[:m])
        o = torch.from_numpy(np.argsort(k, kind="stable"))
        for a in (P, V, ID):
            a[:m] = a[:m][o]
        state["keys"] = k[o.numpy()]

    def bounds():
        pc = CY * CZ
        return tuple(int(np.searchsorted(state["keys"], c * pc, side="left")) for c in (0, 1, 2, 3, max(w - 1, 0), w, w + 1, w + 2))

    # first step as in SlabSystem.begin_step: everything local is "own", ghost planes included in the candidates
    search(n_local)
    b0 = bounds()
    r = slabs.plane_ranges(b0, w)
    r["own"] = (b0[0], b0[7]); r["to_left"] = (b0[0], r["to_left"][1]); r["to_right"] = (r["to_right"][0], b0[7])
    ok = b0[7] == n_local
    for step in range(steps):
        if step > 0:
            # every rank moves the GLOBAL set identically (|dx| < 1 plane), and its own particles accordingly
            dx = rng.uniform(-0.45, 0.45, (n, 3)).astype(np.float32)
            dx[:, 1:] *= 0.2
            if bullets:                                       # ~2% of the particles cross up to four planes in this step
                fast = rng.integers(0, 50, n) == 0
                dx[fast, 0] = rng.uniform(-4.0, 4.0, int(fast.sum())).astype(np.float32)
            newpos = pos + dx
            newpos[:, 0] = np.clip(newpos[:, 0], 0.05, CX - 0.05)
            newpos[:, 1] = np.clip(newpos[:, 1], 0.01, CY - 0.01); newpos[:, 2] = np.clip(newpos[:, 2], 0.01, CZ - 0.01)
            pos = newpos
            a0, a1 = r["own"]
            my_ids = ID[a0:a1].numpy().astype(np.int64)
            P[a0:a1] = torch.from_numpy(pos[my_ids])          # "advect": same slots, new positions
        gathered = None
        if bullets and step > 0:
            # strays (>= 2 planes since the last sort) leave the regular flow: block -> all-gather -> appended below
            a0, a1 = r["own"]
            plane_sorted = state["keys"][a0:a1] // (CY * CZ)
            plane_now = np.floor(P[a0:a1, 0].numpy()).astype(np.int64) - (x0 - 1)
            block = slabs.collect_strays_host([P, V, ID], r["own"], plane_sorted, plane_now, STRAY_CAP)
            state["strays"] = state.get("strays", 0) + int(block[:1].view(torch.int32)[0])
            ok &= int(block[:1].view(torch.int32)[0]) <= STRAY_CAP
            gathered = ex.all_gather(block)
        n_all = slabs.exchange_candidates(ex, [P, V, ID], alt, r["own"], r["to_left"], r["to_right"])
        if gathered is not None:
            n_all += slabs.append_strays_host(gathered, world, STRAY_CAP, [P, V, ID], n_all)
        search(n_all)
        r = slabs.plane_ranges(bounds(), w)
        (o0, o1), (g0, g1), (h0, h1) = r["own"], r["ghost_l"], r["ghost_r"]
        k = keys_of(P[:h1])
        ok &= bool(np.all(np.diff(k) >= 0)) and g0 == 0 and g1 == o0 and o1 == h0
        loc_ids = ID[:h1].numpy().astype(np.int64)
        gplane = np.floor(pos[:, 0]).astype(np.int64)
        ok &= set(loc_ids[o0:o1].tolist()) == set(np.nonzero((gplane >= x0) & (gplane < x1))[0].tolist())
        ok &= len(set(loc_ids[o0:o1].tolist())) == o1 - o0            # no duplicates
        ok &= set(loc_ids[g0:g1].tolist()) == set(np.nonzero(gplane == x0 - 1)[0].tolist())
        ok &= set(loc_ids[h0:h1].tolist()) == set(np.nonzero(gplane == x1)[0].tolist())
        ok &= bool(np.array_equal(P[:h1].numpy(), pos[loc_ids]))
        # field halo: owners publish f = 2*id + step; ghosts must receive exactly that, in the ghost's SORTED order
        # (the ordering contract of exchange_candidates' docstring)
        f = torch.zeros(cap)
        f[o0:o1] = 2 * ID[o0:o1] + step
        ex.exchange(f[r["first"][0]:r["first"][1]].contiguous(), f[r["last"][0]:r["last"][1]].contiguous(), f[g0:g1], f[h0:h1])
        ok &= bool(torch.equal(f[:h1], 2 * ID[:h1] + step))
    allok = [None] * world
    dist.all_gather_object(allok, (bool(ok), state.get("strays", 0)))
    if rank == 0:
        q.put(all(o for o, _ in allok) and (not bullets or sum(c for _, c in allok) > 20))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,bullets", [(2, False), (3, False), (2, True), (3, True)])
def test_slab_assembly_and_halo_gloo(world, bullets):
    """bullets: a few particles per step cross up to four planes -- more than the candidate exchange covers -- and must
    reach their new owner (and its neighbours' ghost planes, in the same order) through the stray routing."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5 if bullets else 4, q, bullets)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_choose_cuts_balances():
    import pkgload
    pkgload.load()
    from cpp_fluid_particles_b200 import slabs
    rng = np.random.default_rng(0)
    plane = rng.integers(18, 83, 200000)                  # the 2M scene occupies planes 18..82 of 100
    for world in (2, 4, 8):
        cuts = slabs.choose_cuts(plane, 100, world)
        counts = [int(((plane >= a) & (plane < b)).sum()) for a, b in zip(cuts, cuts[1:])]
        assert sum(counts) == plane.size and max(counts) <= 1.25 * plane.size / world


def test_plane_ranges_thin_slabs():
    """plane_ranges for slabs of one, two and many planes: owned range, the (at most two) candidate planes per side,
    first / last plane and ghost ranges stay consistent and never reach outside the owned range."""
    import pkgload
    pkgload.load()
    from cpp_fluid_particles_b200.slabs import plane_ranges
    rng = np.random.default_rng(7)
    for w in (1, 2, 3, 9):
        counts = rng.integers(0, 50, w + 2)                       # particles per local plane 0..w+1
        off = np.concatenate([[0], np.cumsum(counts)])            # plane start offsets, off[w+2] = total
        b = (off[0], off[1], off[min(2, w + 2)], off[min(3, w + 2)], off[max(w - 1, 0)], off[w], off[w + 1], off[w + 2])
        r = plane_ranges(tuple(int(x) for x in b), w)
        own = (int(off[1]), int(off[w + 1]))
        assert r["own"] == own and r["ghost_l"] == (0, int(off[1])) and r["ghost_r"] == (int(off[w + 1]), int(off[w + 2]))
        assert r["first"] == (int(off[1]), int(off[2])) and r["last"] == (int(off[w]), int(off[w + 1]))
        for key in ("to_left", "to_right", "first", "last"):
            lo, hi = r[key]
            assert own[0] <= lo <= hi <= own[1], (w, key, r[key], own)
        # candidates = the two outermost owned planes per side (all owned planes when the slab is thinner)
        assert r["to_left"] == (own[0], int(off[min(3, w + 1)]))
        assert r["to_right"] == (int(off[max(w - 1, 1)]), own[1])


def test_choose_cuts_degenerate():
    import pkgload
    pkgload.load()
    from cpp_fluid_particles_b200 import slabs
    # every particle in one plane, more ranks than occupied planes: still strictly increasing cuts covering the grid
    plane = np.full(1000, 5)
    for world in (2, 3, 8):
        cuts = slabs.choose_cuts(plane, 12, world)
        assert cuts[0] == 0 and cuts[-1] == 12 and len(cuts) == world + 1
        assert all(b > a for a, b in zip(cuts, cuts[1:]))


# ---- the NATIVE begin_step (SlabSystem._begin_step_native over csrc/sphk_mg.cu) with the C entry points replaced by
# ---- gloo / numpy stand-ins: same host logic as on the GPUs (assembly straight into the scratch twins, swap, counts
# ---- agreed one step ahead, ordering-contract check, halo ranges), no GPU, no library calls -------------------------
def _native_worker(rank, world, port, steps, q, use_async=False):
    import ctypes as C
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pkgload
    pkgload.load()
    from cpp_fluid_particles_b200 import slabs
    rng = np.random.default_rng(321)
    n = 5000
    pos = np.stack([rng.uniform(0.6, CX - 0.6, n), rng.uniform(0, CY, n), rng.uniform(0, CZ, n)], 1).astype(np.float32)
    plane = pos[:, 0].astype(np.int64)
    cuts = slabs.choose_cuts(plane, CX, world)
    x0, x1 = cuts[rank], cuts[rank + 1]
    w = x1 - x0
    cap = n
    left = rank - 1 if rank > 0 else None
    right = rank + 1 if rank < world - 1 else None

    def fview(ptr, count):          # float32 view of raw memory (what the C side would see)
        return torch.from_numpy(np.ctypeslib.as_array((C.c_float * max(count, 1)).from_address(ptr))[:count])

    def val(x):
        return x.value if hasattr(x, "value") else x

    def p2p(sends, recvs):
        ops = [dist.P2POp(dist.isend, t, p) for t, p in sends] + [dist.P2POp(dist.irecv, t, p) for t, p in recvs]
        if ops:
            for wk in dist.batch_isend_irecv(ops):
                wk.wait()

    class FakeLib:                  # the sphk_* / sphk_mg_* calls _begin_step_native and _halo make
        def sphk_mg_exchange_ints(self, mg, tl, tr, fl, fr, k):
            k = val(k)
            sends, recvs, bufs = [], [], {}
            for side, peer, out in ((0, left, tl), (1, right, tr)):
                if peer is not None:
                    sends.append((torch.tensor(list(out)[:k], dtype=torch.int32), peer))
                    bufs[side] = torch.zeros(k, dtype=torch.int32)
                    recvs.append((bufs[side], peer))
            p2p(sends, recvs)
            for side, arr in ((0, fl), (1, fr)):
                for i in range(k):
                    arr[i] = int(bufs[side][i]) if side in bufs else 0
            return 0

        def sphk_mg_exchange_slices(self, mg, k, src, dst, widths, sl, sr, rl, rr):
            sends, recvs = [], []
            for a in range(val(k)):
                wd = widths[a]
                for peer, (b, c), arr, lst in ((left, sl, src, sends), (right, sr, src, sends), (left, rl, dst, recvs), (right, rr, dst, recvs)):
                    if peer is not None and c > 0:
                        lst.append((fview(arr[a] + 4 * wd * b, wd * c), peer))
            p2p(sends, recvs)
            return 0

        def sphk_mg_halo(self, mg, ctx, scene, what, ptr, width, r):
            base, wd = val(ptr), val(width)
            sends, recvs = [], []
            for peer, sb, sc_, gb, gc in ((left, r[0], r[1], r[4], r[5]), (right, r[2], r[3], r[6], r[7])):
                if peer is not None:
                    if sc_ > 0:
                        sends.append((fview(base + 4 * wd * sb, wd * sc_).clone(), peer))
                    if gc > 0:
                        recvs.append((fview(base + 4 * wd * gb, wd * gc), peer))
            p2p(sends, recvs)
            return 0

        def sphk_copy(self, ctx, dst, src, nf):
            fview(val(dst), val(nf)).copy_(fview(val(src), val(nf)).clone())
            return 0

        def sphk_mg_check(self, mg, err):
            err._obj.value = 0
            return 0

        def sphk_set_active_range(self, ctx, b, c):
            return 0

        # ---- the host-free variants (device-resident ranges: here plain memory behind the same addresses) ----
        def sphk_mg_plane_ranges(self, mg, cs_ptr, plane_cells, w_, dev_ptr, pin_ptr):
            pc, ww = val(plane_cells), val(w_)
            cs = np.ctypeslib.as_array((C.c_int * ((ww + 2) * pc + 1)).from_address(val(cs_ptr)))
            s0, s1, s2, s3 = (int(cs[k * pc]) for k in range(4))
            swm1, sw, sw1, send = int(cs[max(ww - 1, 0) * pc]), int(cs[ww * pc]), int(cs[(ww + 1) * pc]), int(cs[(ww + 2) * pc])
            fb, fe = s1, (s2 if ww >= 2 else sw1)
            lb, le = (sw if ww >= 2 else s1), sw1
            tlb, tle = s1, (min(s3, sw1) if ww >= 2 else sw1)
            trb, tre = (max(swm1, s1) if ww >= 2 else s1), sw1
            out = [s0, s1, s2, s3, swm1, sw, sw1, send, fb, fe - fb, lb, le - lb, s0, s1 - s0, sw1, send - sw1, s1, sw1 - s1,
                   tlb, tle - tlb, trb, tre - trb, tle - tlb, tre - trb]
            for ptr in (val(dev_ptr), val(pin_ptr)):
                np.ctypeslib.as_array((C.c_int * 24).from_address(ptr))[:] = out
            return 0

        def sphk_mg_exchange_ints_async(self, mg, dl, dr, k, pfl, pfr):
            k = val(k)
            sends, recvs, bufs = [], [], {}
            for side, peer, ptr in ((0, left, dl), (1, right, dr)):
                if peer is not None:
                    sends.append((torch.tensor(list((C.c_int * k).from_address(val(ptr))), dtype=torch.int32), peer))
                    bufs[side] = torch.zeros(k, dtype=torch.int32)
                    recvs.append((bufs[side], peer))
            p2p(sends, recvs)
            for side, ptr in ((0, pfl), (1, pfr)):
                arr = (C.c_int * k).from_address(val(ptr))
                for i in range(k):
                    arr[i] = int(bufs[side][i]) if side in bufs else 0
            return 0

        def sphk_mg_check_async(self, mg, ptr):
            (C.c_int * 1).from_address(val(ptr))[0] = 0
            return 0

        def sphk_set_active_range_device(self, ctx, ptr):
            return 0

        def sphk_mg_halo_device(self, mg, ctx, scene, what, ptr, width, ranges_ptr, maxp):
            r = list((C.c_int * 8).from_address(val(ranges_ptr)))
            return self.sphk_mg_halo(mg, ctx, scene, what, ptr, width, r)

        # ---- strays: the two kernels as their torch statements, the all-gather over gloo ----
        def _views(self, k, ptrs, widths):
            return [fview(ptrs[a], cap * widths[a]).reshape(cap, widths[a]) if widths[a] > 1 else fview(ptrs[a], cap) for a in range(val(k))]

        def sphk_strays_collect(self, ctx, cs_ptr, a, cnt, k, ptrs, widths, block_ptr, capacity):
            a, cnt, capacity, pc = val(a), val(cnt), val(capacity), CY * CZ
            arrays = self._views(k, ptrs, widths)
            cs = np.ctypeslib.as_array((C.c_int * ((w + 2) * pc + 1)).from_address(val(cs_ptr)))
            offs = cs[::pc][:w + 3]
            plane_sorted = np.searchsorted(offs, np.arange(a, a + cnt), side="right") - 1
            plane_now = np.floor(arrays[0][a:a + cnt, 0].numpy()).astype(np.int64) - (x0 - 1)
            block = slabs.collect_strays_host(arrays, (a, a + cnt), plane_sorted, plane_now, capacity)
            fview(val(block_ptr), block.numel()).copy_(block)
            return 0

        def sphk_mg_strays_route(self, mg, ctx, block_ptr, gathered_ptr, capacity, k, ptrs, widths, dst_begin):
            capacity = val(capacity)
            nf = slabs.stray_block_floats(capacity, [widths[a] for a in range(val(k))])
            parts = [torch.zeros(nf) for _ in range(world)]
            dist.all_gather(parts, fview(val(block_ptr), nf).clone())
            gathered = fview(val(gathered_ptr), world * nf)
            gathered.copy_(torch.cat(parts))
            slabs.append_strays_host(gathered, world, capacity, self._views(k, ptrs, widths), val(dst_begin))
            return 0

    class Fluid:
        pass

    class FakeSlab:
        _begin_step_native = slabs.SlabSystem._begin_step_native
        _begin_step_async = slabs.SlabSystem._begin_step_async
        _refresh_ranges = slabs.SlabSystem._refresh_ranges
        begin_step = slabs.SlabSystem.begin_step
        _swap_carried = slabs.SlabSystem._swap_carried
        _exchange_ints = slabs.SlabSystem._exchange_ints
        _carried = slabs.SlabSystem._carried
        _bounds = slabs.SlabSystem._bounds
        _search_all = slabs.SlabSystem._search_all
        _halo = slabs.SlabSystem._halo
        _route_strays = slabs.SlabSystem._route_strays
        _collect_strays = slabs.SlabSystem._collect_strays
        _array_args = slabs.SlabSystem._array_args
        stray_cap, _strays_pending = 64, False

        def _s(self):
            return None

        def search_fluid(self):         # stand-in for sphk_neighbor_search: stable sort of pos / vel by local cell key
            m = self.fluid.n
            c = np.floor(self.fluid.pos[:m].numpy()).astype(np.int64)
            lx = c[:, 0] - (x0 - 1)
            ok_ = (lx >= 0) & (lx < w + 2) & (c[:, 1] >= 0) & (c[:, 1] < CY) & (c[:, 2] >= 0) & (c[:, 2] < CZ)
            ncl = (w + 2) * CY * CZ
            k = np.where(ok_, (lx * CY + c[:, 1]) * CZ + c[:, 2], ncl)
            o = torch.from_numpy(np.argsort(k, kind="stable"))
            self.fluid.pos[:m] = self.fluid.pos[:m][o]
            self.fluid.vel[:m] = self.fluid.vel[:m][o]
            self.perm = o                                               # the solver permutes its history array itself
            self.cs_fluid[:] = torch.from_numpy(np.searchsorted(k[o.numpy()], np.arange(ncl + 1), side="left").astype(np.int32))

    s = FakeSlab()
    s.L, s.mg, s.ctx = FakeLib(), object(), None
    s.rank, s.world, s.device, s.cap, s.w, s.plane_cells, s.solver = rank, world, torch.device("cpu"), cap, w, CY * CZ, "dfsph"
    s.ex = type("Ex", (), {"left": left, "right": right})()
    s.fluid = Fluid()
    s.fluid.pos, s.fluid.vel, s.warm = torch.zeros((cap, 3)), torch.zeros((cap, 3)), torch.zeros(cap)
    s.cs_fluid = torch.zeros((w + 2) * CY * CZ + 1, dtype=torch.int32)
    s._ranges, s.use_list, s.comm_s, s.time_assembly, s._assembly_events, s._scene = None, False, 0.0, False, [], None
    s.async_assembly, s.transport, s._async_pending, s._step_async = use_async, 1, False, False
    s._dev24, s._pin24, s._pin_misc = torch.zeros(24, dtype=torch.int32), torch.zeros(24, dtype=torch.int32), torch.zeros(8, dtype=torch.int32)
    s._async_event = type("Ev", (), {"record": lambda self: None, "synchronize": lambda self: None})()
    used_async = strays = 0
    mine = (plane >= x0) & (plane < x1)
    s.n_own = int(mine.sum())
    s.fluid.pos[:s.n_own] = torch.from_numpy(pos[mine])
    s.fluid.vel[:s.n_own, 0] = torch.from_numpy(np.nonzero(mine)[0].astype(np.float32))     # vel.x carries the particle id
    ok = True
    for step in range(steps):
        if step > 0:                      # every rank moves the GLOBAL set identically (< 1 plane), its own particles accordingly
            dx = rng.uniform(-0.45, 0.45, (n, 3)).astype(np.float32)
            dx[:, 1:] *= 0.2
            fast = rng.integers(0, 100, n) == 0                       # ... except 1 %: up to four planes (strays)
            dx[fast, 0] = rng.uniform(-4.0, 4.0, int(fast.sum())).astype(np.float32)
            pos = pos + dx
            pos[:, 0] = np.clip(pos[:, 0], 0.05, CX - 0.05)
            pos[:, 1] = np.clip(pos[:, 1], 0.01, CY - 0.01); pos[:, 2] = np.clip(pos[:, 2], 0.01, CZ - 0.01)
            s._refresh_ranges()
            a0, a1 = s._ranges["own"]
            ids = s.fluid.vel[a0:a1, 0].numpy().astype(np.int64)
            s.fluid.pos[a0:a1] = torch.from_numpy(pos[ids])
            s._collect_strays()               # what SlabSystem.step does before begin_step
            strays += int(s._stray_block[:1].view(torch.int32)[0])
        s.begin_step()
        used_async += int(s._step_async)
        if step % 2 == 0:
            s._refresh_ranges()             # (also exercise a mid-step refresh: the halos below must keep using the device ranges)
        s.warm[:s.fluid.n] = s.warm[:s.fluid.n][s.perm]                  # what step_dfsph's sphk_permute does
        if step % 2 == 1:
            s._refresh_ranges()
        r = s._ranges
        (o0, o1), (g0, g1), (h0, h1) = r["own"], r["ghost_l"], r["ghost_r"]
        ids = s.fluid.vel[:h1, 0].numpy().astype(np.int64)
        gplane = np.floor(pos[:, 0]).astype(np.int64)
        ok &= g0 == 0 and g1 == o0 and o1 == h0
        ok &= set(ids[o0:o1].tolist()) == set(np.nonzero((gplane >= x0) & (gplane < x1))[0].tolist()) and len(set(ids[o0:o1].tolist())) == o1 - o0
        ok &= set(ids[g0:g1].tolist()) == set(np.nonzero(gplane == x0 - 1)[0].tolist())
        ok &= set(ids[h0:h1].tolist()) == set(np.nonzero(gplane == x1)[0].tolist())
        ok &= bool(np.array_equal(s.fluid.pos[:h1].numpy(), pos[ids]))
        # the history array travelled with its particle (owned range): warm = id of the previous step's owner write
        if step > 0:
            ok &= bool(np.array_equal(s.warm[o0:o1].numpy(), (3.0 * ids[o0:o1] + (step - 1)).astype(np.float32)))
        s.warm[o0:o1] = torch.from_numpy((3.0 * ids[o0:o1] + step).astype(np.float32))
        # one halo of a scalar and one of a float3 array: ghosts must receive their owners' values, in sorted order
        f = torch.zeros(cap)
        f[o0:o1] = torch.from_numpy((2.0 * ids[o0:o1] + step).astype(np.float32))
        s._halo(2, f)
        ok &= bool(np.array_equal(f[:h1].numpy(), (2.0 * ids + step).astype(np.float32)))
        g3 = torch.zeros((cap, 3))
        g3[o0:o1] = s.fluid.pos[o0:o1] * 2.0
        s._halo(0, g3)
        ok &= bool(np.array_equal(g3[:h1].numpy(), pos[ids] * np.float32(2.0)))
    ok &= used_async == (steps - 1 if use_async else 0)     # the first step sizes itself synchronously, every later one is host-free
    allok = [None] * world
    dist.all_gather_object(allok, (bool(ok), strays))
    if rank == 0:
        q.put(all(o for o, _ in allok) and sum(c for _, c in allok) > 20)      # (and the stray routing was exercised)
    dist.destroy_process_group()


@pytest.mark.parametrize("use_async", [False, True])
@pytest.mark.parametrize("world", [2, 3, 4])
def test_native_begin_step_host_logic_gloo(world, use_async):
    """use_async: the host-free assembly (_begin_step_async: last step's ranges from pinned memory, this step's ranges
    device-resident) against the same invariants as the synchronous one."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_native_worker, args=(r, world, port, 6, q, use_async)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
