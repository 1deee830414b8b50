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


def _worker(rank, world, port, steps, q):
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
            newpos = pos + dx
            newpos[:, 0] = np.clip(newpos[:, 0], 0.05, CX - 0.05)
            newpos[:, 1] = np.clip(newpos[:, 1], 0.01, CY - 0.01); newpos[:, 2] = np.clip(newpos[:, 2], 0.01, CZ - 0.01)
            pos = newpos
            a0, a1 = r["own"]
            my_ids = ID[a0:a1].numpy().astype(np.int64)
            P[a0:a1] = torch.from_numpy(pos[my_ids])          # "advect": same slots, new positions
        n_all = slabs.exchange_candidates(ex, [P, V, ID], alt, r["own"], r["to_left"], r["to_right"])
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
    dist.all_gather_object(allok, bool(ok))
    if rank == 0:
        q.put(all(allok))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_assembly_and_halo_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 4, q)) for r in range(world)]
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
