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
    mine = (plane >= x0) & (plane < x1)
    cap = n
    P, V, ID = torch.zeros((cap, 3)), torch.zeros((cap, 3)), torch.zeros(cap)
    alt = [torch.zeros((cap, 3)), torch.zeros((cap, 3)), torch.zeros(cap)]
    n_own, gl = int(mine.sum()), 0
    P[:n_own] = torch.from_numpy(pos[mine]); ID[:n_own] = torch.from_numpy(ids[mine])
    state = {}

    def keys_of(p):
        c = np.floor(p.numpy()).astype(np.int64)
        lx = c[:, 0] - (x0 - 1)
        ok = (lx >= 0) & (lx < w + 2) & (c[:, 1] >= 0) & (c[:, 1] < CY) & (c[:, 2] >= 0) & (c[:, 2] < CZ)
        k = (lx * CY + c[:, 1]) * CZ + c[:, 2]
        return np.where(ok, k, (w + 2) * CY * CZ)

    def search(off, m):                                   # stand-in for sphk_neighbor_search: stable sort by key
        k = keys_of(P[off:off + m])
        o = torch.from_numpy(np.argsort(k, kind="stable"))
        for a in (P, V, ID):
            a[off:off + m] = a[off:off + m][o]
        state["keys"] = k[o.numpy()]

    def bounds():
        pc = CY * CZ
        return tuple(int(np.searchsorted(state["keys"], c * pc, side="left")) for c in (0, 1, 2, w, w + 1, w + 2))

    ok = True
    for step in range(steps):
        # every rank moves the GLOBAL set identically (|dx| < 1 plane), and its own particles accordingly
        dx = rng.uniform(-0.45, 0.45, (n, 3)).astype(np.float32)
        dx[:, 1:] *= 0.2
        newpos = pos + dx
        newpos[:, 0] = np.clip(newpos[:, 0], 0.05, CX - 0.05)
        newpos[:, 1] = np.clip(newpos[:, 1], 0.01, CY - 0.01); newpos[:, 2] = np.clip(newpos[:, 2], 0.01, CZ - 0.01)
        pos = newpos
        my_ids = ID[gl:gl + n_own].numpy().astype(np.int64)
        P[gl:gl + n_own] = torch.from_numpy(pos[my_ids])
        gl, n_own, gr = slabs.assemble_slab(ex, [P, V, ID], alt, gl, n_own, search, bounds)
        total = gl + n_own + gr
        search(0, total)                                  # step E: the final sort of the assembled set
        k = keys_of(P[:total])
        ok &= bool(np.all(np.diff(k) >= 0))
        loc_ids = ID[:total].numpy().astype(np.int64)
        gplane = np.floor(pos[:, 0]).astype(np.int64)
        ok &= set(loc_ids[gl:gl + n_own].tolist()) == set(np.nonzero((gplane >= x0) & (gplane < x1))[0].tolist())
        ok &= set(loc_ids[:gl].tolist()) == set(np.nonzero(gplane == x0 - 1)[0].tolist())
        ok &= set(loc_ids[gl + n_own:total].tolist()) == set(np.nonzero(gplane == x1)[0].tolist())
        ok &= bool(np.array_equal(P[:total].numpy(), pos[loc_ids]))
        # field halo: owners publish f = 2*id + step; ghosts must receive exactly that, in the ghost's SORTED order
        # (this is the ordering contract of assemble_slab's docstring)
        f = torch.zeros(cap)
        f[gl:gl + n_own] = 2 * ID[gl:gl + n_own] + step
        s0, s1, s2, sw, sw1, send = bounds()
        ok &= (s1 == gl) and (sw1 == gl + n_own)
        ex.exchange(f[s1:s2].contiguous(), f[sw:sw1].contiguous(), f[0:gl], f[gl + n_own:total])
        ok &= bool(torch.equal(f[:total], 2 * ID[:total] + step))
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
