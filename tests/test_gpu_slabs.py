"""Multi-rank slab driver against the single-GPU result (<= 1e-5 on pos / density, north_star).
On a 1-GPU box the ranks share cuda:0 (gloo, host-staged halo); with >= 2 GPUs the NCCL path runs too."""
import json
import os
import subprocess
import sys

import pytest

from util import ROOT

pytestmark = pytest.mark.gpu


def _run(nproc, extra, timeout=600, env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + nproc + len(extra)),
           os.path.join(ROOT, "tools", "slab_check.py")] + extra
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=timeout, cwd=ROOT,
                       env=dict(os.environ, **(env or {})))
    line = [l for l in r.stdout.splitlines() if l.startswith("SLAB_CHECK ")]
    assert r.returncode == 0 and line, r.stdout[-3000:]
    return json.loads(line[-1][len("SLAB_CHECK "):])


@pytest.mark.parametrize("solver", ["dfsph", "wcsph", "pbd"])
@pytest.mark.parametrize("world", [2, 3])
def test_slabs_match_single_gpu_shared_device(built, solver, world):
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    out = _run(world, ["--backend", "gloo", "--same-gpu", "--solver", solver, "--steps", "3", "--jitter", "0.001"])
    assert out["ok"], out
    assert sum(c[1] for c in out["counts"]) == out["steps"][0]["n"]


@pytest.mark.parametrize("solver,world", [("dfsph", 2), ("pbd", 3), ("wcsph", 3)])
def test_slabs_rebalance_cuts(built, solver, world):
    """SURVEY 8e "re-balance every K steps": the run starts with its interior cuts 3 planes off balance and re-balances
    every 2 steps (each event moves a cut by one plane; the ranks' grids, boundary slices and candidate planes follow).
    The result must stay the single-GPU result, the cuts must come back and the load imbalance must fall below 1.1."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    out = _run(world, ["--backend", "gloo", "--same-gpu", "--solver", solver, "--steps", "12", "--jitter", "0.001", "--skew", "3",
                       "--rebalance", "2"])
    assert out["ok"], out
    assert out["rebalanced"] >= 3 and out["cuts_final"] != out["cuts_initial"], out
    assert out["imbalance_at_last_rebalance"] < 1.1, out
    assert sum(c[1] for c in out["counts"]) == out["steps"][0]["n"]


@pytest.mark.parametrize("solver,world", [("dfsph", 3), ("pbd", 2), ("wcsph", 2)])
def test_slabs_route_fast_particles(built, solver, world):
    """Particles that cross several cell planes in ONE step (here: top-layer particles shot along +-x at ~3 planes per step;
    in the benchmark scene: what the impact on the floor does around step 40) are outside the one-plane contract of the
    candidate exchange.  They must be routed to their new owners and ghost planes (the "strays" of include/sphk.h) and the
    result must stay the single-GPU result."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    out = _run(world, ["--backend", "gloo", "--same-gpu", "--solver", solver, "--steps", "5", "--bullets", "6"])
    assert out["ok"], out
    assert out["strays_routed"] >= 6, out
    assert sum(c[1] for c in out["counts"]) == out["steps"][0]["n"]


@pytest.mark.parametrize("solver", ["dfsph", "pbd"])
def test_slabs_device_side_scene(built, solver):
    """every rank generates its own lattice columns and the boundary shell on the device (no host-side
    particle array, boundary planes cut out of the sorted shell as one slice) -- against the single-GPU run built from
    the host arrays."""
    import torch
    if not torch.cuda.is_available():
        pytest.fail("needs a CUDA device")
    out = _run(2, ["--backend", "gloo", "--same-gpu", "--solver", solver, "--steps", "3", "--device-scene"])
    assert out["ok"], out
    assert sum(c[1] for c in out["counts"]) == out["steps"][0]["n"]


@pytest.mark.parametrize("transport,solver", [("mailbox", "dfsph"), ("mailbox", "wcsph"), ("mailbox", "pbd"),
                                              ("nccl", "dfsph"), ("nccl", "wcsph"), ("torch", "dfsph")])
def test_slabs_match_single_gpu_nccl(built, solver, transport):
    """One rank per GPU over NVLink: the native exchanges of csrc/sphk_mg.cu (peer-memory mailbox halos / NCCL halos)
    and the torch.distributed path must all reproduce the single-GPU result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    env = {"mailbox": {"SPHK_SLAB_NATIVE": "1", "SPHK_SLAB_TRANSPORT": "1"},
           "nccl": {"SPHK_SLAB_NATIVE": "1", "SPHK_SLAB_TRANSPORT": "0"},
           "torch": {"SPHK_SLAB_NATIVE": "0"}}[transport]
    out = _run(2, ["--backend", "nccl", "--solver", solver, "--steps", "3", "--jitter", "0.001"], timeout=100, env=env)
    assert out["ok"], out
    assert sum(c[1] for c in out["counts"]) == out["steps"][0]["n"]


@pytest.mark.parametrize("solver", ["dfsph", "pbd"])
def test_slabs_route_fast_particles_native(built, solver):
    """test_slabs_route_fast_particles over the native transports on 2 GPUs (NCCL all-gather of the stray blocks, host-free
    step assembly)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    out = _run(2, ["--backend", "nccl", "--solver", solver, "--steps", "5", "--bullets", "6"], timeout=100)
    assert out["ok"], out
    assert out["strays_routed"] >= 6, out


@pytest.mark.parametrize("solver,bullets", [("dfsph", 0), ("sph", 0), ("pbd", 0), ("dfsph", 6), ("pbd", 6)])
def test_cpp_slab_system_matches_single_gpu(built, solver, bullets, tmp_path):
    """The class-API system sharded over 2 GPUs (host/sph_slab.hpp: SlabSPHSystem, one process per GPU, file rendezvous,
    native NCCL + peer-memory mailbox halos) against SPHSystem on one GPU, both through the reference's call sites
    (app/sph_headless.cpp): the ranks' owned particles together must be the single-GPU particles, <= 1e-5 on position
    and density (north_star).  Particles are matched by position (the ranks keep their own order)."""
    import numpy as np
    import torch
    from scipy.spatial import cKDTree
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    cli = os.path.join(ROOT, "cpp-fluid-particles_b200", "sph_headless")
    common = ["--solver", solver, "--frames", "6", "--quiet", "--block", "30", "40", "30", "--bullets", str(bullets)]
    one = subprocess.run([cli] + common + ["--dump", str(tmp_path / "one")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=120)
    assert one.returncode == 0, one.stdout[-2000:]
    two = subprocess.run([cli] + common + ["--ranks", "2", "--dump", str(tmp_path / "two")], stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True, timeout=180)
    assert two.returncode == 0, two.stdout[-2000:]
    lines = [json.loads(l) for l in two.stdout.splitlines() if l.startswith("{")]
    assert sorted(l["rank"] for l in lines) == [0, 1], two.stdout[-2000:]
    pos1 = np.fromfile(tmp_path / "one.pos.f32", dtype=np.float32).reshape(-1, 3)
    den1 = np.fromfile(tmp_path / "one.density.f32", dtype=np.float32)
    pos2 = np.concatenate([np.fromfile(tmp_path / f"two.rank{r}.pos.f32", dtype=np.float32).reshape(-1, 3) for r in range(2)])
    den2 = np.concatenate([np.fromfile(tmp_path / f"two.rank{r}.density.f32", dtype=np.float32) for r in range(2)])
    assert pos2.shape == pos1.shape and sum(l["n_owned"] for l in lines) == pos1.shape[0], (pos1.shape, pos2.shape, lines)
    assert min(l["n_owned"] for l in lines) > 0.3 * pos1.shape[0], lines
    assert bullets == 0 or lines[0]["strays_routed"] >= bullets, lines     # (with --bullets: fast particles, routed as strays)
    d, idx = cKDTree(pos2).query(pos1)
    assert np.unique(idx).size == idx.size, "the match between the two runs is not one-to-one"
    scale = max(1.0, float(np.abs(pos1).max()))
    assert float(d.max()) <= 1e-5 * scale, float(d.max())
    rel = np.abs(den2[idx] - den1) / np.maximum(np.abs(den1), 1e-12)
    assert float(rel.max()) <= 1e-5, float(rel.max())
