"""Does any particle cross more than one cell plane in x in one step (the slab decomposition's contract, slabs.py)?
Single GPU: per step max |vx| dt / cell_length over all particles, how many exceed 1, and where they are.

    python tools/plane_skip_probe.py 4m dfsph 70
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import engine
name, solver, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
sc = pkg.scene.benchmark_scene(name, solver)
s = engine.SphkSystem(sc)
cell, dt = float(sc.params.cell_length), float(sc.params.dt)
for k in range(steps):
    s.step()
    st = s.state()
    d = np.abs(st["vel"][:, 0]) * dt / cell
    fast = d > 1.0
    speed = np.linalg.norm(st["vel"], axis=1)
    line = f"step {k:3d}: max |vx| dt/cell {d.max():6.3f}  > 1: {int(fast.sum()):6d}  > 0.5: {int((d > 0.5).sum()):7d}  |v| max {speed.max():8.3f}  density max {st['density'].max():8.3f}"
    if fast.any():
        x = st["pos"][fast, 0]
        line += f"  x of the fast ones: {x.min():.3f} .. {x.max():.3f}"
    print(line, flush=True)
s.close()
