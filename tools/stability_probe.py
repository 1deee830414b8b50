"""Density maximum / neighbour statistics over a long run, ours vs the reference CUDA build (same scene): is a blow-up ours?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import capi
name, solver, steps, which = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
every = int(sys.argv[5]) if len(sys.argv) > 5 else 20
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libsphref.so")
sc = pkg.scene.benchmark_scene(name, solver)
app = capi.SphApp(sc, capi.LIBHOST if which == "ours" else LIBREF)
for k in range(steps):
    ms = app.step()
    if k % every == every - 1:
        st = app.download()
        v = np.linalg.norm(st["vel"], axis=1)
        print(f"{which} step {k:4d}: {ms:8.3f} ms  density max {st['density'].max():10.3f} mean {st['density'].mean():.4f}  |v| max {v.max():8.3f}  "
              f"pos min {st['pos'].min():.4f} max {st['pos'].max():.4f}", flush=True)
        if not np.isfinite(st["pos"]).all() or ms > 500:
            break
app.close()
