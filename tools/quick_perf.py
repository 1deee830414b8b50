"""Quick A/B timing probe (not the bench): ms/step of ours vs the reference CUDA build on one scene."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import capi
name = sys.argv[1] if len(sys.argv) > 1 else "2m"
solvers = sys.argv[2].split(",") if len(sys.argv) > 2 else ["dfsph"]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libsphref.so")
for solver in solvers:
    sc = pkg.scene.benchmark_scene(name, solver)
    for label, lib in (("ours", capi.LIBHOST), ("reference-cuda", LIBREF)):
        if not os.path.exists(lib):
            continue
        t0 = time.time()
        app = capi.SphApp(sc, lib)
        t1 = time.time()
        for _ in range(3):
            app.step()
        ms = [app.step() for _ in range(steps)]
        st = app.download()
        print(f"{name} {solver} {label}: ctor {t1-t0:.2f}s  median {np.median(ms):.3f} ms/step  min {np.min(ms):.3f}  "
              f"-> {sc.fluid.shape[0]/np.median(ms)*1e3/1e6:.1f} M particle-steps/s; dens max {st['density'].max():.4f}", flush=True)
        app.close()
