"""ms/step of this engine and of the reference CUDA build over a LONG dam-break (not only the first 25 steps of the bench):
medians over windows of steps, same scene, same box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import capi
name, solver, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libsphref.so")
sc = pkg.scene.benchmark_scene(name, solver)
res = {}
for label, lib in (("ours", capi.LIBHOST), ("reference", LIBREF)):
    if not os.path.exists(lib):
        continue
    if len(sys.argv) > 4 and label not in sys.argv[4].split(","):
        continue
    app = capi.SphApp(sc, lib)
    ms_all = []
    for k in range(steps):
        ms_all.append(app.step())
        if k % 20 == 19:
            print(f"  {label} steps {k-19}-{k}: median {np.median(ms_all[-20:]):.3f} ms, max {np.max(ms_all[-20:]):.3f} ms", flush=True)
    res[label] = np.array(ms_all)
    st = app.download()
    print(label, "final density max %.3f" % st["density"].max(), flush=True)
    app.close()
w = max(steps // 10, 1)
print(f"{name} {solver}: median ms/step per window of {w} steps")
for k in range(0, steps, w):
    row = "  steps %4d-%4d:" % (k, k + w - 1)
    for label, ms in res.items():
        row += "  %s %.3f" % (label, np.median(ms[k:k + w]))
    if len(res) == 2:
        row += "  ratio %.2f" % (np.median(res["reference"][k:k + w]) / np.median(res["ours"][k:k + w]))
    print(row)
