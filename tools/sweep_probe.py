"""A/B probe of engine variants on one scene (not the bench): per-call CUDA-event times of every C-ABI entry of a
step, for each variant given as `name:opt=value,opt=value` (options are SPHK_OPT_* numbers).

    python tools/sweep_probe.py 2m dfsph 6 base: sched:7=1
"""
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import pkgload

pkg = pkgload.load()
from cpp_fluid_particles_b200 import engine

name, solver, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
variants = []
for v in sys.argv[4:]:
    label, _, opts = v.partition(":")
    variants.append((label, [tuple(int(x) for x in o.split("=")) for o in opts.split(",") if o]))
sc = pkg.scene.benchmark_scene(name, solver)
n = sc.fluid.shape[0]
s = engine.SphkSystem(sc)
METHODS = ["search_fluid", "build_neighbor_list", "gravity", "viscosity", "color_grad", "surface", "density", "fused_density_color_grad",
           "fused_viscosity_surface", "fused_density_alpha_div_error", "fused_pbd_xsph_color_grad", "pressure", "pressure_force", "advect", "dfsph_density_alpha", "dfsph_div_error",
           "dfsph_div_correct", "dfsph_den_error", "dfsph_den_correct", "permute", "copy", "pbd_density_lambda",
           "pbd_delta_pos_apply", "pbd_velocity_from_positions", "pbd_xsph"]
events = defaultdict(list)
timing = [False]


def wrap(meth):
    inner = getattr(s, meth)

    def timed(*a, **k):
        if not timing[0]:
            return inner(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = inner(*a, **k); e1.record()
        events[meth].append((e0, e1))
        return r
    setattr(s, meth, timed)


for m in METHODS:
    if hasattr(s, m):
        wrap(m)
for _ in range(3):
    s.step()
ref_state = None
for label, opts in variants:
    for o, v in opts:
        s.set_option(o, v)
    for _ in range(2):
        s.step()
    events.clear()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timing[0] = True
    a.record()
    for _ in range(steps):
        s.step()
    b.record()
    torch.cuda.synchronize()
    timing[0] = False
    ms = a.elapsed_time(b) / steps
    print(f"== {label}: {ms:.3f} ms/step  ({n / ms * 1e3 / 1e6:.1f} M particle-steps/s)", flush=True)
    for m, ev in sorted(events.items(), key=lambda kv: -sum(x.elapsed_time(y) for x, y in kv[1])):
        t = [x.elapsed_time(y) for x, y in ev]
        print(f"   {m:32s} n/step {len(t) / steps:5.1f}  mean {np.mean(t) * 1e3:8.1f} us  total/step {np.sum(t) / steps:7.3f} ms")
st = s.state()
print("finite:", bool(np.isfinite(st["pos"]).all()), "density max", float(st["density"].max()))
s.close()
