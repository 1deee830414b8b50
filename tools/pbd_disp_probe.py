"""Within-step displacement of the PBD projection (in units of the radius) on a benchmark scene: sizes the list skin.
Per step: the largest displacement and the fraction of particles that moved more than 0.025 / 0.05 / 0.075 R between the
neighbour search and the end of the projection (before the prediction step moves everything)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import engine
name, steps = sys.argv[1], int(sys.argv[2])
sc = pkg.scene.benchmark_scene(name, "pbd")
s = engine.SphkSystem(sc)
R = sc.params.radius
snap = {}
inner_search, inner_advect = s.search_fluid, s.advect
def search():
    inner_search(); snap["p"] = s.fluid.pos.clone()
def advect():
    d = (s.fluid.pos - snap["p"]).norm(dim=1) / R
    snap["row"] = (float(d.max()), float((d > 0.025).float().mean()), float((d > 0.05).float().mean()), float((d > 0.075).float().mean()))
    inner_advect()
s.search_fluid, s.advect = search, advect
for k in range(steps):
    s.step()
    if k % 5 == 4 or k < 3:
        print(f"step {k:4d}: max {snap['row'][0]:.4f} R   frac>0.025R {snap['row'][1]:.5f}  >0.05R {snap['row'][2]:.5f}  >0.075R {snap['row'][3]:.5f}", flush=True)
print("stats", s.list_stats())
s.close()
