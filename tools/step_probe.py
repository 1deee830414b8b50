"""Runs a few steps of one workload through the python mirror (same C-ABI calls as the C++ classes).
Meant to be wrapped by ncu:  ncu ... python tools/step_probe.py 2m dfsph 2"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import engine
name = sys.argv[1] if len(sys.argv) > 1 else "2m"
solver = sys.argv[2] if len(sys.argv) > 2 else "dfsph"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
sc = pkg.scene.benchmark_scene(name, solver)
s = engine.SphkSystem(sc)
print("launches after ctor:", s.launch_count(), flush=True)
for _ in range(steps):
    s.step()
s.synchronize()
print("launches total:", s.launch_count(), s.list_stats() if solver != "pbd" else "")
s.close()
