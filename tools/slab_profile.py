"""Where a multi-rank step spends its time (rank 0 prints): CUDA-event time of begin_step (migration + halo
assembly + searches) vs the solver part, and host time blocked in exchanges."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import pkgload
pkg = pkgload.load()
from cpp_fluid_particles_b200 import slabs
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
scene = sys.argv[1] if len(sys.argv) > 1 else {2: "4m", 4: "8m", 8: "16m"}[world]
sc = pkg.scene.benchmark_scene(scene, "dfsph")
s = slabs.SlabSystem(sc, rank, world, torch.device("cuda", local))
for _ in range(3):
    s.step()
ev = lambda: torch.cuda.Event(enable_timing=True)
# finer split of begin_step: wrap the pieces
import time as _t
marks = {}
def wrap(obj, name, label):
    inner = getattr(obj, name)
    def f(*a, **k):
        e0, e1 = ev(), ev(); e0.record(); t0 = _t.perf_counter(); r = inner(*a, **k); dt = _t.perf_counter() - t0; e1.record()
        marks.setdefault(label, []).append((e0, e1, dt)); return r
    setattr(obj, name, f)
wrap(s, "search_fluid", "search"); wrap(s, "build_neighbor_list", "build")
wrap(s.ex, "exchange_rows", "exchange_rows"); wrap(s, "_bounds", "bounds(sync)"); wrap(s.ex, "exchange", "field_sync")
tb, ts = 0.0, 0.0
K = 10
for _ in range(K):
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); s.begin_step(); e1.record(); s.step_dfsph(); e2.record()
    torch.cuda.synchronize()
    tb += e0.elapsed_time(e1); ts += e1.elapsed_time(e2)
if rank == 0:
    for k, v in marks.items():
        v = v[-K * max(1, len(v) // (K + 3)):]
        print(f"   {k:16s} calls/step {len(v)/K:5.1f}  gpu {sum(a.elapsed_time(b) for a, b, _ in v)/K:7.3f} ms/step  host {sum(d for _, _, d in v)/K*1e3:7.3f} ms/step")
    print(f"world {world} scene {scene}: begin_step {tb/K:.3f} ms  solver {ts/K:.3f} ms  (n_own {s.n_own}, ghosts {s.n_gl}+{s.n_gr})", flush=True)
s.close(); dist.barrier(); dist.destroy_process_group()
