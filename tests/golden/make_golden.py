"""Generates the golden fixtures from the REFERENCE's own CUDA kernels (oracle/_ref/libsphref.so: the
unmodified /root/reference/src/*.cu compiled for sm_100 with the reference's flags).  The reference has
no CPU path and this container has no GPU, so this script is run on a B200 through gpurun:

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'

and the produced files are then copied into tests/golden/ and committed.  Scene: "mini" (1 400 fluid +
3 752 boundary particles, same generator and constants as config 0), fixed-work solver settings of
BASELINE.md.  State after the constructor (step 0, Q3) and after each of 2 explicit steps.
Also records rcp.approx(cellLength) of the device, which lets the CPU oracle reproduce the GPU hash.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pkgload  # noqa: E402

pkg = pkgload.load()
from cpp_fluid_particles_b200 import capi, engine  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)
LIBREF = os.path.join(ROOT, "oracle", "_ref", "libsphref.so")
STEPS = 2
meta = {}
for solver in ("wcsph", "dfsph", "pbd"):
    sc = pkg.scene.benchmark_scene("mini", solver)
    app = capi.SphApp(sc, LIBREF)
    assert app.engine == "reference-cuda"
    data = {"steps": np.int32(STEPS)}
    for k in range(STEPS + 1):
        st = app.download()
        data[f"pos_{k}"], data[f"density_{k}"], data[f"p2c_{k}"] = st["pos"], st["density"], st["p2c"]
        data[f"vel_{k}"] = st["vel"]
        if k < STEPS:
            app.step()
    b = app.download_boundary()
    data["massB"], data["p2cB"] = b["mass"], b["p2c"]
    np.savez_compressed(os.path.join(out, f"mini_{solver}.npz"), **data)
    app.close()
probe = engine.SphkSystem(pkg.scene.make_scene("mini"), step0=False)
rcp = probe.device_rcp(pkg.scene.make_scene("mini").params.cell_length)
meta["rcp_cell_length_bits"] = int(np.float32(rcp).view(np.uint32))
meta["cell_length"] = float(pkg.scene.make_scene("mini").params.cell_length)
import torch  # noqa: E402
meta["device"] = torch.cuda.get_device_name(0)
with open(os.path.join(out, "meta.json"), "w") as f:
    json.dump(meta, f, indent=1)
print("golden written to", out, meta)
