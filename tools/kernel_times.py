"""Prints ms/step and per-kernel CUDA-event times from one bench.py run (helper for tuning)."""
import json, subprocess, sys
out = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
line = [l for l in out.stdout.splitlines() if l.startswith("{")]
if not line:
    print(out.stdout[-2000:], out.stderr[-3000:]); sys.exit(1)
d = json.loads(line[-1])
print("ms/step %.3f  value %.1f M/s  e2e %.1f M/s  launches %d" % (d["ms_per_step"], d["value"] / 1e6, d["e2e"]["value"] / 1e6, d.get("gpu_launches", 0)))
for k in d.get("roofline", {}).get("kernels", []):
    print("  %-50s %.4f ms  %8s GB/s  share %.3f" % (k["kernel"], k["ms"], "%.1f" % k["achieved_gbs"] if k["achieved_gbs"] else "-", k["share_of_step"]))
