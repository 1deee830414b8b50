# the 2-GPU benchmark scene PAST the impact (step ~41: particles crossing tens of planes per step) -- needs the stray routing
set -x
export PYTHONUNBUFFERED=1 SPHK_BENCH_WATCHDOG_S=120
timeout -s KILL 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29721 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/bench_n2_long.json 2> gpurun_out/bench_n2_long.err
tail -3 gpurun_out/bench_n2_long.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_n2_long.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "e2e", d["e2e"].get("ms_per_step"), "parity", d["parity_checked"], d["max_rel_err"], "strays", d["strays"])
except Exception as e:
    print("no bench line:", e)
PY
timeout -s KILL 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29722 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_n2.json | head -2
