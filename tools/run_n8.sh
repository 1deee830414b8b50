set -x
export PYTHONUNBUFFERED=1 SPHK_BENCH_WATCHDOG_S=100
N=${1:-8}
timeout -s KILL 140 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -3 gpurun_out/bench_n$N.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    print("ms_per_step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"].get("ms_per_step"), "parity", d["parity_checked"], d["max_rel_err"], "strays", d["strays"]["collected_per_rank_in_the_last_step"], "imbalance", d["config"]["load_imbalance"])
except Exception as e:
    print("no bench line:", e)
PY
