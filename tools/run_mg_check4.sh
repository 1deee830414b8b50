# 4-GPU check of the native exchanges: interior ranks (two neighbours) + the weak-scaling bench line
export PYTHONUNBUFFERED=1
for cfg in "dfsph 1" "pbd 1" "wcsph 0"; do
  set -- $cfg
  SPHK_SLAB_TRANSPORT=$2 timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29620 \
     tools/slab_check.py --backend nccl --solver $1 --steps 3 --jitter 0.001 > gpurun_out/slab4_$1.log 2>&1
  grep SLAB_CHECK gpurun_out/slab4_$1.log | cut -c1-400 || tail -20 gpurun_out/slab4_$1.log
done
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench4_mg.log 2>&1
tail -1 gpurun_out/bench4_mg.log | cut -c1-300
