set -x
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
timeout -s KILL 330 python -m pytest tests/test_gpu_slabs.py -q -k "nccl and (mailbox or (nccl-dfsph))" > gpurun_out/test_mg.log 2>&1; echo "rc=$?" >> gpurun_out/test_mg.log
tail -c 2500 gpurun_out/test_mg.log
for t in 1 0; do
SPHK_SLAB_TRANSPORT=$t timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench2_t$t.log 2>&1
tail -c 1500 gpurun_out/bench2_t$t.log | grep -o '"ms_per_step": [0-9.]*' | head -1
done
