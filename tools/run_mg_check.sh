set -x
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/topo.log 2>&1
timeout -s KILL 400 python -m pytest tests/test_gpu_slabs.py -q -k "nccl" > gpurun_out/test_mg.log 2>&1; echo "rc=$?" >> gpurun_out/test_mg.log
tail -c 2500 gpurun_out/test_mg.log
for a in 1 0; do
SPHK_SLAB_ASYNC=$a timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 15 --warmup 3 > gpurun_out/bench2_async$a.log 2>&1
tail -c 3000 gpurun_out/bench2_async$a.log | grep -o '"ms_per_step": [0-9.]*\|"parity_ok": [a-z]*\|"host_wall_seconds_per_step_in_begin_step": [0-9.e-]*\|"assembly_ms_per_step_device": [0-9.]*' | head -6
done
