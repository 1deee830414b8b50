# what the per-sweep halos cost a 2-GPU step (diagnostic: the NOHALO run computes wrong results on purpose)
set -x
export PYTHONUNBUFFERED=1
run() { timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 30 --warmup 5 2>gpurun_out/halo_cost_$2.err | tee gpurun_out/halo_cost_$2.json | grep -o '"ms_per_step": [0-9.]*' | head -1; }
run 29711 halos
SPHK_SLAB_DIAG_NOHALO=1 SPHK_BENCH_PARITY=0 run 29712 nohalo
# the same weak-scaling scene through the C++ class API
B="--solver dfsph --iters 4 --frames 40 --quiet --box 8 4 4 --block 256 128 128 --origin 0.725 0.105 0.725"
timeout -s KILL 180 cpp-fluid-particles_b200/sph_headless $B --ranks 2 2>&1 | tee gpurun_out/cpp_slab_4m_2gpu.json
