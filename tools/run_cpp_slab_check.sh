# SlabSPHSystem (C++ class API over 2 GPUs) against SPHSystem: gpurun --gpus 2 -- 'bash tools/run_cpp_slab_check.sh'
set -x
export PYTHONUNBUFFERED=1
nvidia-smi -L
timeout -s KILL 400 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -k cpp_slab 2>&1 | tail -40 | tee gpurun_out/cpp_slab_test.log
# the benchmark scene through the C++ API: 1 GPU against 2 GPUs (2M particles, DFSPH 4+4 fixed iterations like bench.py)
B="--solver dfsph --iters 4 --frames 40 --quiet --box 4 --block 128 128 128 --origin 0.725 0.105 0.725"
timeout -s KILL 120 cpp-fluid-particles_b200/sph_headless $B > gpurun_out/cpp_slab_1gpu.json 2>&1
timeout -s KILL 180 cpp-fluid-particles_b200/sph_headless $B --ranks 2 > gpurun_out/cpp_slab_2gpu.json 2>&1
cat gpurun_out/cpp_slab_1gpu.json gpurun_out/cpp_slab_2gpu.json
