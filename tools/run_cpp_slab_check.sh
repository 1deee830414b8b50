# SlabSPHSystem (C++ class API over 2 GPUs) against SPHSystem: gpurun --gpus 2 -- 'bash tools/run_cpp_slab_check.sh'
set -x
export PYTHONUNBUFFERED=1
timeout -s KILL 300 python -m pytest tests/test_gpu_slabs.py -q -m gpu -k cpp_slab > gpurun_out/cpp_slab_test.log 2>&1
tail -5 gpurun_out/cpp_slab_test.log
