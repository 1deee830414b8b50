# stray routing: shared-device gloo test (1 GPU is enough), then -- with 2 GPUs -- the native transports and the C++ system
set -x
export PYTHONUNBUFFERED=1 SPHK_BENCH_WATCHDOG_S=150
timeout -s KILL 500 python -m pytest tests/test_gpu_slabs.py -q -m gpu -k "route_fast or cpp_slab" > gpurun_out/strays_test_full.log 2>&1; tail -5 gpurun_out/strays_test_full.log
