set -x
export PYTHONUNBUFFERED=1 SPHK_BENCH_WATCHDOG_S=100
timeout -s KILL 640 python -m pytest tests/test_gpu_slabs.py tests/test_gpu_system.py -q -m gpu --durations=8 > gpurun_out/focus_tests.log 2>&1
tail -25 gpurun_out/focus_tests.log
