set -x
export PYTHONUNBUFFERED=1 SPHK_BENCH_WATCHDOG_S=100
timeout -s KILL 560 python -m pytest tests -q -m gpu > gpurun_out/final_gpu_tests.log 2>&1
tail -6 gpurun_out/final_gpu_tests.log
timeout -s KILL 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
