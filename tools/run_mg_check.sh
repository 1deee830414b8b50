set -x
export PYTHONUNBUFFERED=1
# re-balancing over the native transports (mailbox halos, host-free assembly), 2 ranks on 2 GPUs
for solver in dfsph pbd; do
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29651 tools/slab_check.py --backend nccl --solver $solver --steps 12 --jitter 0.001 --skew 3 --rebalance 2 > gpurun_out/rebalance_nccl_$solver.log 2>&1
grep -o 'SLAB_CHECK.*' gpurun_out/rebalance_nccl_$solver.log | python -c "import sys,json; d=json.loads(sys.stdin.read()[11:]); print('ok', d['ok'], 'rebalanced', d['rebalanced'], 'cuts', d['cuts_initial'], '->', d['cuts_final'], 'imbalance', d['imbalance_at_last_rebalance'], 'max errs', max(s['pos'] for s in d['steps']), max(s['density'] for s in d['steps']))" || tail -20 gpurun_out/rebalance_nccl_$solver.log
done
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
grep -o '"ms_per_step": [0-9.]*\|"parity_ok": [a-z]*' gpurun_out/bench_n2.json | head -4
