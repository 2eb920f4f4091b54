#!/bin/bash
# round 2, GPU call E (2 GPUs): full GPU suite after the row-block stored-KL change, bench N=2 / N=1
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r2_pytest_gpu_e.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_e.log
MAE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2_bench_n2_e.json 2> gpurun_out/r2_bench_n2_e.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_e.json')); print('N2 ms/step', d['ms_per_step'], 'launches', d['launches_per_step'], 'parity', d.get('parity_n'))
for k,v in d['workloads'].items():
    if k.startswith('c5'): print(k, v['ms'], v['path'], v['roofline']['frac'])"
tail -3 gpurun_out/r2_bench_n2_e.err
