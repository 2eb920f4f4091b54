#!/bin/bash
# round 2, GPU call L (N GPUs): bench at N with the final kernels (all workloads)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
MAE_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}_l.json 2> gpurun_out/r2_bench_n${N}_l.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n${N}_l.json')); print('N$N ms/step', d['ms_per_step'], 'parity', d.get('parity_n'))
for k,v in d['workloads'].items():
    if k.startswith('c5'): print(k, round(v['ms'],3), v['roofline'].get('frac'))
    elif 'value' in v: print(k, v['value'])"
tail -2 gpurun_out/r2_bench_n${N}_l.err
