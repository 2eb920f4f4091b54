#!/bin/bash
# round 2, GPU call N (2 GPUs): final state - full GPU suite (incl. the 2-GPU distributed tests), smoke, bench N=1 and N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rs > gpurun_out/r2_pytest_gpu_final.log 2>&1
grep -E "passed|failed|SKIPPED" gpurun_out/r2_pytest_gpu_final.log | tail -6
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_final.json 2> gpurun_out/r2_bench_n1_final.err
MAE_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_final.json 2> gpurun_out/r2_bench_n2_final.err
python -c "
import json
for f in ('gpurun_out/r2_bench_n1_final.json', 'gpurun_out/r2_bench_n2_final.json'):
    d=json.load(open(f)); print(f, 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d.get('parity_n'))
    for k,v in d['workloads'].items():
        if k.startswith('c5'): print('  ', k, round(v['ms'],3), v['roofline'].get('frac'))
        elif 'value' in v: print('  ', k, v['value'])"
