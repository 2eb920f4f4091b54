#!/bin/bash
# round 2, GPU call K (1 GPU): full GPU suite + full bench at the final code
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -rs > gpurun_out/r2_pytest_gpu_k.log 2>&1
tail -25 gpurun_out/r2_pytest_gpu_k.log | grep -vE "Warning|warn|^$"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_k.json 2> gpurun_out/r2_bench_n1_k.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_k.json')); print('N1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'], 'traffic', d['roofline']['traffic'])
for k,v in d['workloads'].items():
    if k.startswith('c5'): print(k, round(v['ms'],3), v['roofline'].get('frac'), v['path'][-36:])
    elif k.startswith('n3') or k.startswith('full'): print(k, json.dumps(v)[:330])
    elif 'value' in v: print(k, v['value'])"
tail -3 gpurun_out/r2_bench_n1_k.err
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref_k.json 2> gpurun_out/r2_bench_ref_k.err; head -c 400 gpurun_out/r2_bench_ref_k.json
