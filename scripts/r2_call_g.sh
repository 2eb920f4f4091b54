#!/bin/bash
# round 2, GPU call G (1 GPU): full GPU suite, bench N=1 (all workloads incl. N2), per-role profile + ncu of the fused head kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -rs > gpurun_out/r2_pytest_gpu_g.log 2>&1
tail -4 gpurun_out/r2_pytest_gpu_g.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_g.json 2> gpurun_out/r2_bench_n1_g.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_g.json')); print('N1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'] and d['cpu_baseline']['value'])
for k,v in d['workloads'].items():
    if k.startswith('n2'): print(k, json.dumps(v)[:700])"
tail -3 gpurun_out/r2_bench_n1_g.err
MAE_HEAD_PROF=1 timeout 100 python scripts/dev_dmol_head.py --prof > gpurun_out/r2_head_roles.txt 2>&1
cat gpurun_out/r2_head_roles.txt
timeout 100 python scripts/dev_dmol_head.py --ncu > gpurun_out/r2_head_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:dmol_head -s 1 -c 1 -o gpurun_out/r2_head python scripts/dev_dmol_head.py --ncu > gpurun_out/r2_head_ncu.log 2>&1
tail -3 gpurun_out/r2_head_ncu.log
