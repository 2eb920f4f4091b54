#!/bin/bash
# round 2, GPU call I (1 GPU): graph-capture / calc_nll tests, bench N=1 with the full-model workload, ncu of the benched DMOL kernel in its ring
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py tests/test_gpu_models.py -m gpu -q -x > gpurun_out/r2_pytest_graph.log 2>&1
tail -30 gpurun_out/r2_pytest_graph.log | grep -vE "^$|Warning|warn"
timeout 900 python bench.py --steps 20 --warmup 5 --quick --no-cpu > gpurun_out/r2_bench_n1_i.json 2> gpurun_out/r2_bench_n1_i.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_i.json')); print('N1 ms/step', d['ms_per_step'])
for k,v in d['workloads'].items():
    if k.startswith('full_model') or k.startswith('c5'): print(k, json.dumps(v)[:400])"
tail -3 gpurun_out/r2_bench_n1_i.err
timeout 100 python scripts/prof_dmol_ring.py > gpurun_out/r2_ring_plain.log 2>&1 && ncu --set full --clock-control none --cache-control none -k regex:dmol_bwd -s 20 -c 1 -o gpurun_out/r2_dmol_ring python scripts/prof_dmol_ring.py > gpurun_out/r2_ring_ncu.log 2>&1
tail -2 gpurun_out/r2_ring_ncu.log
