#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --ignore tests/test_gpu_distributed.py > gpurun_out/r2_pytest_gpu_b.log 2>&1
tail -3 gpurun_out/r2_pytest_gpu_b.log
for v in 0 5 6 7; do MAE_DMOL_PERSIST=$v python scripts/dmol_fb_time.py; done 2>&1 | tee gpurun_out/r2_dmol_persist_sweep.txt
MAE_DMOL_PERSIST=6 MAE_DMOL_XSTAGE=1 python scripts/dmol_fb_time.py 2>&1 | tee -a gpurun_out/r2_dmol_persist_sweep.txt
for v in 0 6 7; do MAE_DMOL_PERSIST=$v B=32 python scripts/dmol_fb_time.py; done 2>&1 | tee -a gpurun_out/r2_dmol_persist_sweep.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_bench_n1_b.json 2> gpurun_out/r2_bench_n1_b.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_b.json')); print('ms/step', d['ms_per_step'], 'launches', d['launches_per_step'], 'kernel us', d['roofline']['us_per_launch'], 'c1', d['workloads']['c1_binary_head']['us_per_step'])"
MAE_DMOL_PERSIST=7 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-extra > gpurun_out/r2_bench_n1_b_p7.json 2> gpurun_out/r2_bench_n1_b_p7.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_b_p7.json')); print('persist7 ms/step', d['ms_per_step'], 'kernel us', d['roofline']['us_per_launch'])"
