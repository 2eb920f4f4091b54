#!/bin/bash
# round 2, GPU call H (N GPUs, N = first argument): sharded-head tests of that world size only + bench at N (all workloads)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_h_gpus_n$N.txt 2>&1
T=tests/test_gpu_distributed.py
timeout 1200 python -m pytest "$T::test_sharded_head_c3_shape[$N-True]" "$T::test_sharded_head_c3_shape[$N-False]" "$T::test_sharded_head_rowblock_path[$N]" -m gpu -q -rs -s > gpurun_out/r2_pytest_dist_n$N.log 2>&1
grep -E "sharded|passed|failed|Error" gpurun_out/r2_pytest_dist_n$N.log | head -40
MAE_BENCH_DEBUG=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n${N}_h.json 2> gpurun_out/r2_bench_n${N}_h.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n${N}_h.json')); print('N$N ms/step', d['ms_per_step'], 'launches', d['launches_per_step'], 'parity', d.get('parity_n'))
for k,v in d['workloads'].items():
    if k.startswith('c5'): print(k, round(v['ms'],3), v['path'][:40], v['roofline']['frac'])
    if k.startswith('c4') or k.startswith('n2_fused_head_eval'): print(k, v['value'])"
tail -3 gpurun_out/r2_bench_n${N}_h.err
