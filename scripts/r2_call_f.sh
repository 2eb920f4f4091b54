#!/bin/bash
# round 2, GPU call F (4 GPUs): sharded-head parity at world 2 and 4, the maximum-size C5 test, bench N=4
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_f_gpus.txt 2>&1
timeout 1200 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs -s > gpurun_out/r2_pytest_dist_n4.log 2>&1
grep -E "sharded|passed|failed|Error" gpurun_out/r2_pytest_dist_n4.log | head -40
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "maximum_size" > gpurun_out/r2_pytest_c5max.log 2>&1
tail -3 gpurun_out/r2_pytest_c5max.log
MAE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4_f.json 2> gpurun_out/r2_bench_n4_f.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n4_f.json')); print('N4 ms/step', d['ms_per_step'], 'launches', d['launches_per_step'], 'parity', d.get('parity_n'))
for k,v in d['workloads'].items():
    if k.startswith('c5'): print(k, v['ms'], v['path'][:40], v['roofline']['frac'])
    if k.startswith('c4'): print(k, v['value'])"
tail -3 gpurun_out/r2_bench_n4_f.err
