#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs -s -x > gpurun_out/r2_pytest_dist_n2_c.log 2>&1
grep -E "sharded|passed|failed|Error" gpurun_out/r2_pytest_dist_n2_c.log | head -30
MAE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --quick > gpurun_out/r2_bench_n2_c.json 2> gpurun_out/r2_bench_n2_c.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n2_c.json')); print('N2 ms/step', d['ms_per_step'], 'launches', d['launches_per_step'], 'parity', d.get('parity_n'), d.get('parity_n_rel_err'))"
tail -3 gpurun_out/r2_bench_n2_c.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --quick > gpurun_out/r2_bench_n1_c.json 2> gpurun_out/r2_bench_n1_c.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_c.json')); print('N1 ms/step', d['ms_per_step'], 'launches', d['launches_per_step'], 'c1', d['workloads']['c1_binary_head']['us_per_step'])"
