#!/bin/bash
# round 2, GPU call J (8 GPUs): world-8 sharded-head tests, C3 step A/B: one-launch exchange+regulariser kernel vs gather -> chain
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
T=tests/test_gpu_distributed.py
timeout 600 python -m pytest "$T::test_sharded_head_c3_shape[8-True]" "$T::test_sharded_head_c3_shape[8-False]" "$T::test_sharded_head_rowblock_path[8]" -m gpu -q -rs -s > gpurun_out/r2_pytest_dist_n8.log 2>&1
grep -E "sharded|passed|failed|Error" gpurun_out/r2_pytest_dist_n8.log | head
for mw in 4 8; do
MAE_ROWS_CLUSTER_MAX_WORLD=$mw timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-extra --no-cpu > gpurun_out/r2_bench_n8_maxworld$mw.json 2> gpurun_out/r2_bench_n8_maxworld$mw.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n8_maxworld$mw.json')); print('N8 max_world=$mw ms/step', d['ms_per_step'], 'blocks', d['block_ms'], 'parity', d.get('parity_n'))"
done
