#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/rows_phase_timing.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" | tee gpurun_out/r2_rows_phases_n2_v2.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs -s > gpurun_out/r2_pytest_dist_n2_d.log 2>&1
grep -E "sharded|passed|failed|Error" gpurun_out/r2_pytest_dist_n2_d.log | head -20
