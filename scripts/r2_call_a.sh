#!/bin/bash
# round 2, GPU call A (2 GPUs): measured peaks, distributed parity tests, full GPU suite, bench at N=1 and N=2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_a_gpus.txt 2>&1
python scripts/measure_peaks.py > gpurun_out/r2_peaks.json 2> gpurun_out/r2_peaks.err
timeout 1200 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rs -s > gpurun_out/r2_pytest_dist_n2.log 2>&1
timeout 600 python -m pytest tests -m gpu -q --ignore tests/test_gpu_distributed.py > gpurun_out/r2_pytest_gpu_a.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_a.json 2> gpurun_out/r2_bench_n1_a.err
MAE_BENCH_DEBUG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_a.json 2> gpurun_out/r2_bench_n2_a.err
tail -5 gpurun_out/r2_pytest_dist_n2.log gpurun_out/r2_pytest_gpu_a.log
head -c 600 gpurun_out/r2_bench_n1_a.json; echo; head -c 600 gpurun_out/r2_bench_n2_a.json
