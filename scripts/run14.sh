#!/bin/bash
# fused CG (gather in the operator, dots in update/assemble): parity + timing at 1 and 2 GPUs
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_cg.py -m gpu -x -q > gpurun_out/pytest_cg.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_cg.log
(timeout 300 python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n1.log 2>&1; echo "cg1 rc=$?"; tail -1 gpurun_out/cg_n1.log | cut -c1-700
(timeout 300 python tools/bench_cg.py --nx 64 --ny 64 --nz 64 --iters 30) > gpurun_out/cg_n1_64.log 2>&1; echo "cg1-64 rc=$?"; tail -1 gpurun_out/cg_n1_64.log | cut -c1-700
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n2.log 2>&1; echo "cg2 rc=$?"; tail -1 gpurun_out/cg_n2.log | cut -c1-700
(timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 60 -c 40 --csv --log-file gpurun_out/launches_cg.csv python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 12) > gpurun_out/ncu_cg.log 2>&1; echo "ncu cg rc=$?"
