#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py --steps 50 --warmup 5) > gpurun_out/bench_a.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_a.log | cut -c1-3000
(timeout 200 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-1000
nproc; lscpu | grep "Model name"
