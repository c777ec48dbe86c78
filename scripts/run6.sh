#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_cg.py --nx 64 --ny 64 --nz 64 --iters 30 2>&1 | tail -3
timeout 300 python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30 2>&1 | tail -3
(timeout 400 python bench.py --steps 30 --warmup 5) > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
