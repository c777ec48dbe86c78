#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cg.py -m gpu -q > gpurun_out/pytest_cg.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_cg.log
(timeout 300 python bench.py --config 1 --steps 20 --warmup 3) > gpurun_out/bench_config1.log 2>&1; echo "config1 rc=$?"; tail -1 gpurun_out/bench_config1.log | cut -c1-1800
(timeout 300 python bench.py --config 4 --steps 20 --warmup 3) > gpurun_out/bench_config4.log 2>&1; echo "config4 rc=$?"; tail -1 gpurun_out/bench_config4.log | cut -c1-2500
(timeout 200 python tools/bench_cg.py --nx 64 --ny 64 --nz 64 --iters 20) > gpurun_out/cg_small.log 2>&1; echo "cg rc=$?"; tail -1 gpurun_out/cg_small.log | cut -c1-300
