#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cg.py -m gpu -q -x -k "coefficient_space or full_size or pipeline or fused_gather or golden or hex_all" > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_full.log
(timeout 300 python tools/sweep.py --shapes Hex --nm 2..6 --geom regular --ops Helmholtz --out gpurun_out/sweep_kronfull.jsonl) > gpurun_out/sweep_kronfull.log 2>&1; cut -c1-60,100-260 gpurun_out/sweep_kronfull.jsonl
