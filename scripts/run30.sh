#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hex_all_operators or golden" > gpurun_out/pytest_hex.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_hex.log
(timeout 600 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular,deformed,regular_diag --out gpurun_out/sweep_hex_final.jsonl) > gpurun_out/sweep_hex_final.log 2>&1; echo "hex rc=$?"
