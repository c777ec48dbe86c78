#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hex_all_operators" > gpurun_out/pytest_hex.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_hex.log
(timeout 500 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular --ops Helmholtz,IProductWRTDerivBase --out gpurun_out/sweep_hex_c.jsonl) > gpurun_out/sweep_hex_c.log 2>&1; echo "hex rc=$?"
