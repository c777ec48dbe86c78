#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cg.py -m gpu -q -x -k "hex_helmholtz_coefficient_space_kernel" > gpurun_out/pytest_rows.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_rows.log
(timeout 200 python tools/sweep.py --shapes Hex --nm 7..11 --geom regular_diag --ops Helmholtz --out gpurun_out/sweep_rows.jsonl) > /dev/null 2>&1; cut -c1-60,100-280 gpurun_out/sweep_rows.jsonl
