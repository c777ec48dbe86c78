#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hex_all_operators or golden or edge or misaligned or pipeline or full_size" > gpurun_out/pytest_slab.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_slab.log
(timeout 500 python tools/sweep.py --shapes Hex --nm 2..9 --geom regular --ops BwdTrans,IProductWRTBase --out gpurun_out/sweep_slab.jsonl) > gpurun_out/sweep_slab.log 2>&1; echo "sweep rc=$?"
