#!/bin/bash
# single-shot validation of the general-prism DMMA kernel
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prism_general or prism_extruded or (shape_fast and Prism) or golden" > gpurun_out/pytest_prism_gen.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_prism_gen.log | cut -c1-250
(timeout 40 python tools/sweep.py --shapes Prism --nm 7..7 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_prism_gen.jsonl) 2>&1 | cut -c1-330 | tail -2
