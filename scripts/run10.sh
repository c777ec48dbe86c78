#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular,deformed,regular_diag --out gpurun_out/sweep_hex.jsonl) > gpurun_out/sweep_hex.log 2>&1; echo "hex rc=$?"; tail -2 gpurun_out/sweep_hex.log | cut -c1-300
(timeout 900 python tools/sweep.py --shapes Tet,Prism,Quad,Tri --nm 7..7 --out gpurun_out/sweep_p6.jsonl) > gpurun_out/sweep_p6.log 2>&1; echo "p6 rc=$?"; tail -2 gpurun_out/sweep_p6.log | cut -c1-300
