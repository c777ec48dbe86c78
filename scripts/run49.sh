#!/bin/bash
# full GPU suite with the dense kernel as the default, ncu capture of the Tet P=6 dense kernel, Tri nm=2 both ways
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_full.log
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_helm_kernel -s 3 -c 1 -o gpurun_out/prof_dense_tet_nm7 -f python tools/sweep.py --shapes Tet --nm 7..7 --geom regular --reps 2 --ops Helmholtz) > gpurun_out/ncu_dense.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dense.log | cut -c1-200
for d in 0 1; do (NEKMF_DENSE=$d timeout 120 python tools/sweep.py --shapes Tri --nm 2..2 --geom regular --ops Helmholtz --reps 5 | cut -c1-330); done
