#!/bin/bash
# IPWDB collapsed-shape parity + ncu full captures of the hex nm=5 operator kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "runtime_kernels or golden or pipeline" > gpurun_out/pytest_ipwdb.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ipwdb.log
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:hex_op_kernel -o gpurun_out/prof_hexops_nm5 -f python tools/sweep.py --shapes Hex --nm 5..5 --geom regular --reps 1 --ops BwdTrans,IProductWRTBase,PhysDeriv,Helmholtz,IProductWRTDerivBase) > gpurun_out/ncu_hexops.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_hexops.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
