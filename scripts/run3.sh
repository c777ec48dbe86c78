#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python tools/prof_helm.py --variant regular --reps 10
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:kron_kernel -s 2 -c 1 -o gpurun_out/prof_kron_r01 -f python tools/prof_helm.py --variant regular --reps 4) > gpurun_out/ncu_kron.log 2>&1; echo "ncu rc=$?"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:hex_op_kernel -s 2 -c 1 -o gpurun_out/prof_helmdef_r01 -f python tools/prof_helm.py --variant deformed --reps 4) > gpurun_out/ncu_def.log 2>&1; echo "ncu rc=$?"
