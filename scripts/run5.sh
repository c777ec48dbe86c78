#!/bin/bash
mkdir -p gpurun_out
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:kron_kernel -s 2 -c 1 -o gpurun_out/prof_kron2_r01 -f python tools/prof_helm.py --variant regular --reps 4) > gpurun_out/ncu_kron2.log 2>&1; echo "ncu rc=$?"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_occ tools/fp64_occ.cu && /tmp/fp64_occ
