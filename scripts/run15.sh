#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n1.log 2>&1; echo "cg1 rc=$?"; tail -1 gpurun_out/cg_n1.log | cut -c1-700
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hex_helm_kron|assemble_dot" -s 6 -c 2 -o gpurun_out/prof_cg_fused -f python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 6) > gpurun_out/ncu_cg_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_cg_full.log
