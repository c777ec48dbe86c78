#!/bin/bash
mkdir -p gpurun_out
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:prism_helm_kernel -s 3 -c 1 -o gpurun_out/prof_prism_dmma_nm7 -f python bench.py --config 4 --steps 3 --warmup 3) > gpurun_out/ncu_prism_dmma.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_prism_dmma.log | cut -c1-200
