#!/bin/bash
# compute-sanitizer passes over the kernels written this session
mkdir -p gpurun_out
K="coefficient_space_kernel or segment_operators or fused_gather or hex_all_operators_default_quadrature or runtime_kernels or edge_element or misaligned"
(timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cg.py -m gpu -q -x -k "$K") > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck.log
(timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cg.py -m gpu -q -x -k "$K") > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck.log
