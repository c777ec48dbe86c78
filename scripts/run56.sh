#!/bin/bash
# prism kernel with a 3-stage ring; then compute-sanitizer over the kernels written this session (bounded selection)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prism_extruded" > gpurun_out/pytest_prism.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_prism.log
timeout 300 python bench.py --config 4 > gpurun_out/bench_config4.log 2>gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(k, v['ms'], v.get('frac_dmma'), v['kernel']) for k,v in d['per_shape'].items()]"
K="Tet-5-17 or Tet-7-65 or Tri-7-17 or Pyr-4-17 or (prism_extruded and (7-9 or 8-3 or 2-9)) or Tet-5-False or Prism-4-True or Quad-8-True or Tri-8-False"
(timeout 150 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K") > gpurun_out/sanitizer_racecheck_b.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_b.log | cut -c1-200
(timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K") > gpurun_out/sanitizer_memcheck_b.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/sanitizer_memcheck_b.log | cut -c1-200
