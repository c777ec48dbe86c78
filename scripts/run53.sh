#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "prism_extruded or (shape_fast and Prism) or golden" > gpurun_out/pytest_prism.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_prism.log | cut -c1-300
timeout 300 python bench.py --config 4 > gpurun_out/bench_config4.log 2>gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(k, v['ms'], v.get('frac_dmma'), v['kernel']) for k,v in d['per_shape'].items()]"
tail -3 gpurun_out/bench_config4.err
