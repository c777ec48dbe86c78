#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cg.py -m gpu -q -x -k "quad or config1 or shape_fast or pipeline" > gpurun_out/pytest_quad.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_quad.log
(timeout 300 python bench.py --config 1 --steps 20 --warmup 3) > gpurun_out/bench_config1.log 2>&1; echo "config1 rc=$?"; tail -1 gpurun_out/bench_config1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d['e2e']['ms_per_step'], d['cg'])"
(timeout 300 python tools/sweep.py --shapes Quad --nm 2..8 --geom regular_diag --ops Helmholtz --out gpurun_out/sweep_quad_kron.jsonl) > gpurun_out/sweep_quad.log 2>&1; cut -c1-230 gpurun_out/sweep_quad_kron.jsonl
