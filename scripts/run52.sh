#!/bin/bash
# IProductWRTDerivBase in the compile-time sized Quad/Tri/Prism/Tet kernels: parity, then P=6 and per-order sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shape_fast or golden or runtime or pipeline" > gpurun_out/pytest_ipwdb_shape.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ipwdb_shape.log
(timeout 300 python tools/sweep.py --shapes Prism,Tet --nm 2..9 --ops IProductWRTDerivBase --reps 5 --out gpurun_out/sweep_ipwdb_shape3d.jsonl) > /dev/null 2>&1
(timeout 300 python tools/sweep.py --shapes Quad,Tri --nm 2..9 --geom deformed --ops IProductWRTDerivBase --reps 5 --out gpurun_out/sweep_ipwdb_shape2d_def.jsonl) > /dev/null 2>&1
(NEKMF_QUAD_LANE=0 NEKMF_TRI_LANE=0 timeout 300 python tools/sweep.py --shapes Quad,Tri --nm 5..9 --geom regular --ops IProductWRTDerivBase --reps 5 --out gpurun_out/sweep_ipwdb_shape2d_reg_nolane.jsonl) > /dev/null 2>&1
(timeout 300 python tools/sweep.py --shapes Quad,Tri --nm 5..9 --geom regular --ops IProductWRTDerivBase --reps 5 --out gpurun_out/sweep_ipwdb_shape2d_reg.jsonl) > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_ipwdb_shape*.jsonl')):
    for l in open(f):
        r=json.loads(l)
        if 'op' in r: print(f[24:-6], r['shape'], r['geometry'][:3], r['nm'], r['ms'], r['frac_hbm'], r['kernel'][:34])
PY
