#!/bin/bash
# shape kernels with next-batch coefficient prefetch: parity, then P=6 sweep of the coefficient-input operators
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shape_fast or golden or runtime or dense or prism_extruded or pipeline or edge" > gpurun_out/pytest_prefetch.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_prefetch.log
(timeout 300 python tools/sweep.py --shapes Quad,Tri,Prism,Tet --nm 7..7 --ops BwdTrans,Helmholtz --reps 5 --out gpurun_out/sweep_prefetch_p6.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=0 timeout 300 python tools/sweep.py --shapes Prism,Tet --nm 7..7 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_prefetch_p6_nodense.jsonl) > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_prefetch_p6*.jsonl')):
    for l in open(f):
        r=json.loads(l)
        if 'op' in r: print(r['shape'], r['op'][:5], r['geometry'][:3], r['nm'], r['ms'], r['frac_hbm'], r['kernel'][:40])
PY
