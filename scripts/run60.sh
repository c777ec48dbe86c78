#!/bin/bash
# PhysDeriv with next-batch input prefetch: parity, P=6 sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shape_fast or golden or runtime or pipeline or edge or physderiv" > gpurun_out/pytest_prefetch.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_prefetch.log
(timeout 300 python tools/sweep.py --shapes Quad,Tri,Prism,Tet --nm 7..7 --ops PhysDeriv --reps 5 --out gpurun_out/sweep_pd_prefetch.jsonl) > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_pd_prefetch.jsonl'):
    r=json.loads(l)
    if 'op' in r: print(r['shape'], r['op'][:5], r['geometry'][:3], r['nm'], r['ms'], r['frac_hbm'], r['kernel'][:40])
PY
