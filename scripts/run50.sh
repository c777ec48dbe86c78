#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense" > gpurun_out/pytest_dense.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/pytest_dense.log
(timeout 120 python tools/sweep.py --shapes Tet --nm 2..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tet_1.jsonl) > /dev/null 2>&1
(timeout 120 python tools/sweep.py --shapes Tri --nm 3..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tri_1.jsonl) > /dev/null 2>&1
(timeout 200 python tools/sweep.py --shapes Pyr --nm 2..7 --geom regular --ops Helmholtz --reps 3 --words 16777216 --out gpurun_out/sweep_dense_pyr_1.jsonl) > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_dense_*_1.jsonl')):
    for l in open(f):
        r=json.loads(l)
        if 'op' in r: print(f[-12:-6], r['nm'], r['ms'], r['frac_hbm'], r.get('frac_dmma'), r['kernel'][:30])
PY
