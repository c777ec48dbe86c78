#!/bin/bash
# dense DMMA Helmholtz: parity, then the same sweep with the kernel forced on / off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense" > gpurun_out/pytest_dense.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_dense.log
for d in 1 0; do
(NEKMF_DENSE=$d timeout 120 python tools/sweep.py --shapes Tet --nm 3..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tet_$d.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=$d timeout 120 python tools/sweep.py --shapes Tri --nm 3..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tri_$d.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=$d timeout 200 python tools/sweep.py --shapes Pyr --nm 2..7 --geom regular --ops Helmholtz --reps 3 --words 16777216 --out gpurun_out/sweep_dense_pyr_$d.jsonl) > /dev/null 2>&1
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_dense_*.jsonl')):
    for l in open(f):
        r=json.loads(l)
        if 'op' in r: print(f[-12:-6], r['nm'], r['ms'], r['frac_hbm'], r.get('frac_dmma'), r['kernel'][:30])
PY
