#!/bin/bash
# regular IProductWRTBase with phys-input prefetch: parity, per-order sweep for Prism / Tet
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "shape_fast or golden or runtime or pipeline or edge" > gpurun_out/pytest_prefetch.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_prefetch.log
(timeout 300 python tools/sweep.py --shapes Prism,Tet --nm 2..9 --geom regular --ops IProductWRTBase --reps 5 --out gpurun_out/sweep_iprod_prefetch.jsonl) > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_iprod_prefetch.jsonl'):
    r=json.loads(l)
    if 'op' in r: print(r['shape'], r['op'][:5], r['geometry'][:3], r['nm'], r['ms'], r['frac_hbm'], r['kernel'][:40])
PY
