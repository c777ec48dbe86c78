#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "Tri or golden or runtime or edge" > gpurun_out/pytest_tri.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_tri.log
(timeout 200 python tools/sweep.py --shapes Tri --nm 2..7 --geom regular --ops IProductWRTDerivBase --out gpurun_out/sweep_triipwdb.jsonl) > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_triipwdb.jsonl'):
    r=json.loads(l)
    if 'op' in r: print(r['nm'], r['ms'], r['frac_hbm'], r['kernel'][:24])
PY
