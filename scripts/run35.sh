#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "runtime_kernels or golden or edge or shape_fast or pipeline" > gpurun_out/pytest_gen.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gen.log
(timeout 300 python tools/sweep.py --shapes Pyr,Quad,Tet --nm 7..7 --ops Helmholtz,IProductWRTDerivBase,BwdTrans --out gpurun_out/sweep_gen.jsonl) > /dev/null 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/sweep_gen.jsonl'):
    r=json.loads(l)
    if 'op' in r and 'gen_kernel' in r['kernel']: print(r['shape'], r['op'], r['geometry'], r['ms'], r['frac_hbm'])
PY
