#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hex_all_operators or golden or edge or misaligned or physderiv or pipeline" > gpurun_out/pytest_pd.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_pd.log
(timeout 200 python tools/sweep.py --shapes Hex --nm 2..6 --geom regular --ops PhysDeriv --out gpurun_out/sweep_pd_slab.jsonl) > /dev/null 2>&1
(NEKMF_HEX_PD_SLAB=0 timeout 200 python tools/sweep.py --shapes Hex --nm 2..6 --geom regular --ops PhysDeriv --out gpurun_out/sweep_pd_pencil.jsonl) > /dev/null 2>&1
python - <<'PY'
import json
def load(p): return [r for r in (json.loads(l) for l in open(p)) if 'op' in r]
a=load('gpurun_out/sweep_pd_pencil.jsonl'); b=load('gpurun_out/sweep_pd_slab.jsonl')
for x,y in zip(a,b): print(x['nm'], x['ms'], x['frac_hbm'], '->', y['ms'], y['frac_hbm'], y['kernel'])
PY
