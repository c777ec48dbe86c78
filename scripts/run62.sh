#!/bin/bash
mkdir -p gpurun_out
(timeout 22 python tools/sweep.py --shapes Prism --nm 2..8 --geom regular --ops Helmholtz --reps 3 --out gpurun_out/sweep_prism_gen_on.jsonl) > /dev/null 2>&1 &
(NEKMF_PRISM_GENERAL=0 timeout 22 python tools/sweep.py --shapes Prism --nm 2..8 --geom regular --ops Helmholtz --reps 3 --out gpurun_out/sweep_prism_gen_off.jsonl) > /dev/null 2>&1
wait
python - <<'PY'
import json
for f in ('on','off'):
    for l in open('gpurun_out/sweep_prism_gen_%s.jsonl'%f):
        r=json.loads(l)
        if 'op' in r: print(f, r['nm'], r['ms'], r['kernel'][:28])
PY
