#!/bin/bash
# dense DMMA Helmholtz, 4-warp CTAs: parity, sweep with the kernel forced on, mixed-mesh bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense" > gpurun_out/pytest_dense.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_dense.log
d=1
(NEKMF_DENSE=$d timeout 120 python tools/sweep.py --shapes Tet --nm 2..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tet_$d.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=$d timeout 120 python tools/sweep.py --shapes Tri --nm 3..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tri_$d.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=$d timeout 200 python tools/sweep.py --shapes Pyr --nm 2..7 --geom regular --ops Helmholtz --reps 3 --words 16777216 --out gpurun_out/sweep_dense_pyr_$d.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=0 timeout 120 python tools/sweep.py --shapes Tet --nm 2..2 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tet2_0.jsonl) > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_dense_*_1.jsonl'))+['gpurun_out/sweep_dense_tet2_0.jsonl']:
    for l in open(f):
        r=json.loads(l)
        if 'op' in r: print(f[-12:-6], r['nm'], r['ms'], r['frac_hbm'], r.get('frac_dmma'), r['kernel'][:30])
PY
timeout 300 python bench.py --config 4 > gpurun_out/bench_config4.log 2>gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(k, v['ms'], v['kernel']) for k,v in d['per_shape'].items()]"
