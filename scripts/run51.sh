#!/bin/bash
# end-of-session validation with the dense DMMA Helmholtz as default: full GPU suite, smoke, bench lines, sweeps, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-260
timeout 300 python bench.py --config 4 > gpurun_out/bench_config4.log 2>gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step']); [print(k, v['ms'], v.get('frac_dmma'), v['kernel']) for k,v in d['per_shape'].items()]"
(timeout 120 python tools/sweep.py --shapes Tet --nm 2..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tet_1.jsonl) > /dev/null 2>&1
(timeout 120 python tools/sweep.py --shapes Tri --nm 3..9 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tri_1.jsonl) > /dev/null 2>&1
(timeout 200 python tools/sweep.py --shapes Pyr --nm 2..7 --geom regular --ops Helmholtz --reps 3 --words 16777216 --out gpurun_out/sweep_dense_pyr_1.jsonl) > /dev/null 2>&1
(NEKMF_DENSE=0 timeout 120 python tools/sweep.py --shapes Tet --nm 2..2 --geom regular --ops Helmholtz --reps 5 --out gpurun_out/sweep_dense_tet2_0.jsonl) > /dev/null 2>&1
(timeout 400 python tools/sweep.py --shapes Quad,Tri,Prism,Pyr,Tet --nm 7..7 --reps 5 --out gpurun_out/sweep_p6_final.jsonl) > gpurun_out/sweep_p6_final.log 2>&1; echo "p6 rc=$?"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:dense_helm_kernel -s 3 -c 1 -o gpurun_out/prof_dense_tet_nm7 -f python tools/sweep.py --shapes Tet --nm 7..7 --geom regular --reps 2 --ops Helmholtz) > gpurun_out/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:shape_op_kernel -s 3 -c 1 -o gpurun_out/prof_prism_helm_nm7 -f python tools/sweep.py --shapes Prism --nm 7..7 --geom regular --reps 2 --ops Helmholtz) > gpurun_out/ncu_prism.log 2>&1; echo "ncu prism rc=$?"
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_config4.csv python bench.py --config 4 --steps 2 --warmup 3) > gpurun_out/ncu_launch4.log 2>&1; echo "ncu launches rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep_dense_*_1.jsonl')):
    for l in open(f):
        r=json.loads(l)
        if 'op' in r: print(f[-12:-6], r['nm'], r['ms'], r['frac_hbm'], r.get('frac_dmma'), r['kernel'][:30])
PY
