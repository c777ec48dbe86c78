#!/bin/bash
# end-of-session validation: full GPU suite, smoke, bench lines, refreshed sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; tail -1 gpurun_out/bench.log | cut -c1-200
timeout 300 python bench.py --config 4 > gpurun_out/bench_config4.log 2>gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['bound'], d['roofline']['frac']); [print(k, v['ms'], v.get('frac_dmma'), v['kernel']) for k,v in d['per_shape'].items()]"
(timeout 400 python tools/sweep.py --shapes Quad,Tri,Prism,Pyr,Tet --nm 7..7 --reps 5 --out gpurun_out/sweep_p6_final.jsonl) > gpurun_out/sweep_p6_final.log 2>&1; echo "p6 rc=$?"
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_config4.csv python bench.py --config 4 --steps 2 --warmup 3) > gpurun_out/ncu_launch4.log 2>&1; echo "ncu launches rc=$?"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:prism_helm_kernel -s 3 -c 1 -o gpurun_out/prof_prism_dmma_nm7 -f python bench.py --config 4 --steps 3 --warmup 3) > gpurun_out/ncu_prism_dmma.log 2>&1; echo "ncu rc=$?"
