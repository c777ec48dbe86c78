#!/bin/bash
# final validation of the round: full GPU suite, smoke, refreshed P=6 sweep of every shape / operator
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
(timeout 400 python tools/sweep.py --shapes Quad,Tri,Prism,Pyr,Tet --nm 7..7 --reps 5 --out gpurun_out/sweep_p6_final.jsonl) > gpurun_out/sweep_p6_final.log 2>&1; echo "p6 rc=$?"
timeout 300 python bench.py --config 4 > gpurun_out/bench_config4.log 2>gpurun_out/bench_config4.err; tail -1 gpurun_out/bench_config4.log | cut -c1-150
