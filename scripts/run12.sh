#!/bin/bash
# session-6 baseline: full GPU suite, smoke, both bench arms, launch list, hex + shape sweeps
mkdir -p gpurun_out
nvidia-smi -L
(timeout 120 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.log | cut -c1-400
(timeout 400 python bench.py --steps 50 --warmup 5) > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
(timeout 500 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular,deformed,regular_diag --out gpurun_out/sweep_hex.jsonl) > gpurun_out/sweep_hex.log 2>&1; echo "hex rc=$?"; tail -2 gpurun_out/sweep_hex.log | cut -c1-300
(timeout 500 python tools/sweep.py --shapes Tet,Prism,Quad,Tri --nm 7..7 --out gpurun_out/sweep_p6.jsonl) > gpurun_out/sweep_p6.log 2>&1; echo "p6 rc=$?"; tail -2 gpurun_out/sweep_p6.log | cut -c1-300
(timeout 300 python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n1.log 2>&1; echo "cg1 rc=$?"; tail -1 gpurun_out/cg_n1.log | cut -c1-600
