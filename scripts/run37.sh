#!/bin/bash
# end-of-round validation and recorded numbers
mkdir -p gpurun_out
(timeout 120 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
(timeout 300 python bench.py --impl reference --steps 20 --warmup 3) > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
(timeout 400 python bench.py --steps 50 --warmup 5) > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-200
(timeout 300 python bench.py --config 1 --steps 20 --warmup 3) > gpurun_out/bench_config1.log 2>&1; echo "config1 rc=$?"
(timeout 300 python bench.py --config 4 --steps 20 --warmup 3) > gpurun_out/bench_config4.log 2>&1; echo "config4 rc=$?"
(timeout 600 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular,deformed,regular_diag --out gpurun_out/sweep_hex_final.jsonl) > /dev/null 2>&1; echo "hex rc=$?"
(timeout 600 python tools/sweep.py --shapes Quad,Tri,Prism,Pyr,Tet --nm 7..7 --out gpurun_out/sweep_p6_final.jsonl) > /dev/null 2>&1; echo "p6 rc=$?"
(timeout 300 python tools/sweep.py --shapes Quad --nm 2..8 --geom regular_diag,regular,deformed --out gpurun_out/sweep_quad_final.jsonl) > /dev/null 2>&1; echo "quad rc=$?"
(timeout 300 python tools/sweep.py --shapes Tri --nm 2..8 --geom regular,deformed --out gpurun_out/sweep_tri_final.jsonl) > /dev/null 2>&1; echo "tri rc=$?"
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
(timeout 300 python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n1.log 2>&1; echo "cg1 rc=$?"; tail -1 gpurun_out/cg_n1.log | cut -c1-400
