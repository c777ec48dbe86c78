#!/bin/bash
# final per-order sweeps of the round
mkdir -p gpurun_out
(timeout 600 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular,deformed,regular_diag --out gpurun_out/sweep_hex_final.jsonl) > gpurun_out/sweep_hex_final.log 2>&1; echo "hex rc=$?"
(timeout 600 python tools/sweep.py --shapes Quad,Tri,Prism,Pyr,Tet --nm 7..7 --out gpurun_out/sweep_p6_final.jsonl) > gpurun_out/sweep_p6_final.log 2>&1; echo "p6 rc=$?"
(timeout 300 python tools/sweep.py --shapes Quad --nm 2..8 --geom regular_diag,regular,deformed --out gpurun_out/sweep_quad_final.jsonl) > gpurun_out/sweep_quad_final.log 2>&1; echo "quad rc=$?"
(timeout 400 python bench.py --steps 50 --warmup 5) > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-300
(timeout 300 python bench.py --config 4 --steps 20 --warmup 3) > gpurun_out/bench_config4.log 2>&1; echo "config4 rc=$?"
