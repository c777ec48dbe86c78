#!/bin/bash
mkdir -p gpurun_out
(timeout 120 python __graft_entry__.py smoke) > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -4 gpurun_out/smoke.log
(timeout 60 ./tools/fp64_peak) > gpurun_out/fp64_peak.log 2>&1; cat gpurun_out/fp64_peak.log
(timeout 400 python bench.py --steps 30 --warmup 5) > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench.log
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline) > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
(timeout 400 ncu --set full --clock-control none --import-source on -k regex:hex_op_kernel -s 6 -c 2 -o gpurun_out/prof_helm_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline) > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
