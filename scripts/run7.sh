#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
(timeout 400 python bench.py --steps 50 --warmup 5) > gpurun_out/bench_r01.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r01.log | cut -c1-900
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline) > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:kron_kernel -s 2 -c 1 -o gpurun_out/prof_kron3_r01 -f python tools/prof_helm.py --variant regular --reps 4) > gpurun_out/ncu_kron3.log 2>&1; echo "ncu rc=$?"
(timeout 300 ncu --set full --clock-control none --import-source on -k regex:hex_op_kernel -s 2 -c 1 -o gpurun_out/prof_sheared_r01 -f python tools/prof_helm.py --variant sheared --reps 4) > gpurun_out/ncu_sheared.log 2>&1; echo "ncu rc=$?"
timeout 100 python tools/prof_helm.py --variant sheared --reps 6
