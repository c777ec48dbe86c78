#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q -k "pipeline or two_gpus or cg" > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu2.log
(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline) > gpurun_out/bench_b.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_b.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 30 --warmup 5) > gpurun_out/bench_n2.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_n2.log | cut -c1-600
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n2.log 2>&1; echo "cg2 rc=$?"; tail -1 gpurun_out/cg_n2.log
(timeout 300 python tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n1.log 2>&1; echo "cg1 rc=$?"; tail -1 gpurun_out/cg_n1.log
