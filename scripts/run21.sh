#!/bin/bash
# 8-GPU scaling: operator apply (weak, no collective) and sharded CG (z-slabs, NCCL interface exchange)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8 4; do
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 50 --warmup 5) > gpurun_out/bench_n$n.log 2>&1; echo "bench$n rc=$?"; tail -1 gpurun_out/bench_n$n.log | cut -c1-420
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n tools/bench_cg.py --nx 64 --ny 128 --nz 128 --iters 30) > gpurun_out/cg_n$n.log 2>&1; echo "cg$n rc=$?"; tail -1 gpurun_out/cg_n$n.log | cut -c1-700
done
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/bench_cg.py --nx 128 --ny 128 --nz 512 --iters 30) > gpurun_out/cg_n8_weak.log 2>&1; echo "cg8 weak rc=$?"; tail -1 gpurun_out/cg_n8_weak.log | cut -c1-700
