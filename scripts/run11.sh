#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shape_fast or runtime_kernels" > gpurun_out/pytest_shape.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_shape.log
(timeout 600 python tools/sweep.py --shapes Tet,Prism,Quad,Tri --nm 7..7 --ops BwdTrans,IProductWRTBase,PhysDeriv,Helmholtz --out gpurun_out/sweep_p6b.jsonl) > gpurun_out/sweep_p6b.log 2>&1; echo "p6 rc=$?"; cut -c1-250 gpurun_out/sweep_p6b.log | tail -40
for mb in 4 16 32; do NEKMF_HOST_CHUNK_MB=$mb timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $mb MB', d['e2e']['ms_per_step'])"; done
