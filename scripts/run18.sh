#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or misaligned or edge or cpp" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pipe.log
for mb in 8 16 32 64; do NEKMF_HOST_CHUNK_MB=$mb timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $mb MB e2e ms', d['e2e']['ms_per_step'], 'clocks', d['clocks'])"; done
