#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipeline or misaligned or edge or cpp or pinned" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_pipe.log
for mb in 4 8 16 32; do NEKMF_HOST_CHUNK_MB=$mb NEKMF_HOST_ZEROCOPY=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('zerocopy-out chunk $mb e2e ms', d['e2e']['ms_per_step'], d['e2e']['value'], d['checksum_l2'])"; done
