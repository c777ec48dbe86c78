#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hex_all_operators or golden or edge or misaligned or pipeline" > gpurun_out/pytest_hex.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_hex.log
(timeout 120 python tools/pcie_probe.py) > gpurun_out/pcie_probe.json 2>gpurun_out/pcie_probe.err; cat gpurun_out/pcie_probe.json
(timeout 500 python tools/sweep.py --shapes Hex --nm 2..11 --geom regular,deformed --ops BwdTrans,IProductWRTBase,PhysDeriv --out gpurun_out/sweep_hex_b.jsonl) > gpurun_out/sweep_hex_b.log 2>&1; echo "hex rc=$?"
