#!/bin/bash
mkdir -p gpurun_out
(timeout 500 python tools/sweep.py --shapes Hex --nm 6..6 --geom regular --ops BwdTrans,IProductWRTBase --out gpurun_out/sweep_slab6.jsonl) > gpurun_out/sweep_slab6.log 2>&1; cut -c1-200 gpurun_out/sweep_slab6.jsonl
