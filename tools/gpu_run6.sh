#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3; do echo "== timeline CCB_DEBUG=$d"; CCB_DEBUG=$d timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd|bwd|ratio|slowest"; done
echo "== timeline warps=32"; CCB_DEN_WARPS=32 timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd|bwd|ratio|slowest"
echo "== ncu full den"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/prof_den6 -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full6.log 2>&1; tail -2 gpurun_out/ncu_full6.log | cut -c1-200
