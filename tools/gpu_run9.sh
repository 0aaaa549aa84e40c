#!/bin/bash
mkdir -p gpurun_out
echo "== timeline"; timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd|bwd|ratio|slowest|rror"; cp gpurun_out/timeline.npz gpurun_out/timeline9.npz
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -2 gpurun_out/bench9.err; python -c "
import json;j=json.load(open('gpurun_out/bench9.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'])"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/prof_den9 -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full9.log 2>&1; tail -1 gpurun_out/ncu_full9.log | cut -c1-100
