#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest10.log 2>&1; tail -5 gpurun_out/pytest10.log
echo "== timeline"; timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd|bwd|ratio|slowest|rror"
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench10.json 2> gpurun_out/bench10.err; tail -2 gpurun_out/bench10.err; python -c "
import json;j=json.load(open('gpurun_out/bench10.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'])"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/prof_den10 -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full10.log 2>&1; tail -1 gpurun_out/ncu_full10.log | cut -c1-100
