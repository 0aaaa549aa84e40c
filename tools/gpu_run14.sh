#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest14.log 2>&1; tail -5 gpurun_out/pytest14.log
echo "== timeline"; timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd:|bwd:|ratio"
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench14.json 2> gpurun_out/bench14.err; tail -2 gpurun_out/bench14.err; python -c "
import json;j=json.load(open('gpurun_out/bench14.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'])"
