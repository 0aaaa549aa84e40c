#!/bin/bash
mkdir -p gpurun_out
for d in 0 2; do echo "== timeline CCB_DEBUG=$d"; CCB_DEBUG=$d timeout 300 python tools/timeline.py --T 200 2>&1 | grep -E "fwd|bwd|ratio|slowest"; cp gpurun_out/timeline.npz gpurun_out/timeline_d$d.npz; done
echo "== bench"; timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench7.json 2> gpurun_out/bench7.err; tail -2 gpurun_out/bench7.err; python -c "
import json;j=json.load(open('gpurun_out/bench7.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'])"
echo "== bench w32"; CCB_DEN_WARPS=32 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench7w32.json 2> gpurun_out/bench7.err; tail -2 gpurun_out/bench7.err; python -c "
import json;j=json.load(open('gpurun_out/bench7w32.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'])"
