#!/bin/bash
# numerator || den forward overlap: event trace of the fused call for each launch order / carve-out setting (under gpurun)
mkdir -p gpurun_out
for o in 0 1; do for c in 1 0; do
  echo "== order=$o carve=$c"
  CCB_OVERLAP=1 CCB_OVERLAP_TRACE=1 CCB_OVERLAP_ORDER=$o CCB_OVERLAP_CARVE=$c timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-strong 2>&1 >/dev/null | grep "overlap trace" | tail -3
done; done
