#!/bin/bash
# A/B of the den kernels under gpurun: usage  bash tools/gpu_ab.sh <tag> "ENV=.. ENV=.." "ENV=.." ...   (one bench per env set)
tag=$1; shift
mkdir -p gpurun_out
i=0
for envs in "$@"; do
  i=$((i+1))
  echo "== [$i] $envs"
  env $envs timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-strong > gpurun_out/${tag}_ab$i.json 2> gpurun_out/${tag}_ab$i.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_ab$i.json"))
    k = d["roofline"]["kernels"]
    print("   step %.2f ms  den %.2f ms  " % (d["ms_per_step"], d["roofline"]["den_ms"]) + "  ".join("%s %.2f" % (n.replace("den_", "").replace("_kernel", ""), v["ms"]) for n, v in k.items()))
except Exception as e:
    print("   failed:", e); print(open("gpurun_out/${tag}_ab$i.err").read()[-800:])
PY
done
