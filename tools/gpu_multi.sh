#!/bin/bash
# Multi-GPU artefacts for profiles/ (under `gpurun --gpus N`): the NCCL sharding test, the restated trainer under DDP x N, and
# bench.py launched the way the driver launches it (torchrun, one rank per GPU) with its strong-scaling block.
# usage: bash tools/gpu_multi.sh <N> [tag]
N=${1:-2}; tag=${2:-mg}
mkdir -p gpurun_out
nvidia-smi -L | head -$N
if [ "$N" -ge 2 ]; then
  echo "== NCCL sharding test (world 2)"; timeout 300 python -m pytest tests/test_dist_nccl.py -m gpu -q --timeout 280 2>&1 | tail -2
fi
echo "== trainer smoke, DDP x $N"; timeout 300 python tools/trainer_smoke.py --steps 30 --gpus $N > gpurun_out/${tag}_trainer_ddp$N.log 2>&1; tail -4 gpurun_out/${tag}_trainer_ddp$N.log
echo "== bench --gpus $N (torchrun)"
if [ "$N" -ge 2 ]; then
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${tag}_bench_g$N.json 2> gpurun_out/${tag}_bench_g$N.err
else
  timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/${tag}_bench_g$N.json 2> gpurun_out/${tag}_bench_g$N.err
fi
tail -2 gpurun_out/${tag}_bench_g$N.err
python - <<PY
import json
try:
    d = json.loads([l for l in open("gpurun_out/${tag}_bench_g$N.json") if l.startswith("{")][-1])
    print("value %.0f frames/s  %.2f ms/step  e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
    for k, v in (d.get("strong") or {}).items():
        print("  strong %s: %.1f ms/step  %.0f frames/s  local utterances %d" % (k, v["ms_per_step"], v["frames_per_s"], v["local_utterances"]))
except Exception as e:
    print("failed:", e)
PY
