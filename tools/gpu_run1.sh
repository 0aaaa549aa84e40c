#!/bin/bash
# first GPU contact: smoke, parity tests, a short bench.  Every stage under its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_gpu.log
echo "== bench small"; timeout 300 python bench.py --steps 3 --warmup 3 --T 300 --no-cpu-baseline > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
echo "== bench full"; timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.json
