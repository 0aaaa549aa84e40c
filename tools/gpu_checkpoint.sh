#!/bin/bash
# checkpoint artefacts for profiles/: tests, full bench (both arms), ncu launch list (same command), ncu full capture of the den kernels
# usage (under gpurun): bash tools/gpu_checkpoint.sh <tag>
tag=${1:-ck}
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 400 > gpurun_out/${tag}_pytest.log 2>&1; tail -3 gpurun_out/${tag}_pytest.log
echo "== smoke"; timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; tail -2 gpurun_out/${tag}_bench.err; cut -c1-600 gpurun_out/${tag}_bench.json
echo "== bench reference arm"; timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; cut -c1-400 gpurun_out/${tag}_bench_ref.json
echo "== ncu launch list"; timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda --no-strong > gpurun_out/${tag}_ncu_bench.log 2>&1; tail -1 gpurun_out/${tag}_ncu_bench.log | cut -c1-200
echo "== ncu full"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/${tag}_prof_den -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda --no-strong > gpurun_out/${tag}_ncu_full.log 2>&1; tail -1 gpurun_out/${tag}_ncu_full.log | cut -c1-200
