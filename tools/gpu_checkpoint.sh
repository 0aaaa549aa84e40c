#!/bin/bash
# checkpoint artefacts for profiles/: tests, full bench, ncu launch list (same command), ncu full capture of the den kernels
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/ck_pytest.log 2>&1; tail -3 gpurun_out/ck_pytest.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/ck_bench.json 2> gpurun_out/ck_bench.err; tail -2 gpurun_out/ck_bench.err; cut -c1-1500 gpurun_out/ck_bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ck_bench_ref.json 2> gpurun_out/ck_bench_ref.err; cat gpurun_out/ck_bench_ref.json | cut -c1-600
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/ck_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ck_ncu_bench.log 2>&1; tail -1 gpurun_out/ck_ncu_bench.log | cut -c1-200
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/ck_prof_den -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ck_ncu_full.log 2>&1; tail -1 gpurun_out/ck_ncu_full.log | cut -c1-200
