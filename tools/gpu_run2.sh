#!/bin/bash
mkdir -p gpurun_out
echo "== ref den"; timeout 120 python tools/ref_probe.py den 2>&1 | tail -4
echo "== ref ctc (sanitizer)"; timeout 300 compute-sanitizer --print-limit 3 python tools/ref_probe.py ctc > gpurun_out/ref_ctc_sanitizer.log 2>&1; grep -E "Invalid|Illegal|at 0x|in |=========     by|ref ctc" gpurun_out/ref_ctc_sanitizer.log | head -20
echo "== l2 gather"; timeout 120 ./tools/l2_gather_bench 2>&1 | tee gpurun_out/l2_gather.txt
echo "== pytest (no ref ctc)"; timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_parity.py::test_vs_reference_cuda > gpurun_out/pytest_gpu2.log 2>&1; tail -15 gpurun_out/pytest_gpu2.log
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --T 200 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
echo "== ncu full den"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/prof_den -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log | cut -c1-300; ls -la gpurun_out/
