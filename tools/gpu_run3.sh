#!/bin/bash
mkdir -p gpurun_out
{
echo "# reference CUDA sources compiled for sm_100a, run on $(nvidia-smi --query-gpu=name --format=csv,noheader)"
for v in unmodified sm100fix; do for w in den ctc; do
  echo "== variant=$v part=$w"; CCB_REF_VARIANT=$v timeout 120 python tools/ref_probe.py $w 2>&1 | grep -E "variant|ref den|ref ctc|rror" | cut -c1-400 | head -6
done; done
} > gpurun_out/ref_on_b200.txt 2>&1; cat gpurun_out/ref_on_b200.txt
echo "== golden"; timeout 300 python tests/golden/make_ref_golden.py 2>&1 | tail -5
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu3.log 2>&1; tail -25 gpurun_out/pytest_gpu3.log
for w in 16 32; do
echo "== bench warps=$w"; CCB_DEN_WARPS=$w timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err; tail -2 gpurun_out/bench_w$w.err; python -c "
import json;j=json.load(open('gpurun_out/bench_w$w.json'));print(j['value'],j['ms_per_step'],j['roofline']['kernels'],j['roofline']['frac'],j['e2e']['value'],j['reference_cuda_same_gpu'])"
done
echo "== ncu full den"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:den_ -c 2 -o gpurun_out/prof_den3 -f python bench.py --T 100 --steps 1 --warmup 0 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_full3.log 2>&1; tail -2 gpurun_out/ncu_full3.log | cut -c1-200
