#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest22.log 2>&1; tail -4 gpurun_out/pytest22.log
show() { python -c "
import json,sys
j=json.load(open('$1')); r=j['roofline']
print('$2', round(j['value']), round(j['ms_per_step'],2), {k:round(v['ms'],2) for k,v in r['kernels'].items()}, round(r['frac'],4), 'e2e', round(j['e2e']['ms_per_step'],2), j.get('raw_logit_entry'))"; }
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench22.json 2> gpurun_out/bench22.err || tail -3 gpurun_out/bench22.err
show gpurun_out/bench22.json default
for cfg in "u4b8:" "ubwd2:CCB_U_BWD=2" "ufwd2_ubwd2:CCB_U_FWD=2 CCB_U_BWD=2"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda --N 256 --T 3000 --varlen > gpurun_out/bench22_$name.json 2> gpurun_out/bench22_$name.err || tail -3 gpurun_out/bench22_$name.err
  show gpurun_out/bench22_$name.json varlen256_$name
done
echo "== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches22.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu22.log 2>&1; tail -1 gpurun_out/ncu22.log | cut -c1-200
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/launches22.csv')) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4][:70]; v=float(r[-1]); a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in agg.items(): print("%-72s x%-3d avg %.3f ms"%(k,c,t/c/1e6 if t/c>1e4 else t/c/1e3))
PY
