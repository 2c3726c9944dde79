#!/bin/bash
# training-step check: tests + config4 bench (graph) + launch list of the step
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-tr}
timeout 900 python -m pytest -m gpu -x -q tests/test_gpu_train.py > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --workload config4 --no-cpu-baseline > $OUT/${TAG}_bench4.json 2> $OUT/${TAG}_bench4.err; echo "bench rc=$?"; tail -3 $OUT/${TAG}_bench4.err
python - <<PY
import json
for f in ('gpurun_out/${TAG}_bench4.json',):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), d['train']['cuda_graph'][:40], d['train']['gpu_launches_per_step'])
    except Exception as e: print(f,'ERR',e)
PY
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 400 -c 300 --csv --log-file $OUT/${TAG}_train_launches.csv \
   python bench.py --workload config4 --no-graph --no-cpu-baseline --steps 12 --warmup 3 > $OUT/${TAG}_ncu.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/${TAG}_train_launches.csv')) if len(r)>5 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    k=r[4][:60]; agg.setdefault(k,[0,0.0]); agg[k][0]+=1; agg[k][1]+=float(r[-1].replace(',',''))
tot=sum(v[1] for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print(f"{v[1]/1000:9.1f} us {v[0]:4d}  {k}")
print('total us', tot/1000, 'launches', len(rows))
PY
fi
