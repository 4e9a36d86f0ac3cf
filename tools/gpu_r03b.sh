#!/bin/bash
mkdir -p gpurun_out
T=r03b
( python -m pytest tests -m gpu -q 2>&1 | tail -5 ) > gpurun_out/${T}_pytest.log 2>&1
tail -3 gpurun_out/${T}_pytest.log
python tools/train_step_probe.py 2>&1 | grep -E "^step|repack|adam" 
ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 1500 --csv --log-file gpurun_out/${T}_c3_launches.csv python tools/bench_configs.py c3 > gpurun_out/${T}_c3_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/r03b_c3_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ki=h.index('Kernel Name'); vi=h.index('Metric Value')
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[hi+1:]:
    if len(r)<=vi: continue
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg[r[ki][:80]][0]+=1; agg[r[ki][:80]][1]+=v
tot=sum(v[1] for v in agg.values())
print('total captured ms', tot/1e6)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{k:80s} n={v[0]:4d} avg={v[1]/v[0]/1e3:9.2f} us share={100*v[1]/tot:5.1f}%")
PY
python tools/bench_configs.py c3 2>&1 | tail -2
