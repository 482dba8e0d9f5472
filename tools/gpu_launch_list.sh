#!/bin/bash
# ncu launch list (gpu__time_duration.sum, --clock-control none) of the bench command for a config;
# per-launch times are cold-cache and serialised: use the SHARES, not the absolutes.
mkdir -p gpurun_out
CFG=${1:-C2}
ncu --metrics gpu__time_duration.sum --clock-control none -c ${2:-400} --csv \
    --log-file gpurun_out/launches_${CFG}.csv python bench.py --config $CFG --steps 2 --warmup 1 --no-cpu \
    > gpurun_out/launches_${CFG}.log 2>&1
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_${CFG}.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
tot=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    v*= {'ns':1e-3,'us':1,'ms':1e3}.get(r[ui],1e-3)
    k=r[ki].split('(')[0].replace('void jrb::','').replace('jrb::','')
    t=tot.setdefault(k,[0,0.0]); t[0]+=1; t[1]+=v
s=sum(t[1] for t in tot.values())
print('kernel launches total_us share')
for k,(n,v) in sorted(tot.items(), key=lambda kv:-kv[1][1])[:25]:
    print(f'{k[:40]:40s} {n:5d} {v:10.1f} {100*v/s:5.1f}%')
PY
