#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_er8.csv python bench.py --config C2 --steps 2 --warmup 1 --no-cpu --emulate-ranks 8 > gpurun_out/launches_er8.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_er8.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
names=[(r[ki].split('(')[0].replace('void jrb::','').replace('jrb::',''), float(r[vi].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}.get(r[ui],1e-3)) for r in rows[1:]]
# print the launches of the 3rd evaluation (between k_focc markers)
idx=[i for i,(n,_) in enumerate(names) if n.startswith('k_focc')]
a,b=idx[2],idx[3] if len(idx)>3 else len(names)
tot=0
for n,v in names[a:b]:
    print(f'{n[:44]:44s} {v:8.1f}'); tot+=v
print('sum',tot)
PY
