#!/bin/bash
# ncu launch list (gpu__time_duration) of one evaluation: CONFIG, EMULATE ranks; prints the last evaluation
mkdir -p gpurun_out
CFG=${CFG:-C2}; ER=${ER:-1}; TAG=${TAG:-x}
ncu --metrics gpu__time_duration.sum --clock-control none -c ${COUNT:-600} --csv --log-file gpurun_out/r2_launches_${TAG}.csv python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu --emulate-ranks $ER > gpurun_out/r2_launches_${TAG}.log 2>&1
python - <<PY
import csv, re
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2_launches_${TAG}.csv') if l.startswith('"'))]
hdr=rows[0]; rows=rows[1:]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); gi=hdr.index('Grid Size')
seq=[(re.sub(r'\(.*','',r[ki]).replace('jrb::','').replace('void ',''), float(r[vi])/1e3, r[gi]) for r in rows]
idx=[i for i,(n,_,_) in enumerate(seq) if n.startswith('k_pack_energies')]
a,b=(idx[0]+1, idx[1]+1) if len(idx)>1 else (0,len(seq))
tot=0
for n,t,g in seq[a:b]:
    tot+=t; print(f'{t:8.1f} {g:>16} {n[:60]}')
print('total', round(tot,1), 'launches', b-a)
PY
