#!/bin/bash
for sk in 0 300 700 1200 2000 4000; do
JRB_FUSED_SKEW=$sk python bench.py --config C2 --steps 4 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('skew $sk', round(d['value'],2), {k:round(v,2) for k,v in d['phases_ms'].items() if k in ('density','hpsi')})"
done
