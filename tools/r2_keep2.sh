#!/bin/bash
O=gpurun_out/keep; mkdir -p $O
for m in 6 12 24; do
for k in 0 1; do
PGF_LS_KEEP=$k timeout 300 python bench.py --workload sweep48_b8 --no-cpu-baseline --no-e2e --models-per-gpu $m --steps $((2400/m)) > $O/k_m${m}_$k.json 2>/dev/null
python - $O/k_m${m}_$k.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1].split('/')[-1], round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'])
except Exception as e: print(sys.argv[1],'FAILED',e)
PY
done
done
