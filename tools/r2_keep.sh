#!/bin/bash
O=gpurun_out/keep; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep_step.py tests/test_gpu_kernels.py -q -x > $O/t.log 2>&1; echo "tests rc=$?"; tail -3 $O/t.log
B="timeout 300 python bench.py --workload sweep48_b8 --no-cpu-baseline --no-e2e --models-per-gpu 6 --steps 400"
for r in 1 2; do
PGF_LS_KEEP=0 $B > $O/m6_keep0_$r.json 2> $O/e0.err
$B > $O/m6_keep1_$r.json 2> $O/e1.err
done
PGF_LS_KEEP=0 timeout 300 python bench.py --workload sweep48_b8 --no-cpu-baseline --no-e2e --models-per-gpu 12 --steps 200 > $O/m12_keep0.json 2>/dev/null
PGF_LS_KEEP=1 timeout 300 python bench.py --workload sweep48_b8 --no-cpu-baseline --no-e2e --models-per-gpu 12 --steps 200 > $O/m12_keepall.json 2>/dev/null
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['kernel'].split(' models')[0].replace('linear_','').replace(' B=8',''):k['avg_launch_ms'] for k in d['kernels'][:6]}
    print(sys.argv[1].split('/')[-1], round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'], ks)
except Exception as e: print(sys.argv[1],'FAILED',e)
PY
done
