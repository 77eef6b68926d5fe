#!/bin/bash
# round-2 record of the fused sweep step: full GPU test suite, then the B=8 sweep bench (BASELINE config 3 as written) in its variants
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2/t_all.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/r2/t_all.log
B="timeout 300 python bench.py --workload sweep48_b8 --steps 400 --no-cpu-baseline"
$B --models-per-gpu 6 > gpurun_out/r2/b8_m6.json 2> gpurun_out/r2/b8_m6.err
$B --models-per-gpu 6 --no-pdl --no-e2e > gpurun_out/r2/b8_m6_nopdl.json 2>&1
$B --models-per-gpu 6 --graph-steps 0 --no-e2e > gpurun_out/r2/b8_m6_direct.json 2>&1
$B --models-per-gpu 6 --plan off --steps 100 --no-e2e > gpurun_out/r2/b8_m6_planoff.json 2>&1
$B --models-per-gpu 48 --steps 100 --no-e2e > gpurun_out/r2/b8_m48.json 2>&1
$B --models-per-gpu 128 --steps 50 --no-e2e > gpurun_out/r2/b8_m128.json 2>&1
for f in gpurun_out/r2/b8_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1]).read()[-800:])
PY
done
