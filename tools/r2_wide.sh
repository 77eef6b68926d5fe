#!/bin/bash
O=gpurun_out/wide; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_sweep_step.py tests/test_gpu_kernels.py tests/test_gpu_train.py -q -x > $O/t.log 2>&1; echo "tests rc=$?"; tail -4 $O/t.log
B="timeout 300 python bench.py --workload sweep48_b8 --no-cpu-baseline --no-e2e"
$B --models-per-gpu 48 --steps 100 > $O/r2_bench_sweep48_b8_m48.json 2> $O/m48.err
$B --models-per-gpu 128 --steps 50 > $O/r2_bench_sweep_b8_m128.json 2> $O/m128.err
$B --models-per-gpu 24 --steps 100 > $O/m24.json 2> $O/m24.err
PGF_LINEAR_WIDE=0 $B --models-per-gpu 24 --steps 100 > $O/m24_ring.json 2> $O/m24r.err
$B --models-per-gpu 12 --steps 200 > $O/m12.json 2> $O/m12.err
$B --models-per-gpu 6 --steps 400 > $O/m6.json 2> $O/m6.err
for f in $O/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['kernel'].split(' models')[0].replace('linear_','').replace(' B=8',''):k['avg_launch_ms'] for k in d['kernels'][:6]}
    print(sys.argv[1].split('/')[-1], round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], ks)
except Exception as e: print(sys.argv[1],'FAILED',e); print(open(sys.argv[1].replace('.json','.err').replace('r2_bench_sweep48_b8_','').replace('r2_bench_sweep_b8_','')).read()[-600:])
PY
done
