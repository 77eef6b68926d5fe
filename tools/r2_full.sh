#!/bin/bash
O=gpurun_out/full; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -s > $O/t_all.log 2>&1; echo "gpu tests rc=$?"; grep -E "passed|failed|Error" $O/t_all.log | tail -5; grep -A30 "deviation from the reference" $O/t_all.log | head -40
timeout 600 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/full/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline'].get('step_tensor_frac'), 'e2e', d['e2e']['value'])
for k in d['kernels']: print(k['kernel'][:70], k['avg_launch_ms'], k['frac'], k['share_of_step'])
PY
