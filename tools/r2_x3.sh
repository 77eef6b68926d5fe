#!/bin/bash
O=gpurun_out/x3; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_x3.py -q -s > $O/t_x3.log 2>&1; echo "x3 tests rc=$?"; grep -E "^\[|rel err|passed|failed|Error|error" $O/t_x3.log | head -80
timeout 600 python bench.py --no-cpu-baseline > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"; tail -3 $O/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x3/bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step']); print(json.dumps(d['config']['also']['fp32x3_synth64k'], indent=1))
PY
timeout 300 python bench.py --workload dp64k --precision fp32x3 --no-cpu-baseline --no-e2e > $O/bench_dp64k_x3.json 2> $O/bench_dp64k_x3.err; echo "dp64k x3 rc=$?"; tail -3 $O/bench_dp64k_x3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/x3/bench_dp64k_x3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['roofline'])
for k in d['kernels']: print(k['kernel'], k['avg_launch_ms'], k['frac'], k['share_of_step'])
PY
