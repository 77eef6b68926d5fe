#!/bin/bash
# the default bench line at N GPUs of one box, as the driver launches it
N=${1:-4}; O=gpurun_out/n; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $R bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_default_n$N.json 2> $O/default_n$N.err; echo "default n=$N rc=$?"
timeout 300 $R bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/ref_n$N.json 2> $O/ref_n$N.err; echo "reference arm rc=$?"; tail -c 300 $O/ref_n$N.json
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/n/r2_bench_default_n{N}.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ratio', round(d['e2e']['value']/d['value'],4), d['e2e'].get('input_path','')[:70])
a=d['config']['also']
print(' dp64k', round(a['dp64k']['value']), a['dp64k']['ms_per_step']); print(' b8', round(a['sweep48_b8']['value']), a['sweep48_b8']['ms_per_step'], round(a['sweep48_b8']['e2e']['value'])); print(' x3', round(a['fp32x3_synth64k']['fp32x3']['value']))
print(' cpu_baseline', d.get('cpu_baseline'))
PY
