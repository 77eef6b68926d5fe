#!/bin/bash
N=${1:-2}; O=gpurun_out/dp; mkdir -p $O
bash tools/r2_dp.sh $N
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
timeout 600 $R bench.py --gpus $N > $O/default_n$N.json 2> $O/default_n$N.err; echo "default n=$N rc=$?"; tail -2 $O/default_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
d=json.loads(open(f'gpurun_out/dp/default_n{N}.json').read().strip().splitlines()[-1])
print('default', round(d['value']), d['ms_per_step'], 'e2e', d['e2e']['value'])
a=d['config']['also']
print(json.dumps(a['dp64k'])); print(json.dumps(a['fp32x3_synth64k'])); print(a['sweep48_b8']['value'])
PY
