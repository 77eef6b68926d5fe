#!/bin/bash
N=${1:-2}; shift; O=gpurun_out/fan; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for c in "$@"; do
timeout 400 $R bench.py --gpus $N --no-also --no-cpu-baseline --fanout on --fanout-ctas $c > $O/fan_n${N}_c$c.json 2> $O/fan_n${N}_c$c.err; echo "n=$N ctas=$c rc=$?"
python - $N $c <<'PY'
import json,sys
N,v=sys.argv[1:]
try:
    d=json.loads(open(f'gpurun_out/fan/fan_n{N}_c{v}.json').read().strip().splitlines()[-1])
    print('ctas',v, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ratio', round(d['e2e']['value']/d['value'],4))
except Exception as e: print(v,'FAILED',e)
PY
done
