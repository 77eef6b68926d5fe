#!/bin/bash
N=${1:-2}; shift; O=gpurun_out/fan; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
for v in "$@"; do
case $v in
 off) F="--fanout off";;
 p2p) F="--fanout on --fanout-mode p2p";;
 p2pr*) F="--fanout on --fanout-mode p2p --fanout-reserve ${v#p2pr}";;
 nccl*) F="--fanout on --fanout-mode nccl --fanout-ctas ${v#nccl}";;
esac
timeout 400 $R bench.py --gpus $N --no-also --no-cpu-baseline $F > $O/fan_n${N}_$v.json 2> $O/fan_n${N}_$v.err; echo "n=$N $v rc=$?"; grep -vE "OMP_NUM|\*\*\*\*" $O/fan_n${N}_$v.err | tail -3 | cut -c1-400
python - $N $v <<'PY'
import json,sys
N,v=sys.argv[1:]
try:
    d=json.loads(open(f'gpurun_out/fan/fan_n{N}_{v}.json').read().strip().splitlines()[-1])
    print(v, 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ratio', round(d['e2e']['value']/d['value'],4), d['e2e']['h2d_bytes_per_step'], d['e2e'].get('input_path'))
except Exception as e: print(v,'FAILED',e)
PY
done
