#!/bin/bash
O=gpurun_out/b8k; mkdir -p $O
run() { tag=$1; m=$2; shift; shift; env "$@" timeout 200 python bench.py --workload sweep48_b8 --models-per-gpu $m --steps 40 --no-e2e --no-cpu-baseline > $O/$tag.json 2> $O/$tag.err; python - $O/$tag.json $tag <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    ks={k['kernel'].split(' models')[0].replace('linear_','').replace(' B=8',''):k['avg_launch_ms'] for k in d['kernels'][:6]}
    print(sys.argv[2], round(d['value']), 'ms', round(d['ms_per_step'],3), ks)
except Exception as e: print(sys.argv[2],'FAILED',e)
PY
}
for m in 128 48; do
run m${m}_default $m A=1
for bp in 1 2 4 9 18; do run m${m}_adam_bp$bp $m PGF_LS_ADAM_BP=$bp; done
for bp in 1 2 4 8; do run m${m}_dx_bp$bp $m PGF_LS_DX_BP=$bp; done
done
