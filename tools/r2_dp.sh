#!/bin/bash
# data-parallel mode (BASELINE config 4) at this call's GPU count: overlapped bucket exchange on / off
N=${1:-2}; O=gpurun_out/dp; mkdir -p $O
if [ "$N" = "1" ]; then
timeout 300 python -m pytest tests/test_gpu_model.py -q -k "bucketed or call_plan" 2>&1 | tail -5
fi
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
[ "$N" = "1" ] && R="python"
for ov in on off; do
timeout 300 $R bench.py --gpus $N --workload dp64k --no-cpu-baseline --no-e2e --steps 40 --dp-overlap $ov > $O/dp64k_n${N}_$ov.json 2> $O/dp64k_n${N}_$ov.err; echo "dp64k n=$N overlap=$ov rc=$?"; tail -2 $O/dp64k_n${N}_$ov.err
done
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for ov in ("on","off"):
    try:
        d=json.loads(open(f'gpurun_out/dp/dp64k_n{N}_{ov}.json').read().strip().splitlines()[-1])
        print(ov, round(d['value']), 'ms/step', round(d['ms_per_step'],4), d['config']['parallelism'])
    except Exception as e:
        print(ov,'FAILED',e)
PY
