#!/bin/bash
N=${1:-2}; O=gpurun_out/dp; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
run() { tag=$1; shift; timeout 200 env "${ENVV[@]}" $R bench.py --gpus $N --workload dp64k --no-cpu-baseline --no-e2e --steps 40 "$@" > $O/v_$tag.json 2> $O/v_$tag.err; python - $tag <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/dp/v_{sys.argv[1]}.json').read().strip().splitlines()[-1]); print(sys.argv[1], round(d['value']), 'ms/step', round(d['ms_per_step'],4))
except Exception as e: print(sys.argv[1],'FAILED',e)
PY
}
ENVV=(A=1)
run off --dp-overlap off
run on5 --dp-overlap on --dp-chunks 5
run on5_r8 --dp-overlap on --dp-chunks 5 --dp-reserve-sms 8
run on5_r16 --dp-overlap on --dp-chunks 5 --dp-reserve-sms 16
run on5_r32 --dp-overlap on --dp-chunks 5 --dp-reserve-sms 32
run on1_r16 --dp-overlap on --dp-chunks 1 --dp-reserve-sms 16
ENVV=(NCCL_MAX_NCHANNELS=4)
run on5_ch4 --dp-overlap on --dp-chunks 5
run on5_ch4_r8 --dp-overlap on --dp-chunks 5 --dp-reserve-sms 8
run off_ch4 --dp-overlap off
ENVV=(NCCL_DEBUG=INFO)
timeout 100 env NCCL_DEBUG=INFO $R bench.py --gpus $N --workload dp64k --no-cpu-baseline --no-e2e --steps 5 --dp-overlap off 2>&1 | grep -iE "channels|NVLS|nChannels|Connected" | head -8
