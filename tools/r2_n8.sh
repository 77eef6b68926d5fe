#!/bin/bash
N=${1:-8}; O=gpurun_out/n8; mkdir -p $O
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
timeout 500 $R bench.py --gpus $N --no-cpu-baseline > $O/r2_bench_default_n$N.json 2> $O/default_n$N.err; echo "default (fanout auto) rc=$?"
timeout 300 $R bench.py --gpus $N --no-cpu-baseline --no-also --fanout off > $O/r2_bench_default_n${N}_fanout_off.json 2> $O/off_n$N.err; echo "fanout off rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
for f in (f'r2_bench_default_n{N}.json', f'r2_bench_default_n{N}_fanout_off.json'):
    try:
        d=json.loads(open('gpurun_out/n8/'+f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'ratio', round(d['e2e']['value']/d['value'],4), d['e2e'].get('input_path','')[:80])
        a=d['config'].get('also')
        if a:
            print(' dp64k', json.dumps(a['dp64k'])[:400]); print(' b8', a['sweep48_b8']['value'], a['sweep48_b8']['ms_per_step'], a['sweep48_b8']['e2e']); print(' x3', a['fp32x3_synth64k']['fp32x3'])
    except Exception as e: print(f,'FAILED',e)
PY
grep -vE "OMP|\*\*\*" $O/default_n$N.err | tail -3 | cut -c1-300
