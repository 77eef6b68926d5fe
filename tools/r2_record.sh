#!/bin/bash
# round-2 record at HEAD: GPU tests, smoke, the default bench line, the B=8 sweep (BASELINE config 3 as written) and dp64k at one GPU,
# and the ncu launch lists of the default bench and of the B=8 sweep step
O=gpurun_out/r2; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "gpu tests rc=$?"; tail -3 $O/t_all.log
timeout 200 python __graft_entry__.py --smoke > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python bench.py > $O/r2_bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_cpu.json 2> $O/bench_ref.err
B="timeout 300 python bench.py --workload sweep48_b8 --no-cpu-baseline"
$B --models-per-gpu 6 --steps 400 > $O/r2_bench_sweep48_b8_m6.json 2> $O/b8_m6.err
$B --models-per-gpu 48 --steps 100 --no-e2e > $O/r2_bench_sweep48_b8_m48.json 2> $O/b8_m48.err
$B --models-per-gpu 128 --steps 50 --no-e2e > $O/r2_bench_sweep_b8_m128.json 2> $O/b8_m128.err
timeout 300 python bench.py --workload dp64k --no-cpu-baseline > $O/r2_bench_dp64k_n1.json 2> $O/dp64k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also > $O/ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_sweep48_b8.csv python bench.py --workload sweep48_b8 --models-per-gpu 6 --steps 4 --warmup 3 --graph-steps 0 --no-e2e --no-cpu-baseline > $O/ncu_launches_b8.log 2>&1
for f in $O/r2_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'frac', (d.get('roofline') or {}).get('frac'), 'launches', d.get('gpu_launches'), 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print(sys.argv[1], 'FAILED', e); print(open(sys.argv[1]).read()[-800:])
PY
done
# single-GPU view of one data-parallel rank at 8 GPUs (8,192 rows per GPU, no exchange): where its step time goes
timeout 300 python bench.py --workload dp64k --batch 8192 --no-cpu-baseline --no-e2e --steps 100 > $O/r2_bench_dp64k_rank_of_8.json 2> $O/dp8192.err
# ncu --set full captures of the round-2 kernels
mkdir -p /tmp/rep $O/prof; cp profiles/ncu_traffic.json $O/prof/ 2>/dev/null
cap() { stem=$1; regex=$2; shift; shift; timeout 240 ncu --set full --clock-control none --import-source on --launch-skip 2 --launch-count 1 -k regex:"$regex" -o /tmp/rep/$stem -f python tools/ncu_targets.py "$@" > $O/ncu_$stem.log 2>&1; }
cap r2_gemm_x3_fwd1 gemm_bf16_tc x3_fwd1
cap r2_gemm_x3_dW1 gemm_bf16_tc x3_dW1
cap r2_split3 split3_kernel split3
cap r2_wide_adam_W1 wide_adam_kernel wide_adam_W1
cap r2_wide_dx_W1 wide_dx_kernel wide_dx_W1
python tools/ncu_summary.py --outdir=$O/prof /tmp/rep/r2_*.ncu-rep > $O/ncu_summary.log 2>&1; tail -6 $O/ncu_summary.log
