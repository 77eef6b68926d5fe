set -x
mkdir -p /tmp/rep gpurun_out/prof
python bench.py > gpurun_out/bench_default.log 2>gpurun_out/bench_default.err
python tools/kbench.py > gpurun_out/kbench_all.log 2>&1
python tools/kbench.py --sweep 48 > gpurun_out/kbench_sweep48.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/prof/r1_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
for k in gemm_fwd1_bits gemm_fwd2 gemm_dX_dDP_fused gemm_dZ1_bits_db1 gemm_dW1 gemm_dW2; do
  timeout 240 ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 1 -k regex:"gemm_bf16_tc" -o /tmp/rep/r1_$k -f python tools/kbench.py --only $k --iters 1 > gpurun_out/ncu_$k.log 2>&1
done
for k in linear_fwd_W1 linear_dx_W1 linear_adam_W1; do
  kk=$(echo $k | sed 's/_W1//; s/linear_dx/linear_dx_kernel/; s/linear_fwd/linear_fwd_kernel/; s/linear_adam/linear_adam_kernel/')
  timeout 240 ncu --set full --clock-control none --import-source on --launch-skip 3 --launch-count 1 -k regex:"$kk" -o /tmp/rep/r1_$k -f python tools/kbench.py --sweep 48 --only $k --iters 1 > gpurun_out/ncu_$k.log 2>&1
done
python tools/ncu_summary.py --outdir=gpurun_out/prof /tmp/rep/r1_*.ncu-rep > gpurun_out/ncu_summary.log 2>&1
ls -la gpurun_out/prof
tail -3 gpurun_out/ncu_summary.log
