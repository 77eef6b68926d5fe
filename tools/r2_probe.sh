#!/bin/bash
mkdir -p gpurun_out/r2
for d in 0 4 8; do
PGF_LS_DBG=$d timeout 60 python tools/plan_probe.py 6 200 4 1
done
for m in 6 48; do
for d in 0 4; do
PGF_LS_DBG=$d timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2/pl_$d.csv python tools/plan_probe.py $m 2 0 0 > gpurun_out/r2/ncu_plan.log 2>&1
python - $d $m <<'PY'
import csv,sys
rows=[r for r in csv.reader(open(f'gpurun_out/r2/pl_{sys.argv[1]}.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
out=[]; tot=0
for r in rows[1:][-13:]:
    out.append(r[ik].split('(')[0].split('::')[-1][:14]+'='+r[iv]); tot+=float(r[iv])
print('M',sys.argv[2],'dbg',sys.argv[1],' '.join(out), 'sum', tot)
PY
done
done
