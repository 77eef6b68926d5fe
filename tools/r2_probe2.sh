#!/bin/bash
for m in 6 48; do
python tools/ls_probe.py $m
PGF_LS_DBG=1 python tools/ls_probe.py $m
PGF_LS_DBG=2 python tools/ls_probe.py $m
PGF_LS_DBG=3 python tools/ls_probe.py $m
PGF_LS_STAGES=4 python tools/ls_probe.py $m
PGF_LS_STAGES=4 PGF_LS_DBG=1 python tools/ls_probe.py $m
done
