#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for wl in c2 c4; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_ldl_$wl.csv python scripts/ldl_once.py $wl > $O/r02_ncu_ldl_$wl.log 2>&1
done
