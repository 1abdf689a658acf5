#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
CB_SOLVE_MINB=4 timeout 300 python scripts/df_trace_solve.py c2 > $O/r02_trace_solve_c2.txt 2>&1
CB_SOLVE_MINB=4 timeout 600 python scripts/df_trace_solve.py c4 > $O/r02_trace_solve_c4.txt 2>&1
CB_SOLVE_MINB=4 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/r02_launches_ldl_c4.csv python scripts/ldl_once.py c4 > $O/r02_ncu_ldl_c4.log 2>&1
head -c 3000 $O/r02_trace_solve_c4.txt
