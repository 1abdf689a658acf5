#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
CB_SOLVE_MINB=3 timeout 600 python scripts/df_trace_solve.py c4 > $O/r02_trace_solve_c4.txt 2>&1
head -14 $O/r02_trace_solve_c4.txt
