#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve2 -c 2 -f -o $O/r02_prof_solve_c4 python scripts/ldl_once.py c4 > $O/r02_prof_solve_c4.log 2>&1
echo "prof solve c4 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve2 -c 2 -f -o $O/r02_prof_solve_c2 python scripts/ldl_once.py c2 > $O/r02_prof_solve_c2.log 2>&1
echo "prof solve c2 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_factor|k_invert|k_fwd_leafw|k_bwd_leafw" -s 12 -c 14 -f -o $O/r02_prof_factor_c4 python scripts/ldl_once.py c4 > $O/r02_prof_factor_c4.log 2>&1
echo "prof factor c4 exit $?"
ls -la $O/*.ncu-rep
