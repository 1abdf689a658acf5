#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_ldl_gpu.py tests/test_zz_shard_gpu.py tests/test_ipm_gpu.py -x -q -m gpu > $O/r02_call9_tests.log 2>&1
echo "tests exit $?" > $O/r02_call9_summary.txt
for mb in 3 2; do
for wl in c2 c4; do
CB_SOLVE_MINB=$mb timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $O/r02_c9_${wl}_mb$mb.json 2> $O/r02_c9_${wl}_mb$mb.err
echo "bench $wl minb $mb exit $?" >> $O/r02_call9_summary.txt
done; done
cat $O/r02_call9_summary.txt; tail -n 3 $O/r02_call9_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_c9_*.json")):
    try:
        d=json.load(open(f)); print(f, "it/s %.2f ms/it %.2f refactor %.3f ldl %.3f kkt %.3f e2e %.2f setup %.2f frac %.3f solves/it %.2f %s %d" % (d["value"], d["ms_per_step"], d["refactor_ms"], d["ldl_solve_ms"], d["kkt_solve_ms"], d["e2e"]["value"], d["e2e"]["setup_s"], d["roofline"]["frac"], d["ldl_solves_per_iteration"], d["status"], d["iterations"]))
    except Exception as e: print(f, "ERR", e)
PY
for wl in c2 c4; do
CB_SOLVE_MINB=3 timeout 600 python scripts/df_trace_solve.py $wl > $O/r02_trace_solve_$wl.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/r02_launches_ldl_$wl.csv python scripts/ldl_once.py $wl > $O/r02_ncu_ldl_$wl.log 2>&1
done
