#!/bin/bash
# round 2, call 4 (1 GPU): solve v2 (slab tasks, inverted pivot blocks, two right-hand sides per sweep)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_ldl_gpu.py tests/test_zz_shard_gpu.py tests/test_ipm_gpu.py tests/test_configs_gpu.py -x -q -m gpu > $O/r02_call4_tests.log 2>&1
echo "tests exit $?" > $O/r02_call4_summary.txt
for mb in 3 2 4; do
CB_SOLVE_MINB=$mb CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c2 --no-cpu-baseline > $O/r02_c4b_c2_mb$mb.json 2> $O/r02_c4b_c2_mb$mb.err
echo "bench c2 minb $mb exit $?" >> $O/r02_call4_summary.txt
CB_SOLVE_MINB=$mb CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c4 --no-cpu-baseline > $O/r02_c4b_c4_mb$mb.json 2> $O/r02_c4b_c4_mb$mb.err
echo "bench c4 minb $mb exit $?" >> $O/r02_call4_summary.txt
done
cat $O/r02_call4_summary.txt; tail -n 5 $O/r02_call4_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02_c4b_*.json")):
    try:
        d=json.load(open(f)); print(f, "it/s %.2f ms/it %.2f refactor %.3f ldl %.3f kkt %.3f e2e %.2f setup %.2f frac %.3f solves/it %.2f %s %d" % (d["value"], d["ms_per_step"], d["refactor_ms"], d["ldl_solve_ms"], d["kkt_solve_ms"], d["e2e"]["value"], d["e2e"]["setup_s"], d["roofline"]["frac"], d["ldl_solves_per_iteration"], d["status"], d["iterations"]))
    except Exception as e: print(f, "ERR", e)
PY
grep "solve plan:" $O/r02_c4b_c2_mb3.err $O/r02_c4b_c4_mb3.err | tail -n 4
