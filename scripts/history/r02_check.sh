#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 $O/r02_smoke.log
timeout 300 python -m pytest tests/test_ldl_gpu.py tests/test_zz_shard_gpu.py -x -q -m gpu 2>&1 | tail -n 1
CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_check_c4.json 2> $O/r02_check_c4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_check_c4.json").readline())
print("c4 it/s %.2f e2e %.3f setup %.2f total %.2f refactor %.2f ldl %.3f %s %d" % (d["value"], d["e2e"]["value"], d["e2e"]["setup_s"], d["e2e"]["total_s"], d["refactor_ms"], d["ldl_solve_ms"], d["status"], d["iterations"]))
PY
grep "uploads + device alloc\|big allocations" $O/r02_check_c4.err | tail -n 2
