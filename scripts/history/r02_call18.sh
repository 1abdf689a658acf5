#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_gpu_suite.log 2>&1
echo "gpu suite exit $?"
tail -n 3 $O/r02_gpu_suite.log
for wl in c4 c2 c5 c3; do
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > $O/r02_c18_$wl.json 2> $O/r02_c18_$wl.err
echo "bench $wl exit $?"
done
python - <<'PY'
import json
for wl in ["c4","c2","c5","c3"]:
    try:
        d=json.loads(open("gpurun_out/r02_c18_%s.json" % wl).readline()); print(wl, "it/s %.2f ms/it %.2f refactor %.3f (%.0f GFLOP/s) ldl %.3f kkt %.3f e2e %.2f setup %.2f %s %d" % (d["value"], d["ms_per_step"], d["refactor_ms"], d["roofline_other"]["fp64_gflops"], d["ldl_solve_ms"], d["kkt_solve_ms"], d["e2e"]["value"], d["e2e"]["setup_s"], d["status"], d["iterations"]))
    except Exception as e: print(wl, "ERR", e)
PY
