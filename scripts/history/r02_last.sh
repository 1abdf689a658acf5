#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > $O/r02_gpu_suite.log 2>&1
echo "gpu suite exit $?"; tail -n 2 $O/r02_gpu_suite.log
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
echo "bench c4 exit $?"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c2u --no-cpu-baseline > $O/r02_bench_c2u.json 2> $O/r02_bench_c2u.err
echo "bench c2u exit $?"
python - <<'PY'
import json
for wl in ["c4","c2u"]:
    try:
        d=json.loads(open("gpurun_out/r02_bench_%s.json" % wl).readline())
        cpu=d.get("cpu_baseline") or {}
        print(wl, "it/s %.3f ms/it %.2f e2e %.3f setup %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["setup_s"]), "refactor %.3f ldl %.3f kkt %.3f" % (d["refactor_ms"], d["ldl_solve_ms"], d["kkt_solve_ms"]), "frac %.4f" % d["roofline"]["frac"], "| cpu %s" % cpu.get("value"), d["status"], d["iterations"], d["config"]["nnzL"], d["config"]["levels"])
    except Exception as e: print(wl, "ERR", e)
PY
