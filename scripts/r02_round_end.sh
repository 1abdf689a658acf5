#!/bin/bash
# round 2: the driver's round-end sequence on one GPU (full GPU suite, smoke, reference arm, default bench) + the other configurations
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > $O/r02_gpu_suite.log 2>&1
echo "gpu suite exit $?" > $O/r02_call16_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
echo "smoke exit $?" >> $O/r02_call16_summary.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_reference_c4.json 2> $O/r02_bench_reference_c4.err
echo "reference arm exit $?" >> $O/r02_call16_summary.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
echo "bench c4 exit $?" >> $O/r02_call16_summary.txt
for wl in c2 c3 c5; do
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --workload $wl > $O/r02_bench_$wl.json 2> $O/r02_bench_$wl.err
echo "bench $wl exit $?" >> $O/r02_call16_summary.txt
done
cat $O/r02_call16_summary.txt; tail -n 4 $O/r02_gpu_suite.log; cat $O/r02_smoke.log | tail -n 2
python - <<'PY'
import json
for wl in ["reference_c4","c4","c2","c3","c5"]:
    f="gpurun_out/r02_bench_%s.json" % wl
    try:
        d=json.loads(open(f).readline())
        cpu=d.get("cpu_baseline") or {}
        print(wl, "it/s %.3f ms/it %.2f e2e %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "refactor %s ldl %s kkt %s" % (d.get("refactor_ms"), d.get("ldl_solve_ms"), d.get("kkt_solve_ms")), "frac %s" % (d.get("roofline",{}).get("frac")), "| cpu %s it/s refactor %s ms kkt %s ms" % (cpu.get("value"), cpu.get("refactor_ms"), cpu.get("kkt_solve_ms")), d.get("status"), d.get("iterations"))
    except Exception as e: print(wl, "ERR", e)
PY
