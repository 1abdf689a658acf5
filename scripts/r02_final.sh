#!/bin/bash
# round 2, final: what the driver runs at round end on one GPU (smoke, reference arm, default bench)
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
echo "smoke exit $?"; tail -n 1 $O/r02_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_reference_c4.json 2> $O/r02_bench_reference_c4.err
echo "reference arm exit $?"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
echo "bench c4 exit $?"
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c2 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
echo "bench c2 exit $?"
python - <<'PY'
import json
for wl in ["reference_c4","c4","c2"]:
    try:
        d=json.loads(open("gpurun_out/r02_bench_%s.json" % wl).readline())
        cpu=d.get("cpu_baseline") or {}
        print(wl, "it/s %.3f ms/it %.2f e2e %.3f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), "refactor %s ldl %s kkt %s" % (d.get("refactor_ms"), d.get("ldl_solve_ms"), d.get("kkt_solve_ms")), "frac %s" % (d.get("roofline",{}).get("frac")), "| cpu %s it/s refactor %s ms kkt %s ms" % (cpu.get("value"), cpu.get("refactor_ms"), cpu.get("kkt_solve_ms")), d.get("status"), d.get("iterations"))
    except Exception as e: print(wl, "ERR", e)
PY
