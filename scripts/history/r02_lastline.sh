#!/bin/bash
set -u
mkdir -p gpurun_out
CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4.json 2> gpurun_out/r02_bench_c4.err
echo "bench c4 exit $?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_c4.json").readline())
cpu=d.get("cpu_baseline") or {}
print("c4 it/s %.2f e2e %.3f setup %.2f total %.2f refactor %.2f ldl %.3f frac %.4f cpu %s %s %d" % (d["value"], d["e2e"]["value"], d["e2e"]["setup_s"], d["e2e"]["total_s"], d["refactor_ms"], d["ldl_solve_ms"], d["roofline"]["frac"], cpu.get("value"), d["status"], d["iterations"]))
PY
grep "ordering + symbolic\|uploads + device" gpurun_out/r02_bench_c4.err | tail -n 2
