#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
./scripts/ubench/dmma_tile > $O/r02_dmma_tile.txt 2>&1
cat $O/r02_dmma_tile.txt
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c5 > $O/r02_bench_c5.json 2> $O/r02_bench_c5.err
echo "bench c5 exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/r02_launches_c5.csv python bench.py --gpus 1 --steps 3 --warmup 3 --workload c5 --no-cpu-baseline --no-process-warmup > $O/r02_ncu_c5.log 2>&1
echo "ncu c5 exit $?"
python - <<'PY'
import json, csv, collections
try:
    d=json.loads(open("gpurun_out/r02_bench_c5.json").readline())
    cpu=d.get("cpu_baseline") or {}
    print("c5 it/s %.3f ms/it %.2f e2e %.3f refactor %.3f ldl %.3f kkt %.3f other %.3f | cpu %s it/s %s %d" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["refactor_ms"], d["ldl_solve_ms"], d["kkt_solve_ms"], d.get("other_ms_per_step") or 0, cpu.get("value"), d["status"], d["iterations"]))
except Exception as e: print("c5 ERR", e)
tot=collections.Counter(); cnt=collections.Counter()
rows=[r for r in csv.reader(l for l in open("gpurun_out/r02_launches_c5.csv") if not l.startswith("=="))]
h=rows[0]; ik=h.index("Kernel Name"); iv=h.index("Metric Value")
for r in rows[1:]:
    try: tot[r[ik][:70]]+=float(r[iv].replace(",","")); cnt[r[ik][:70]]+=1
    except: pass
for k,v in tot.most_common(14): print("%-72s n=%5d total %10.1f us mean %8.2f us" % (k,cnt[k],v/1e3,v/1e3/cnt[k]))
PY
