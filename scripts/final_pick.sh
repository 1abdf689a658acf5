#!/bin/bash
# picks the better nd_leaf of two candidates on this box, then validates and benches with it
for L in 1600 3200; do timeout 120 python scripts/ldl_tune.py one $L 0 2>&1 | grep refactor_ms; done > gpurun_out/tune_ndleaf_final.txt
cat gpurun_out/tune_ndleaf_final.txt
BEST=$(python - <<'PY'
import json
best=None
for l in open("gpurun_out/tune_ndleaf_final.txt"):
    d=json.loads(l); c=d["refactor_ms"]+5.9*d["solve_ms"]
    if best is None or c<best[0]: best=(c,d["nd_leaf"])
print(best[1])
PY
)
echo "picked nd_leaf=$BEST"
echo $BEST > gpurun_out/picked_nd_leaf.txt
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
CB_ND_LEAF=$BEST timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err
python -c "
import json;d=json.load(open('gpurun_out/bench_final.json'));print(d['value'],d['refactor_ms'],d['ldl_solve_ms'],d['kkt_solve_ms'],d['e2e']['value'],d['e2e']['setup_s'],d['e2e_resolve']['value'],d['cpu_baseline']['value'],d['roofline']['frac'],d['config']['levels'])"
