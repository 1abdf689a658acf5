#!/bin/bash
# round 2: bench.py under torchrun on N GPUs -- default (ONE C4 problem split over the N GPUs, NCCL all-gathers) and replicas
set -u
N=${1:-4}
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
NCCL_DEBUG=INFO timeout 900 $RUN --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_c4_n$N.json 2> $O/r02_bench_c4_n$N.err
echo "bench c4 sharded n$N exit $?" > $O/r02_scale_n${N}_summary.txt
timeout 900 $RUN --master-port 29516 bench.py --gpus $N --steps 10 --warmup 3 --replicas --workload c2 > $O/r02_bench_c2_replicas_n$N.json 2> $O/r02_bench_c2_replicas_n$N.err
echo "bench c2 replicas n$N exit $?" >> $O/r02_scale_n${N}_summary.txt
cat $O/r02_scale_n${N}_summary.txt
python - <<PY
import json
for f in ["gpurun_out/r02_bench_c4_n$N.json","gpurun_out/r02_bench_c2_replicas_n$N.json"]:
    try:
        d=json.loads(open(f).readline()); print(f, "it/s %.2f ms/it %.2f refactor %.3f ldl %.3f kkt %.3f e2e %.2f setup %.2f %s %d %s %s" % (d["value"], d["ms_per_step"], d["refactor_ms"], d["ldl_solve_ms"], d["kkt_solve_ms"], d["e2e"]["value"], d["e2e"]["setup_s"], d["status"], d["iterations"], d["config"]["parallelism"][:44], d.get("collectives")))
    except Exception as e: print(f, "ERR", e)
PY
grep -m2 "nranks $N" $O/r02_bench_c4_n$N.err | cut -c1-200; grep -c "NCCL INFO" $O/r02_bench_c4_n$N.err; tail -n 3 $O/r02_bench_c4_n$N.err | cut -c1-300
