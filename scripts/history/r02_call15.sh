#!/bin/bash
# round 2 (N GPUs): sharded C4 with stream-ordered NCCL exchanges, bench.py under torchrun (default = one C4 problem over N GPUs)
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_ldl_gpu.py tests/test_zz_shard_gpu.py tests/test_ipm_gpu.py -x -q -m gpu > $O/r02_call15_tests.log 2>&1
echo "tests exit $?" > $O/r02_call15_summary.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02_c15_c4_n1.json 2> $O/r02_c15_c4_n1.err
echo "bench n1 exit $?" >> $O/r02_call15_summary.txt
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --workload c2 > $O/r02_c15_c2_n1.json 2> $O/r02_c15_c2_n1.err
echo "bench c2 n1 exit $?" >> $O/r02_call15_summary.txt
timeout 600 $RUN --master-port 29512 scripts/shard_bench.py --workload c4 --reps 5 > $O/r02_shard_ldl_c4_n$N.json 2> $O/r02_shard_ldl_c4_n$N.err
echo "shard ldl c4 exit $?" >> $O/r02_call15_summary.txt
NCCL_DEBUG=INFO timeout 900 $RUN --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_c4_n$N.json 2> $O/r02_bench_c4_n$N.err
echo "bench c4 n$N exit $?" >> $O/r02_call15_summary.txt
CB_SHARD_TRANSPORT=torch timeout 900 $RUN --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_c4_n${N}_torchtransport.json 2> $O/r02_bench_c4_n${N}_torchtransport.err
echo "bench c4 n$N torch transport exit $?" >> $O/r02_call15_summary.txt
cat $O/r02_call15_summary.txt; tail -n 3 $O/r02_call15_tests.log
python - <<PY
import json,glob
for f in ["gpurun_out/r02_c15_c4_n1.json","gpurun_out/r02_c15_c2_n1.json","gpurun_out/r02_bench_c4_n$N.json","gpurun_out/r02_bench_c4_n${N}_torchtransport.json"]:
    try:
        d=json.loads(open(f).readline()); print(f, "it/s %.2f ms/it %.2f refactor %.3f ldl %.3f kkt %.3f e2e %.2f setup %.2f solves/it %.2f %s %d %s" % (d["value"], d["ms_per_step"], d["refactor_ms"], d["ldl_solve_ms"], d["kkt_solve_ms"], d["e2e"]["value"], d["e2e"]["setup_s"], d["ldl_solves_per_iteration"], d["status"], d["iterations"], d["config"]["parallelism"][:40]))
    except Exception as e: print(f, "ERR", e)
PY
head -c 700 $O/r02_shard_ldl_c4_n$N.json; echo; grep -c "NCCL INFO" $O/r02_bench_c4_n$N.err; grep -m3 "nranks\|NVLS\|Connected all" $O/r02_bench_c4_n$N.err; tail -n 4 $O/r02_bench_c4_n$N.err
