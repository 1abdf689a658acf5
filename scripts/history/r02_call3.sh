#!/bin/bash
# round 2, call 3 (N GPUs): C4 sharded (LDL level + whole IPM)
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29512 scripts/shard_bench.py --workload c4 --reps 5 > $O/r02_shard_ldl_c4_n$N.json 2> $O/r02_shard_ldl_c4_n$N.err
echo "shard ldl c4 exit $?" > $O/r02_call3_summary.txt
timeout 900 $RUN --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --shard --workload c4 > $O/r02_bench_shard_c4_n$N.json 2> $O/r02_bench_shard_c4_n$N.err
echo "bench --shard c4 exit $?" >> $O/r02_call3_summary.txt
cat $O/r02_call3_summary.txt; head -c 700 $O/r02_shard_ldl_c4_n$N.json; echo; head -c 1200 $O/r02_bench_shard_c4_n$N.json; echo; tail -n 5 $O/r02_shard_ldl_c4_n$N.err; tail -n 5 $O/r02_bench_shard_c4_n$N.err
