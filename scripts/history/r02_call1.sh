#!/bin/bash
# round 2, call 1 (2 GPUs): first NCCL run of the subtree-sharded factorisation on C2 + the sharded IPM bench
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 400 $RUN --master-port 29511 scripts/shard_bench.py --workload c2 --reps 10 > $O/r02_shard_ldl_c2_n$N.json 2> $O/r02_shard_ldl_c2_n$N.err
echo "shard ldl c2 exit $?" > $O/r02_shard_summary.txt
NCCL_DEBUG=INFO timeout 400 $RUN --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --shard > $O/r02_bench_shard_c2_n$N.json 2> $O/r02_bench_shard_c2_n$N.err
echo "bench --shard c2 exit $?" >> $O/r02_shard_summary.txt
CB_TIMING=1 timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
echo "bench c2 exit $?" >> $O/r02_shard_summary.txt
cat $O/r02_shard_summary.txt; head -c 600 $O/r02_shard_ldl_c2_n$N.json; echo; head -c 1500 $O/r02_bench_shard_c2_n$N.json; tail -5 $O/r02_shard_ldl_c2_n$N.err $O/r02_bench_shard_c2_n$N.err
