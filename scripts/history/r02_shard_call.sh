#!/bin/bash
# Multi-GPU call of round 2 (N = 2 by default):  /usr/local/graft/bin/gpurun --gpus 2 --timeout 1500 -- 'bash scripts/r02_shard_call.sh 2'
# First real-GPU run of the subtree-sharded factorisation over NCCL: LDL level (scripts/shard_bench.py), then the whole
# interior-point solve split over the GPUs (bench.py --shard), next to the replica run of the same N.
set -u
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $RUN --master-port 29511 scripts/shard_bench.py --workload c2 --reps 10 > $O/r02_shard_ldl_c2_n$N.json 2> $O/r02_shard_ldl_c2_n$N.err
echo "shard ldl c2 exit $?" > $O/r02_shard_summary.txt
timeout 900 $RUN --master-port 29512 scripts/shard_bench.py --workload c4 --reps 5 > $O/r02_shard_ldl_c4_n$N.json 2> $O/r02_shard_ldl_c4_n$N.err
echo "shard ldl c4 exit $?" >> $O/r02_shard_summary.txt
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --shard > $O/r02_bench_shard_c2_n$N.json 2> $O/r02_bench_shard_c2_n$N.err
echo "bench --shard c2 exit $?" >> $O/r02_shard_summary.txt
timeout 900 $RUN --master-port 29514 bench.py --gpus $N --steps 5 --warmup 3 --shard --workload c4 > $O/r02_bench_shard_c4_n$N.json 2> $O/r02_bench_shard_c4_n$N.err
echo "bench --shard c4 exit $?" >> $O/r02_shard_summary.txt
timeout 600 $RUN --master-port 29515 bench.py --gpus $N --steps 10 --warmup 3 > $O/r02_bench_replicas_c2_n$N.json 2> $O/r02_bench_replicas_c2_n$N.err
echo "bench replicas c2 exit $?" >> $O/r02_shard_summary.txt
cat $O/r02_shard_summary.txt; head -c 400 $O/r02_shard_ldl_c2_n$N.json; echo; head -c 400 $O/r02_bench_shard_c2_n$N.json
