#!/bin/bash
# round 2, call 2 (1 GPU): C4 on the hub-separator ordering
set -u
mkdir -p gpurun_out
O=gpurun_out
CB_TIMING=1 timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c4 --no-cpu-baseline > $O/r02_bench_c4.json 2> $O/r02_bench_c4.err
echo "bench c4 exit $?" > $O/r02_call2_summary.txt
timeout 600 python -m pytest tests/test_configs_gpu.py tests/test_ldl_gpu.py -x -q -m gpu > $O/r02_call2_tests.log 2>&1
echo "tests exit $?" >> $O/r02_call2_summary.txt
cat $O/r02_call2_summary.txt; tail -n 3 $O/r02_call2_tests.log; head -c 1500 $O/r02_bench_c4.json; echo; grep "cb timing" $O/r02_bench_c4.err | awk '$NF=="s" && $(NF-1)>0.05' | tail -n 40
