#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c4 --no-cpu-baseline > $O/r02_timing_c4.json 2> $O/r02_timing_c4.err
CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload c2 --no-cpu-baseline > $O/r02_timing_c2.json 2> $O/r02_timing_c2.err
nproc; grep -m1 "model name" /proc/cpuinfo
