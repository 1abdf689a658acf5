#!/bin/bash
# First gpurun call of round 2: everything that was written without a GPU at the end of round 1 gets its first run,
# then the standing evidence is refreshed.  Usage (from the repo root, on the dev container):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash scripts/r02_first_call.sh'
# Output lands in gpurun_out/r02_*; copy what should be judged into profiles/.
set -u
mkdir -p gpurun_out
O=gpurun_out
# 1. the tests that never ran on a device (nonsymmetric cones, generalised power cone, presolve), verbose, no -x
timeout 600 python -m pytest tests/test_zz_nonsym_gpu.py -q -m gpu -rA > $O/r02_nonsym_tests.log 2>&1
echo "nonsym tests exit $?" > $O/r02_summary.txt
timeout 600 python -m pytest tests/test_zz_shard_gpu.py tests/test_zz_golden.py tests/test_zz_psd_large_gpu.py tests/test_zz_equilibration_gpu.py tests/test_zz_data_updating_gpu.py tests/test_zz_algebra_gpu.py -q -m gpu -rA > $O/r02_shard_golden_tests.log 2>&1
echo "shard + golden tests exit $?" >> $O/r02_summary.txt
# 2. the whole GPU suite as the driver runs it
timeout 900 python -m pytest tests -x -q -m gpu > $O/r02_gpu_tests.log 2>&1
echo "gpu suite exit $?" >> $O/r02_summary.txt
# 3. smoke
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1
echo "smoke exit $?" >> $O/r02_summary.txt
# 4. bench: C2 (headline), with setup timing marks on stderr; then the nonsymmetric workload
CB_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_c2.json 2> $O/r02_bench_c2.err
echo "bench c2 exit $?" >> $O/r02_summary.txt
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --workload expmix > $O/r02_bench_expmix.json 2> $O/r02_bench_expmix.err
echo "bench expmix exit $?" >> $O/r02_summary.txt
# 5. launch list of the bench command (share of each kernel in the step)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/r02_launches.csv \
  python bench.py --gpus 1 --steps 2 --warmup 1 --no-process-warmup > $O/r02_ncu_bench.log 2>&1
echo "ncu launch list exit $?" >> $O/r02_summary.txt
tail -3 $O/r02_nonsym_tests.log $O/r02_gpu_tests.log; cat $O/r02_summary.txt; head -c 600 $O/r02_bench_c2.json
# 6. ordering knob written without a GPU: halo-AMD on the nested-dissection leaves (about 6 % fewer flops on C2 by the
#    symbolic counts, one more level) -- same bench, knob on; adopt as the default only if ms_per_step drops
CB_ND_HALO=1 timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > $O/r02_bench_c2_halo.json 2> $O/r02_bench_c2_halo.err
echo "bench c2 halo exit $?" >> $O/r02_summary.txt
