#!/bin/bash
# Round-end evidence run on one B200 box: GPU tests, bench lines (config[1] default, full-size config[2] and config[3]),
# launch lists and ncu --set full captures of the main kernels.  Everything lands in gpurun_out/final_*.
set -x
R=gpurun_out
python -m pytest tests -m gpu -x -q > $R/final_pytest_gpu.log 2>&1; tail -3 $R/final_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > $R/final_bench_reference_config1.json 2> $R/final_bench_reference_config1.err
python bench.py > $R/final_bench_config1.json 2> $R/final_bench_config1.err
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 DART_BENCH_REF_PAIRS=100000 timeout 900 python bench.py > $R/final_bench_config2_fullsize.json 2> $R/final_bench_config2_fullsize.err
DART_BENCH_WORKLOAD=c4 DART_BENCH_SCALE=1.0 DART_BENCH_PAIRS=100000 DART_BENCH_MIS=10 DART_BENCH_REF_PAIRS=20000 timeout 900 python bench.py > $R/final_bench_config3_fullsize.json 2> $R/final_bench_config3_fullsize.err
# launch lists (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/final_launches_config1.csv python tools/profile_step.py 1000000 2 > $R/final_ncu_l1.log 2>&1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/final_launches_config2_fullsize.csv python tools/profile_step.py 1000000 2 > $R/final_ncu_l2.log 2>&1
DART_BENCH_WORKLOAD=c4 DART_BENCH_SCALE=1.0 DART_BENCH_MIS=10 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/final_launches_config3_fullsize.csv python tools/profile_step.py 100000 2 > $R/final_ncu_l3.log 2>&1
# full captures of the second step's main kernels
KR='regex:k_search|k_phase|k_kmer_scan|k_nw_thread|k_nw$|k_read_final|k_sort_cluster_small'
python tools/profile_step.py 500000 2 > $R/final_plain1.log 2>&1 && ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 12 --launch-count 12 -o $R/final_prof_config1 -f python tools/profile_step.py 500000 2 > $R/final_ncu_f1.log 2>&1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 timeout 900 ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 12 --launch-count 12 -o $R/final_prof_config2_fullsize -f python tools/profile_step.py 500000 2 > $R/final_ncu_f2.log 2>&1
ls -la $R/final_*
