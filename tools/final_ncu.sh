#!/bin/bash
# ncu part of the round-end evidence (see final_profiles.sh): launch lists + full captures only.
set -x
R=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/final_launches_config1.csv python tools/profile_step.py 1000000 2 > $R/final_ncu_l1.log 2>&1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/final_launches_config2_fullsize.csv python tools/profile_step.py 1000000 2 > $R/final_ncu_l2.log 2>&1
DART_BENCH_WORKLOAD=c4 DART_BENCH_SCALE=1.0 DART_BENCH_MIS=10 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/final_launches_config3_fullsize.csv python tools/profile_step.py 100000 2 > $R/final_ncu_l3.log 2>&1
KR='regex:k_search|k_phase|k_kmer_scan|k_nw_thread|k_nw$|k_read_final|k_sort_cluster_small'
ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 12 --launch-count 12 -o $R/final_prof_config1 -f python tools/profile_step.py 500000 2 > $R/final_ncu_f1.log 2>&1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 timeout 900 ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 12 --launch-count 12 -o $R/final_prof_config2_fullsize -f python tools/profile_step.py 500000 2 > $R/final_ncu_f2.log 2>&1
ls -la $R/final_prof_*
