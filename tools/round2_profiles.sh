#!/bin/bash
# Round-2 ncu evidence, run on the GPU box AFTER the same commands have exited 0 without ncu (see tools/round2_collect.sh for
# how the outputs become profiles/r02_*).  Launch lists: cold-cache and serialised -> compare shares, not absolutes.
set -x
R=gpurun_out
python tools/profile_step.py 1000000 2 > $R/r02_ps_c1.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/r02_launches_config1.csv python tools/profile_step.py 1000000 2 > /dev/null 2>&1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 python tools/profile_step.py 1000000 2 > $R/r02_ps_c2.log 2>&1 || exit 1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/r02_launches_config2_fullsize.csv python tools/profile_step.py 1000000 2 > /dev/null 2>&1
DART_BENCH_WORKLOAD=c4 DART_BENCH_SCALE=1.0 DART_BENCH_MIS=10 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/r02_launches_config3_fullsize.csv python tools/profile_step.py 100000 2 > /dev/null 2>&1
KR='regex:k_search|k_phase|k_kmer_scan|k_nw_thread|k_nw$|k_read_final|k_write_records|k_pair_prune|k_sort_cluster_small|k_sam_write|k_encode'
ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 11 --launch-count 13 -o $R/r02_prof_config1 -f python tools/profile_step.py 500000 2 > $R/r02_ncu_f1.log 2>&1
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 timeout 900 ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 11 --launch-count 13 -o $R/r02_prof_config2_fullsize -f python tools/profile_step.py 500000 2 > $R/r02_ncu_f2.log 2>&1
ls -la $R/r02_prof_*
