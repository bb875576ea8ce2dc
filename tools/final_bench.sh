R=gpurun_out
python bench.py --impl reference --steps 2 --warmup 1 > $R/final_bench_reference_config1.json 2> $R/final_bench_reference_config1.err
python bench.py > $R/final_bench_config1.json 2> $R/final_bench_config1.err
DART_BENCH_WORKLOAD=c3 DART_BENCH_SCALE=1.0 DART_BENCH_REF_PAIRS=100000 timeout 900 python bench.py > $R/final_bench_config2_fullsize.json 2> $R/final_bench_config2_fullsize.err
DART_BENCH_WORKLOAD=c4 DART_BENCH_SCALE=1.0 DART_BENCH_PAIRS=100000 DART_BENCH_MIS=10 DART_BENCH_REF_PAIRS=20000 timeout 900 python bench.py > $R/final_bench_config3_fullsize.json 2> $R/final_bench_config3_fullsize.err
DART_BENCH_WORKLOAD=c5 DART_BENCH_SCALE=1.0 DART_BENCH_REF_PAIRS=50000 timeout 600 python bench.py > $R/final_bench_config4.json 2> $R/final_bench_config4.err
python -m pytest tests -m gpu -x -q > $R/final_pytest_gpu.log 2>&1; tail -2 $R/final_pytest_gpu.log
