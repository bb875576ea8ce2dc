# A/B launch lists of the NW stage: old thread kernel (libdartgpu_oldnw.so) vs the current one. Not a bench number.
for lib in libdartgpu_oldnw.so libdartgpu.so; do
  for wl in c2 c4; do
    export DARTGPU_LIB=$PWD/dart_b200/$lib DART_BENCH_WORKLOAD=$wl DART_BENCH_SCALE=0.06
    [ $wl = c4 ] && export DART_BENCH_MIS=10 || unset DART_BENCH_MIS
    pairs=500000; [ $wl = c4 ] && pairs=100000
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ab_${lib%.so}_$wl.csv python tools/profile_step.py $pairs 2 > /dev/null 2>&1
    echo "== $lib $wl"; python tools/launch_table.py gpurun_out/ab_${lib%.so}_$wl.csv k_encode | grep -E "k_nw|total"
  done
done
