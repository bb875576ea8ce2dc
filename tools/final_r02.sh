# Round-2 closing run on one GPU: tests, smoke, the bench line, the reference arm, spin-vs-sleep, refreshed config[1] ncu evidence.
R=gpurun_out
[ -n "$SKIP_TESTS" ] || { python -m pytest tests -m gpu -x -q > $R/r02_final_pytest.log 2>&1; tail -3 $R/r02_final_pytest.log; }
python -c "import __graft_entry__ as g; g.smoke()" > $R/r02_final_smoke.log 2>&1; tail -2 $R/r02_final_smoke.log
python bench.py > $R/r02_bench_config1.json 2> $R/r02_bench_config1.err; tail -2 $R/r02_bench_config1.err
python bench.py --impl reference --steps 3 --warmup 1 > $R/r02_bench_reference_config1.json 2>/dev/null
[ -n "$SKIP_TESTS" ] || DART_BENCH_HUMAN=0 DART_BENCH_TOOL=0 DARTGPU_SYNC=spin python bench.py > $R/r02_bench_config1_spin.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $R/r02_launches_config1.csv python tools/profile_step.py 1000000 2 > /dev/null 2>&1
KR='regex:k_search|k_phase|k_kmer_scan|k_nw_thread|k_nw$|k_read_final|k_write_records|k_pair_prune|k_sort_cluster_small|k_encode'
ncu --set full --clock-control none --import-source on -k "$KR" --launch-skip 11 --launch-count 13 -o $R/r02_prof_config1 -f python tools/profile_step.py 500000 2 > $R/r02_ncu_f1.log 2>&1
python - <<'PY'
import json
import os
for f in [x for x in ("r02_bench_config1", "r02_bench_config1_spin") if os.path.exists(f"gpurun_out/{x}.json")]:
    l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    h = l.get("human_scale") or {}
    print(f, "value %.1f M copy %.1f M e2e %.1f M" % (l["value"] / 1e6, l["value_with_result_copy"]["value"] / 1e6, l["e2e"]["value"] / 1e6), "roof", round(l["roofline"]["frac"], 3),
          "| human", round(h.get("value", 0) / 1e6, 1), round((h.get("e2e") or {}).get("value", 0) / 1e6, 1), (h.get("roofline_hbm") or {}).get("frac"), "| tool", (l.get("fastq_to_sam") or {}).get("gpu_reads_per_s"))
PY
