python tools/fullsize.py --devices 0,1,2,3,4,5,6,7 --pairs2 20000000 --pairs3 0 --pairs5 4000000 --skip-ref --expect profiles/r02_fullsize_reference_record.json --out gpurun_out/r02_fullsize_8gpu.json > gpurun_out/r02_fullsize_8gpu.log 2>&1; tail -12 gpurun_out/r02_fullsize_8gpu.log
python -m pytest tests/test_gpu_end_to_end.py -m gpu -q -k "two_gpus or sharding" 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -c 300 gpurun_out/r02_bench_8gpu.err
DART_BENCH_HUMAN=0 DART_BENCH_TOOL=0 python bench.py --steps 10 > gpurun_out/r02_bench_1of8.json 2>/dev/null
python - <<'PY'
import json
for f in ("r02_bench_8gpu", "r02_bench_1of8"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        h = l.get("human_scale") or {}
        print(f, l["n_gpus"], "value %.1f M  e2e %.1f M  ms %.3f" % (l["value"] / 1e6, l["e2e"]["value"] / 1e6, l["ms_per_step"]), "| human", {k: (round(v / 1e6, 1) if k == "value" else v) for k, v in h.items() if k in ("value", "ms_per_step", "skipped", "failed")}, "e2e", round((h.get("e2e") or {}).get("value", 0) / 1e6, 1))
    except Exception as ex:
        print(f, "failed", ex)
PY
nproc; free -g | head -2
