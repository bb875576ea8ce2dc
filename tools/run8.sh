# 8-GPU weak-scaling run on one box (the driver runs its own at round end): topology, the bench at N = 8, and N = 1 on the same box.
lscpu | grep -E "Model name|Socket|NUMA node|^CPU\(s\)" > gpurun_out/r02_8gpu_box.txt; nvidia-smi topo -m >> gpurun_out/r02_8gpu_box.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err; tail -c 300 gpurun_out/r02_bench_8gpu.err
DART_BENCH_HUMAN=0 DART_BENCH_TOOL=0 python bench.py --steps 10 > gpurun_out/r02_bench_1of8.json 2>/dev/null
python - <<'PY'
import json
for f in ("r02_bench_8gpu", "r02_bench_1of8"):
    try:
        l = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        h = l.get("human_scale") or {}
        print(f, l["n_gpus"], "value %.1f M  copy %.1f M  e2e %.1f M  ms %.3f" % (l["value"] / 1e6, l["value_with_result_copy"]["value"] / 1e6, l["e2e"]["value"] / 1e6, l["ms_per_step"]), l["host_link"], l["host_thread_bound_to_gpu_numa_node"])
        if h: print("   human value %.1f M copy %.1f M e2e %.1f M" % (h.get("value", 0) / 1e6, h.get("value_with_result_copy", 0) / 1e6, (h.get("e2e") or {}).get("value", 0) / 1e6), h.get("skipped"), h.get("failed"))
    except Exception as ex:
        print(f, "failed", ex)
PY
cat gpurun_out/r02_8gpu_box.txt | head -20
