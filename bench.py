#!/usr/bin/env python
"""bench.py — reads/s of the per-read mapping hot path on BASELINE.json config[1]
("same genome [4.6 Mbp synthetic], 1M paired-end 2x101 bp reads, SAM output, 1 B200").

One step = one pass of the whole hot path (seeds, candidates, 8-mer re-seeding, NW gap fill, pairing, reports)
over one batch of 1 M synthetic pairs (2 M reads) per GPU.
  value  reads/s with the read batch already resident in HBM when the timed region starts
         (dartgpu_map_reads_resident), whole job over all ranks
  e2e    the same through the public C-ABI call with HOST buffers (dartgpu_map_reads): read H2D and
         result D2H inside the timed region
Ranks shard reads (each GPU maps its own contiguous 1 M-pair range, index replicated in every HBM): no collective
on the data path; torch.distributed only provides the barrier and the max-over-ranks of the timing.

`--impl reference` times the UNMODIFIED reference (oracle/_ref/dart_ref, multithreaded, all host cores) on a bounded
sample of the same workload, index-load time subtracted.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
WORK = os.environ.get("DART_BENCH_DIR", "/tmp/dart_b200_bench")
PAIRS_PER_GPU = int(os.environ.get("DART_BENCH_PAIRS", 1_000_000))
REF_SAMPLE_PAIRS = int(os.environ.get("DART_BENCH_REF_PAIRS", 200_000))
READ_LEN = 101
MIS = os.environ.get("DART_BENCH_MIS")  # None = as BASELINE names the config (no -mis)
# Default = BASELINE config[1] (what `metric` is quoted on at 1 GPU).  DART_BENCH_WORKLOAD=c3 switches to a scaled
# config[2] (multi-contig genome with gene models, spliced pairs; DART_BENCH_SCALE x 3.1 Gbp) whose Occ table no longer
# fits L2 — used for the HBM-bound roofline of k_search in profiles/, never for the headline line.
CONTEXTS = int(os.environ.get("DART_BENCH_CONTEXTS", 4))
WORKLOAD = os.environ.get("DART_BENCH_WORKLOAD", "c2")
SCALE = float(os.environ.get("DART_BENCH_SCALE", "0.06"))


def prepare_genome():
    """Config-1/2 genome (4.6 Mbp, seed 1001) + the reference's own index builder. Input preparation, not timed."""
    from dart_b200 import synth
    os.makedirs(WORK, exist_ok=True)
    idx = os.path.join(WORK, "idx" if WORKLOAD == "c2" else f"idx_{'c5' if WORKLOAD == 'c5' else 'c3'}_{SCALE}")
    g = (synth.config_genome(2) if WORKLOAD == "c2" else synth.config_genome(5, SCALE) if WORKLOAD == "c5"
         else synth.config_genome(3, SCALE))
    if not all(os.path.exists(idx + e) for e in (".bwt", ".sa", ".pac", ".ann", ".amb")):
        fa = os.path.join(WORK, "genome.fa" if WORKLOAD == "c2" else f"genome_{WORKLOAD}.fa")
        if g.total_len <= 50_000_000:
            synth.write_fasta(fa, g)
        builder = os.path.join(ROOT, "oracle", "_ref", "bwt_index")
        if g.total_len > 50_000_000:
            # the reference's single-threaded builder needs ~100 s per 186 Mbp (hours for 3.1 Gbp): stage large genomes
            # with the GPU builder, whose files are byte-identical (tests/test_index_build.py).  Staging, never timed.
            from dart_b200 import capi
            capi.index_build(g, idx + ".tmp")
        else:
            if not os.path.exists(builder):
                raise SystemExit("bench: oracle/_ref/bwt_index (the reference's index builder) is not built")
            subprocess.run([builder, fa, idx + ".tmp"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        for e in (".bwt", ".sa", ".pac", ".ann", ".amb"):
            os.replace(idx + ".tmp" + e, idx + e)
    return g, idx


def make_pairs(g, n_pairs, rank):
    from dart_b200 import synth
    if WORKLOAD == "c2":
        return synth.simulate_pairs(g, n_pairs, READ_LEN, 0.01, seed=2002 + rank)
    if WORKLOAD == "c4":   # config[3]: 2x250, 3 % substitutions + 1-3 bp indels, -mis 10
        return synth.simulate_pairs(g, n_pairs, 250, 0.03, seed=2004 + rank, frag_mean=600, frag_sd=50, frag_min=500, frag_max=900,
                                    p_ins=0.002, p_del=0.002)
    if WORKLOAD == "c5":   # config[4]: repeat-rich genome, -m -max_dup 10000 -all_sj
        return synth.simulate_pairs(g, n_pairs, READ_LEN, 0.01, seed=2005 + rank)
    return synth.simulate_pairs(g, n_pairs, READ_LEN, 0.01, seed=2003 + rank, spliced=True, frag_min=202, frag_max=500)


def as_batch(m1, m2):
    """Interleave mates as the reference's reader leaves them: mate 2 reverse-complemented (GetData.cpp:157-168)."""
    from dart_b200 import capi, synth
    n, L = m1.shape
    codes = np.empty((2 * n, L), dtype=np.uint8)
    codes[0::2] = m1
    codes[1::2] = synth.revcomp_codes(m2)
    return capi.ReadBatch.from_codes(codes)


class ClockSampler(threading.Thread):
    """nvidia-smi-equivalent sampling (NVML) of SM clock and throttle reasons during the timed region."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, device):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz, self.stop_flag = [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        self.stop_flag = True
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


def time_reference(idx, r1, r2, n_reads, cores, extra):
    """One run of the reference binary; returns wall seconds of mapping (index load measured separately)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dart_ref")
    out = os.path.join(WORK, "ref_out.sam")
    cmd = [exe, "-i", idx, "-f", r1, "-f2", r2, "-t", str(cores), "-o", out, "-j", os.path.join(WORK, "ref.junc")] + extra
    t0 = time.perf_counter()
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return time.perf_counter() - t0


def reference_setup(g, idx, n_pairs):
    from dart_b200 import synth
    m1, m2 = make_pairs(g, n_pairs, 0)
    r1, r2 = os.path.join(WORK, f"ref_{WORKLOAD}_{n_pairs}_1.fq"), os.path.join(WORK, f"ref_{WORKLOAD}_{n_pairs}_2.fq")
    if not (os.path.exists(r1) and os.path.exists(r2)):
        synth.write_fastq(r1, m1, 1); synth.write_fastq(r2, m2, 2)
    e1, e2 = os.path.join(WORK, "one_1.fq"), os.path.join(WORK, "one_2.fq")
    synth.write_fastq(e1, m1[:1], 1); synth.write_fastq(e2, m2[:1], 2)
    return r1, r2, e1, e2


def run_reference_arm(args, rank):
    if rank != 0:
        return
    g, idx = prepare_genome()
    cores = os.cpu_count() or 1
    extra = (["-mis", MIS] if MIS else []) + (["-m", "-max_dup", "10000", "-all_sj"] if WORKLOAD == "c5" else [])
    r1, r2, e1, e2 = reference_setup(g, idx, REF_SAMPLE_PAIRS)
    load = min(time_reference(idx, e1, e2, 2, cores, extra) for _ in range(2))
    for _ in range(args.warmup):
        time_reference(idx, r1, r2, 2 * REF_SAMPLE_PAIRS, cores, extra)
    t = [max(time_reference(idx, r1, r2, 2 * REF_SAMPLE_PAIRS, cores, extra) - load, 1e-6) for _ in range(args.steps)]
    per_step = float(np.mean(t))
    v = 2 * REF_SAMPLE_PAIRS / per_step
    sample = f"{REF_SAMPLE_PAIRS} pairs 2x{READ_LEN} of config[1] per step, dart_ref -t {cores}, index-load time ({load:.2f} s) subtracted"
    print(json.dumps({
        "impl": "reference", "metric": "reads/sec mapped", "value": v, "unit": "reads/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
        "config": workload_config(sample_pairs=REF_SAMPLE_PAIRS),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(sample_pairs=None):
    wl = ("BASELINE config[1]: synthetic 4.6 Mbp random genome (seed 1001), paired-end 2x101 bp, 1% substitutions, FR fragments ~N(300,30)"
          if WORKLOAD == "c2" else
          f"BASELINE config[4] scaled x{SCALE}: repeat-rich {4.6 * SCALE:.1f} Mbp genome (30% segmental duplications), 2x101 bp, -m -max_dup 10000 -all_sj" if WORKLOAD == "c5" else
          f"BASELINE config[3] scaled x{SCALE}: {int(3.1e9 * SCALE / 1e6)} Mbp genome, pairs 2x250 bp, 3% substitutions + indels" if WORKLOAD == "c4" else
          f"BASELINE config[2] scaled x{SCALE}: {int(3.1e9 * SCALE / 1e6)} Mbp genome in 24 contigs with gene models, spliced pairs 2x101 bp, 1% substitutions")
    return {"workload": wl, "pairs_per_gpu": sample_pairs or PAIRS_PER_GPU, "read_len": 250 if WORKLOAD == "c4" else READ_LEN,
            "flags": ("-mis " + MIS) if MIS else "as named (no -mis: MaxMismatch=0, SURVEY.md F3)",
            "sharding": "contiguous read range per GPU, index replicated per HBM, no collective",
            "l2": "read batch (226 MB of codes per GPU) is larger than L2; " + ("the 4.6 MB Occ table of this config is L2-resident by nature"
                                                                                 if WORKLOAD == "c2" else "the Occ table is larger than L2 (HBM gathers)")}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    # One host thread per context, and a thread that waits for its stream spins (lowest latency; measured at 2 GPUs: waiting
    # on a blocking event instead costs 23 % of `value`).  So the contexts of all ranks together must not exceed the host's
    # cores: 4 per GPU up to 4 GPUs on this 16-core box, 2 per GPU at 8.  Only if even one context per GPU does not fit
    # do the threads wait asleep (DARTGPU_SYNC=block).
    global CONTEXTS
    cores_total = os.cpu_count() or 1
    CONTEXTS = max(1, min(CONTEXTS, cores_total // world))
    if world * CONTEXTS > cores_total:
        os.environ.setdefault("DARTGPU_SYNC", "block")
    import torch
    import torch.distributed as dist
    from dart_b200 import capi
    from dart_b200.shard import shard_bounds
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        g, idx = prepare_genome()
    barrier()
    if rank != 0:
        g, idx = prepare_genome()
    params = dict(pair_end=1)
    if MIS:
        params["max_mismatch"] = int(MIS)
    if WORKLOAD == "c5":
        params.update(multi_hit=1, max_dup=10000, all_sj=1)
    # CONTEXTS contexts (one host thread each, the C-ABI's unit of concurrency) share this GPU and split the step's batch:
    # their H2D / kernels / D2H overlap on separate streams.  Host cores are divided among ranks and contexts.
    cores = os.cpu_count() or 1
    params["host_threads"] = max(1, cores // (world * CONTEXTS))
    mappers = [capi.Mapper(idx, device=local, **params) for _ in range(CONTEXTS)]
    M = mappers[0]
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    batch = as_batch(*make_pairs(g, PAIRS_PER_GPU, rank))
    n_reads = batch.n
    bounds = shard_bounds(n_reads, CONTEXTS, True)
    subs = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        off = batch.offsets[a:b + 1]
        subs.append(capi.ReadBatch(batch.bases[off[0]:off[-1]].copy(), (off - off[0]).copy()).pin())   # e2e: inputs in pinned host memory
    pool = ThreadPoolExecutor(CONTEXTS)

    def run_steps(resident, steps):
        """`steps` passes over the batch.  Every context maps its own slice `steps` times back to back; the contexts are
        not re-synchronised between steps (a barrier per step would run them in lockstep: all in their kernels, then all
        in their D2H, and nothing would overlap), only at the two ends of the timed region."""
        def worker(m, sb):
            for _ in range(steps):
                m.map_reads(sb, resident, False)
        futs = [pool.submit(worker, m, sb) for m, sb in zip(mappers, subs)]
        for f in futs:
            f.result()

    def timed(resident, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(resident, steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- device-resident arm ----
    for m, sb in zip(mappers, subs):
        m.upload_reads(sb)
    run_steps(True, args.warmup)
    sampler = ClockSampler(local); sampler.start()
    ms_res = timed(True, args.steps)
    sts = [m.stats() for m in mappers]
    # ---- end-to-end arm (host buffers in, host results out) ----
    run_steps(False, max(1, args.warmup // 2))
    ms_e2e = timed(False, args.steps)
    sts_e2e = [m.stats() for m in mappers]
    clocks = sampler.result()
    st = {k: sum(x[k] for x in sts) for k in sts[0]}          # work and kernel time summed over the contexts of this rank
    st_e2e = {k: sum(x[k] for x in sts_e2e) for k in sts_e2e[0]}

    # one context alone on the GPU over the whole batch: per-kernel times without the other contexts' kernels in between
    # (the per-step sums above are CUDA-event intervals on streams that share the device)
    M.upload_reads(batch)
    M.map_reads(batch, True, False)
    M.map_reads(batch, True, False)
    alone = M.stats()
    int32_peak = M.int32_peak()
    # kernel-only seeding over the whole batch on one context (no host orchestration, no copies): roofline of the dominant kernel
    M.seed_resident()
    ms_k = []
    for _ in range(max(3, args.steps)):
        M.seed_resident()
        ms_k.append(M.stats()["ms_search"])
    sk = M.stats()
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        search_bytes = 64 * sk["ext_blocks"] + (sk["read_bases"] + 3) // 4 + 16 * sk["seeds"]
        search_ms = float(np.mean(ms_k))
        achieved = search_bytes / (search_ms * 1e-3) / 1e9
        traffic = None
        try:   # DRAM bytes of the ncu capture, scaled from its launch (400 k reads) to this launch
            t = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_ncu.json")))
            key = {"c2": "config1", "c3": "config2_fullsize"}.get(WORKLOAD if WORKLOAD == "c2" or SCALE == 1.0 else "")
            traffic = int(t[key]["dram_bytes_per_launch"] * n_reads / t[key]["reads_in_launch"]) if key in t else None
        except Exception:
            pass
        line = {
            "metric": "reads/sec mapped", "value": world * n_reads * args.steps / (ms_res * 1e-3), "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64 (2-bit bases, 64-bit FM intervals, int32 NW)",
            "data": "synthetic", "config": workload_config(), "clocks": clocks,
            "e2e": {"value": world * n_reads * args.steps / (ms_e2e * 1e-3), "unit": "reads/s",
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"])},
            "gpu_launches": int(st["kernel_launches"]) * args.steps, "contexts_per_gpu": CONTEXTS,
            "host_sync": os.environ.get("DARTGPU_SYNC", "spin"), "host_cores": os.cpu_count(),
            "roofline": {"bound": "hbm", "kernel": "k_search (FM-index forward extension)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": int(search_bytes), "kernel_ms": search_ms,
                         "note": ("config[1]'s 4.6 MB Occ table is L2-resident, so this kernel is bound by the integer pipes / L2, "
                                  "not HBM; the fraction is reported against the HBM copy peak as the contract asks") if WORKLOAD == "c2"
                                 else "Occ table larger than L2: 64-byte gathers from HBM"},
            "kernels_ms_per_step": {k: st[k] for k in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_report", "ms_d2h", "ms_host")},
            "work_per_step": {k: int(st[k]) for k in ("ext_steps", "ext_blocks", "lf_steps", "hits", "seeds", "nw_jobs", "nw_cells", "kmer_jobs", "kmer_window_bases")},
            "kernels_ms_one_context_alone": {k: alone[k] for k in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_report", "ms_d2h")},
            "nw_gcups": (alone["nw_cells"] / (alone["ms_nw"] * 1e-3) / 1e9) if alone["ms_nw"] > 0 else None,
            "seed_gbs": achieved,
        }
        if line["nw_gcups"]:
            ops_per_cell = 22    # adds, max and compares of the restated recurrence + traceback flags (nw_kernel.cu), per cell
            line["nw_roofline"] = {"bound": "int32", "achieved_gcups": line["nw_gcups"], "int32_ops_per_s_measured": int32_peak,
                                   "ops_per_cell": ops_per_cell, "peak_gcups": int32_peak / ops_per_cell / 1e9,
                                   "frac": line["nw_gcups"] * 1e9 * ops_per_cell / int32_peak,
                                   "note": "whole NW stage of one context alone (shape sort + both kernels + tracebacks), cells = sum of m*n"}
        if world == 1:
            try:
                cores = os.cpu_count() or 1
                extra = (["-mis", MIS] if MIS else []) + (["-m", "-max_dup", "10000", "-all_sj"] if WORKLOAD == "c5" else [])
                sp = min(REF_SAMPLE_PAIRS, 100_000 if WORKLOAD != "c4" else 40_000)
                r1, r2, e1, e2 = reference_setup(g, idx, sp)
                load = min(time_reference(idx, e1, e2, 2, cores, extra) for _ in range(2))
                t = max(time_reference(idx, r1, r2, 2 * sp, cores, extra) - load, 1e-6)
                line["cpu_baseline"] = {"value": 2 * sp / t, "unit": "reads/s", "cores": cores, "kind": "reference",
                                        "sample": f"{sp} pairs of the same workload, oracle/_ref/dart_ref -t {cores}, index-load time subtracted"}
            except Exception as ex:  # the baseline is reported, never a reason to lose the measurement
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
        print(json.dumps(line))
    for m in mappers:
        m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
