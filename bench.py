#!/usr/bin/env python
"""bench.py — reads/s of the per-read mapping hot path on BASELINE.json config[1]
("same genome [4.6 Mbp synthetic], 1M paired-end 2x101 bp reads, SAM output, 1 B200").

One step = one pass of the whole hot path (seeds, candidates, 8-mer re-seeding, NW gap fill, pairing, reports)
over one batch of 1 M synthetic pairs (2 M reads) per GPU.  ONE host thread per GPU drives CONTEXTS contexts through
dartgpu_submit / dartgpu_wait (PARTS batches per step) and sleeps while it waits.
  value   reads/s with the reads AND the result records resident in HBM (no PCIe in the timed region), whole job over all ranks
          (`value_with_result_copy`: the records copied to page-locked host memory every step — round 1's definition)
  e2e     the same through the public C-ABI with page-locked HOST buffers: read H2D and result D2H inside the timed region
  roofline        k_search on this config: L2-bound (index is L2-resident) against the L2 gather peak measured in this run
  human_scale /   BASELINE config[2] at FULL size (3.1 Gbp genome, spliced pairs) in the same run: value, e2e and the
  roofline_hbm    HBM-bound roofline of k_search against MEASURED_PEAKS.json
  cpu_baseline    the stock reference binary on the box's cores; cpu_baseline_hotpath: its per-read functions only (no parse / format / IO)
  fastq_to_sam    the tool: dart_b200_map vs dart_ref on the same FASTQ files
  host_link       what the host side of PCIe gives all ranks at once
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
CONTEXTS = int(os.environ.get("DART_BENCH_CONTEXTS", 4))   # batches in flight per GPU (one context each, one host thread for all)
PARTS = int(os.environ.get("DART_BENCH_PARTS", 2))         # sub-batches a step's batch is cut into
NW_OPS_PER_CELL = 26   # SASS instructions per cell in k_nw_thread's inner loop (4 cells per iteration; listing: profiles/r02_nw_thread_sass.txt)
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


def hotpath_baseline(g, idx, cores):
    """The reference's hot-path calls ONLY (IdentifySeedPairs ... EvaluateMAPQ, src/Mapping.cpp:600-639) over pre-parsed
    reads on all host cores, through oracle/_ref/libdartref.so: no FASTQ parsing, no SAM text, no file IO — the like-for-like
    denominator of `value` / `e2e` next to the stock binary's number."""
    from oracle import pyoracle as po
    sp = 100_000
    m1, m2 = make_pairs(g, sp, 0)
    b = as_batch(m1, m2)
    R = po.Reference(idx)
    R.set_params(max_mismatch=int(MIS) if MIS else 0, pair_end=1, **({"multi_hit": 1, "max_dup": 10000, "all_sj": 1} if WORKLOAD == "c5" else {}))
    R.hotpath(b.bases, b.offsets, 1, cores)               # warm-up (page in the index)
    sec = R.hotpath(b.bases, b.offsets, 1, cores)
    return {"value": b.n / sec, "unit": "reads/s", "cores": cores, "kind": "reference",
            "sample": f"{sp} pairs of the same workload, the reference's own per-read functions (libdartref.so) on {cores} threads, "
                      "reads pre-parsed in memory, no output formatting"}


def tool_level(g, idx, cores):
    """FASTQ files in, SAM file out: dart_b200_map (reader thread + GPU-side FASTQ parse, mapping and SAM text + pwrite pool) next to
    the stock reference binary on the SAME files (tmpfs when the box has one), index load excluded on both sides.  This is the
    number a user of the tool sees; it is bounded by file IO, not by the GPU."""
    import re
    from dart_b200 import synth
    import shutil
    pairs = int(os.environ.get("DART_BENCH_TOOL_PAIRS", 4_000_000))
    need = pairs * 1200                       # FASTQ + SAM + the reference's sample, with room to spare
    d = os.path.join(WORK, "tool")
    if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > 2 * need:
        d = "/dev/shm/dart_b200_tool"         # tmpfs: the number then measures the tool, not the disk
    os.makedirs(d, exist_ok=True)
    if shutil.disk_usage(d).free < need:
        return {"skipped": f"not enough free space under {d} for {pairs} pairs"}
    r1, r2 = os.path.join(d, f"t{pairs}_1.fq"), os.path.join(d, f"t{pairs}_2.fq")
    if not (os.path.exists(r1) and os.path.exists(r2)):
        with open(r1, "wb") as f1, open(r2, "wb") as f2:
            done = 0
            while done < pairs:
                k = min(1_000_000, pairs - done)
                m1, m2 = make_pairs(g, k, 100 + done // 1_000_000)
                f1.write(synth.fastq_bytes(m1, 1, first_id=done)); f2.write(synth.fastq_bytes(m2, 2, first_id=done))
                done += k
    tool = os.path.join(ROOT, "dart_b200", "dart_b200_map")
    best = None
    for _ in range(3):
        p = subprocess.run([tool, "-i", idx, "-f", r1, "-f2", r2, "-o", os.path.join(d, "gpu.sam"), "-j", os.path.join(d, "gpu.junc")],
                           capture_output=True, text=True, check=True)
        sec = float(re.search(r"processed in ([0-9.]+) seconds", p.stdout).group(1))
        best = sec if best is None else min(best, sec)
    # the reference on a quarter of the same files (it is ~5x slower), index-load time subtracted
    rp = min(pairs, 500_000)
    q1, q2 = os.path.join(d, "q_1.fq"), os.path.join(d, "q_2.fq")
    for src, dst in ((r1, q1), (r2, q2)):
        with open(src, "rb") as f, open(dst, "wb") as o:
            for _ in range(4 * rp):
                o.write(f.readline())
    e1, e2 = os.path.join(d, "one_1.fq"), os.path.join(d, "one_2.fq")
    for src, dst in ((r1, e1), (r2, e2)):
        with open(src, "rb") as f, open(dst, "wb") as o:
            for _ in range(4):
                o.write(f.readline())
    load = min(time_reference(idx, e1, e2, 2, cores, []) for _ in range(2))
    t_ref = max(time_reference(idx, q1, q2, 2 * rp, cores, []) - load, 1e-6)
    out = {"gpu_reads_per_s": 2 * pairs / best, "gpu_seconds": best, "pairs": pairs,
           "fastq_bytes": os.path.getsize(r1) + os.path.getsize(r2), "sam_bytes": os.path.getsize(os.path.join(d, "gpu.sam")),
           "reference_reads_per_s": 2 * rp / t_ref, "reference_pairs": rp, "reference_threads": cores,
           "what": "dart_b200_map vs dart_ref -t <cores> on the same FASTQ files, FASTQ -> SAM + junctions.tab, index load excluded, best of 3 (GPU) / one run (reference)"}
    for f in (os.path.join(d, "gpu.sam"), q1, q2, r1, r2):      # leave nothing behind (tmpfs is memory)
        try:
            os.remove(f)
        except OSError:
            pass
    return out


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


class Lanes:
    """K contexts of one GPU driven by ONE host thread: dartgpu_submit on a context, dartgpu_wait when its next batch is due,
    so K batches are in flight and the copies / kernels of consecutive batches overlap.  The thread sleeps in dartgpu_wait.
    A step's batch is cut into PARTS sub-batches; sub-batch j of step s goes to context (s * PARTS + j) mod K."""

    def __init__(self, mappers, subs):
        self.mappers, self.subs = mappers, subs          # subs[k] = the sub-batch context k maps (k mod PARTS)

    def upload(self):
        for m, sb in zip(self.mappers, self.subs):
            m.upload_reads(sb)

    def results_on_device(self, on):
        for m in self.mappers:
            m.results_on_device(on)
        self.on_device = on

    def run(self, resident, steps):
        """`steps` passes over the batch.  No barrier between steps: a context is only waited for when it is needed again."""
        K = len(self.mappers)
        parts = PARTS
        busy = [False] * K
        how = "device" if getattr(self, "on_device", False) else False
        k = 0
        for _ in range(steps * parts):
            m = self.mappers[k]
            if busy[k]:
                m.wait(copy=how)
            m.submit(None if resident else self.subs[k])
            busy[k] = True
            k = (k + 1) % K
        for i in range(K):             # oldest first
            j = (k + i) % K
            if busy[j]:
                self.mappers[j].wait(copy=how)

    def stats(self):
        """Work and event times of one step: the last batch of every context, scaled from K batches to PARTS."""
        sts = [m.stats() for m in self.mappers]
        f = PARTS / len(sts)
        return {k: sum(x[k] for x in sts) * f for k in sts[0]}


def make_lanes(capi, idx, local, params, batch, contexts):
    from dart_b200.shard import shard_bounds
    assert contexts % PARTS == 0
    mappers = [capi.Mapper(idx, device=local, **params) for _ in range(contexts)]
    bounds = shard_bounds(batch.n, PARTS, True)
    parts = []
    for a, b in zip(bounds[:-1], bounds[1:]):
        off = batch.offsets[a:b + 1]
        parts.append((batch.bases[off[0]:off[-1]], off - off[0]))
    # every context gets its own page-locked copy of its sub-batch (e2e: inputs in pinned host memory)
    subs = [capi.ReadBatch(parts[k % PARTS][0].copy(), parts[k % PARTS][1].copy()).pin() for k in range(contexts)]
    return Lanes(mappers, subs)


def search_roofline(M, batch, reps, peaks, l2_peak=None):
    """k_search over the whole batch on one context: roofline of the dominant kernel.
    Work: the device's counters from the stage entry point (dartgpu_seed_and_cluster_resident runs the kernel WITH its work
    counters): algorithmic bytes = SURVEY.md 8d (64 B per BWA block the reference's steps touch + packed read + 16 B per record),
    requested bytes = the 32-byte sectors the kernel really loads.  Time: CUDA events around the kernel as the whole-path call
    launches it (without the counters, which cost ~10 % of its instructions), averaged over `reps` calls."""
    M.upload_reads(batch)
    M.seed_resident()
    sk = M.stats()
    ms = []
    M.map_reads(batch, True, False)
    for _ in range(reps):
        M.map_reads(batch, True, False)
        ms.append(M.stats()["ms_search"])
    kms = float(np.mean(ms))
    sk["ms_search_with_counters"] = sk["ms_search"]
    alg = 64 * sk["ext_blocks"] + (sk["read_bases"] + 3) // 4 + 16 * sk["seeds"]
    req = 32 * sk["search_sector_loads"] + (sk["read_bases"] + 1) // 2 + 16 * sk["seeds"]   # packed read view: 8 B per 16 bases
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    out = {"kernel": "k_search (FM-index forward extension)", "kernel_ms": kms, "reads_in_launch": batch.n,
           "algorithmic_bytes_per_launch": int(alg), "requested_bytes_per_launch": int(req),
           "algorithmic_gbs": alg / (kms * 1e-3) / 1e9, "requested_gbs": req / (kms * 1e-3) / 1e9,
           "hbm_peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
           "kernel_ms_with_work_counters": sk["ms_search_with_counters"]}
    if l2_peak:
        out["l2_peak_gbs"] = l2_peak / 1e9
    return out, sk


def human_leg(args, rank, world, local, capi, barrier, allmax, peaks):
    """BASELINE config[2] at FULL size (3.1 Gbp genome in 24 contigs with gene models, spliced pairs 2x101): the
    HBM-bound case of the path, measured in the same run (extra keys of the bench line).  Index staged by the GPU builder
    (byte-identical files, tests/test_index_build.py); untimed preparation."""
    import psutil
    from dart_b200 import synth
    need = world * 14e9 + 8e9
    if psutil.virtual_memory().available < need:
        return {"skipped": f"host memory: {psutil.virtual_memory().available / 1e9:.0f} GB available, {need / 1e9:.0f} GB wanted for {world} ranks"}
    t0 = time.perf_counter()
    scale = float(os.environ.get("DART_BENCH_HUMAN_SCALE", "1.0"))
    g = synth.config_genome(3, scale)
    idx = os.path.join(WORK, f"idx_human_{scale}")
    if rank == 0 and not all(os.path.exists(idx + e) for e in (".bwt", ".sa", ".pac", ".ann", ".amb")):
        capi.index_build(g, idx + ".tmp", device=local)
        for e in (".bwt", ".sa", ".pac", ".ann", ".amb"):
            os.replace(idx + ".tmp" + e, idx + e)
    barrier()
    pairs = int(os.environ.get("DART_BENCH_HUMAN_PAIRS", PAIRS_PER_GPU))
    m1, m2 = synth.simulate_pairs(g, pairs, READ_LEN, 0.01, seed=2003 + rank, spliced=True, frag_min=202, frag_max=500)
    batch = as_batch(m1, m2)
    del g, m1, m2
    params = dict(pair_end=1, host_threads=max(1, (os.cpu_count() or 1) // world))
    lanes = make_lanes(capi, idx, local, params, batch, CONTEXTS)
    prep_s = time.perf_counter() - t0
    steps, warm = args.steps, max(3, args.warmup)
    lanes.upload()
    lanes.results_on_device(True)
    lanes.run(True, warm)
    ms_res = allmax(timed_region(barrier, lambda: lanes.run(True, steps)))
    st = lanes.stats()
    lanes.results_on_device(False)
    lanes.run(True, warm)
    ms_res_copy = allmax(timed_region(barrier, lambda: lanes.run(True, steps)))
    lanes.run(False, warm)
    ms_e2e = allmax(timed_region(barrier, lambda: lanes.run(False, steps)))
    st_e2e = lanes.stats()
    out = {"workload": f"BASELINE config[2] at x{scale} of full size: {int(3.1e9 * scale / 1e6)} Mbp genome in 24 contigs with gene models, "
                       f"spliced pairs 2x101 bp, 1% substitutions; {pairs} pairs per GPU and step",
           "value": world * batch.n * steps / (ms_res * 1e-3), "unit": "reads/s", "ms_per_step": ms_res / steps,
           "value_with_result_copy": world * batch.n * steps / (ms_res_copy * 1e-3),
           "e2e": {"value": world * batch.n * steps / (ms_e2e * 1e-3), "unit": "reads/s",
                   "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"])},
           "n_gpus": world, "steps": steps, "warmup": warm, "prep_seconds_untimed": prep_s,
           "kernels_ms_per_step": {k: st[k] for k in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_report", "ms_d2h")}}
    if rank == 0:
        M = lanes.mappers[0]
        rf, sk = search_roofline(M, batch, max(3, steps), peaks)
        ach, peak = rf["algorithmic_gbs"], rf["hbm_peak_gbs"]
        traffic = None
        try:   # DRAM bytes of the committed ncu capture of the same kernel on the same workload, scaled to this launch
            t = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_ncu.json")))["config2_fullsize"]
            traffic = int(t["dram_bytes_per_launch"] * batch.n / t["reads_in_launch"]) if scale == 1.0 else None
        except Exception:
            pass
        out["roofline_hbm"] = {"bound": "hbm", "kernel": rf["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                               "traffic": traffic, "kernel_ms": rf["kernel_ms"], "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"],
                               "requested_bytes_per_launch": rf["requested_bytes_per_launch"], "peak_source": rf["peak_source"],
                               "note": "3.1 GB Occ table + 1 GB start table: random sector gathers from HBM (L2 hit ~6 %)"}
    for m in lanes.mappers:
        m.close()
    return out


def host_link_probe(torch, barrier, allmax, world):
    """What the host side of PCIe gives ALL ranks at once: every rank copies 256 MB page-locked <-> device at the same time
    (after a barrier), CUDA-event timed, slowest rank reported.  `e2e` moves ~168 bytes per read over this link (ASCII bases in,
    records out); at N GPUs x 0.3-0.4 G reads/s per GPU that is the shared resource of a path without any collective."""
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    out = {}
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True)),
                     ("both", None)):
        if fn is None:
            s2 = torch.cuda.Stream()
            h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")

            def fn():
                d.copy_(h, non_blocking=True)
                with torch.cuda.stream(s2):
                    h2.copy_(d2, non_blocking=True)
        fn(); torch.cuda.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            fn()
        if name == "both":
            torch.cuda.current_stream().wait_stream(s2)
        e1.record(); torch.cuda.synchronize()
        ms = allmax(e0.elapsed_time(e1))
        out[name + "_gbs_per_gpu"] = (2 if name == "both" else 1) * 4 * n / (ms * 1e-3) / 1e9
    out["ranks_copying_at_once"] = world
    out["aggregate_both_gbs"] = out["both_gbs_per_gpu"] * world
    return out


def timed_region(barrier, fn):
    import torch
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


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

    # ONE host thread per GPU (this one) drives CONTEXTS contexts through dartgpu_submit / dartgpu_wait and sleeps while it
    # waits: the number of batches in flight is no longer tied to the host's cores (round 1: one spinning thread per context,
    # 32 spinners on a 32-core box at 8 GPUs, weak-scaling efficiency 0.58).
    import torch
    import torch.distributed as dist
    from dart_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(ms):
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    numa_bound = capi.bind_host_thread(local)       # page-locked buffers and the driving thread next to this rank's GPU
    if rank == 0:
        g, idx = prepare_genome()
    barrier()
    if rank != 0:
        g, idx = prepare_genome()
    link = host_link_probe(torch, barrier, allmax, world)
    params = dict(pair_end=1)
    if MIS:
        params["max_mismatch"] = int(MIS)
    if WORKLOAD == "c5":
        params.update(multi_hit=1, max_dup=10000, all_sj=1)
    cores = os.cpu_count() or 1
    params["host_threads"] = max(1, cores // world)      # OpenMP threads of the pageable-buffer staging only
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    batch = as_batch(*make_pairs(g, PAIRS_PER_GPU, rank))
    n_reads = batch.n
    lanes = make_lanes(capi, idx, local, params, batch, CONTEXTS)
    M = lanes.mappers[0]

    # ---- device-resident arm: reads AND records stay in HBM, nothing crosses PCIe in the timed region ----
    # every arm warms EVERY context up (a step uses PARTS of the CONTEXTS contexts): first use allocates page-locked result buffers
    warm_steps = max(args.warmup, -(-CONTEXTS // PARTS))
    lanes.upload()
    lanes.results_on_device(True)
    lanes.run(True, warm_steps)
    sampler = ClockSampler(local); sampler.start()
    ms_res = allmax(timed_region(barrier, lambda: lanes.run(True, args.steps)))
    st = lanes.stats()
    # ---- the same with the records copied to page-locked host memory every step (round 1's definition of `value`) ----
    lanes.results_on_device(False)
    lanes.run(True, warm_steps)
    ms_res_copy = allmax(timed_region(barrier, lambda: lanes.run(True, args.steps)))
    st_copy = lanes.stats()
    # ---- end-to-end arm (host buffers in, host results out) ----
    lanes.run(False, warm_steps)
    ms_e2e = allmax(timed_region(barrier, lambda: lanes.run(False, args.steps)))
    st_e2e = lanes.stats()
    clocks = sampler.result()

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    line = None
    if rank == 0:
        # one context alone on the GPU over the whole batch: per-kernel times without the other contexts' kernels in between
        # (the per-step sums above are CUDA-event intervals on streams that share the device)
        M.upload_reads(batch)
        M.map_reads(batch, True, False)
        M.map_reads(batch, True, False)
        alone = M.stats()
        int32_peak = M.int32_peak()
        l2_peak = M.l2_peak(16 << 20)
        rf, sk = search_roofline(M, batch, max(3, args.steps), peaks, l2_peak)
        l2_resident = WORKLOAD == "c2" or (WORKLOAD == "c5" and SCALE <= 2)
        if l2_resident:
            # config[1]'s 4.6 MB Occ table + 16 MB start table live in L2 (ncu: L2 hit 92 %, DRAM 3 % of peak): the roof of the
            # kernel is the L2's random-sector gather rate, measured on this GPU by dartgpu_measure_l2_peak, against the
            # bytes the kernel really requests (32 B per sector load, counted on the device)
            ach, peak = rf["requested_gbs"], l2_peak / 1e9
            dram = None
            try:   # DRAM bytes of the committed ncu capture of the same kernel on the same workload, scaled to this launch
                t = json.load(open(os.path.join(ROOT, "profiles", "search_kernel_ncu.json")))["config1"]
                dram = int(t["dram_bytes_per_launch"] * batch.n / t["reads_in_launch"]) if WORKLOAD == "c2" else None
            except Exception:
                pass
            roof = {"bound": "l2", "kernel": rf["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": dram,
                    "traffic_note": "dram__bytes_read + dram__bytes_write of the ncu capture: 5 % of the requested bytes reach HBM, the rest is served by L2",
                    "peak_source": "dartgpu_measure_l2_peak: random 32-byte sector gathers over a 16 MB L2-resident table, all SMs, CUDA events (this run)",
                    "requested_bytes_per_launch": rf["requested_bytes_per_launch"], "kernel_ms": rf["kernel_ms"],
                    "reference_algorithmic": {"bytes_per_launch": rf["algorithmic_bytes_per_launch"], "gbs": rf["algorithmic_gbs"],
                                              "frac_of_hbm_peak": rf["algorithmic_gbs"] / rf["hbm_peak_gbs"],
                                              "note": "SURVEY 8d bytes (64 B per BWA block of the reference's steps, incl. the steps the start table skips); "
                                                      "these bytes never reach HBM on this index, so this is not a roofline fraction"},
                    "note": "index is L2-resident: bound = L2 sector-gather rate (and instruction issue), not HBM; the HBM-bound case is `human_scale.roofline_hbm`"}
        else:
            ach, peak = rf["algorithmic_gbs"], rf["hbm_peak_gbs"]
            roof = {"bound": "hbm", "kernel": rf["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                    "peak_source": rf["peak_source"], "algorithmic_bytes_per_launch": rf["algorithmic_bytes_per_launch"], "kernel_ms": rf["kernel_ms"],
                    "requested_bytes_per_launch": rf["requested_bytes_per_launch"], "note": "Occ table larger than L2: sector gathers from HBM"}
        line = {
            "metric": "reads/sec mapped", "value": world * n_reads * args.steps / (ms_res * 1e-3), "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64 (2-bit bases, 64-bit FM intervals, int32 NW)",
            "data": "synthetic", "config": workload_config(), "clocks": clocks,
            "e2e": {"value": world * n_reads * args.steps / (ms_e2e * 1e-3), "unit": "reads/s",
                    "h2d_bytes_per_step": int(st_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(st_e2e["d2h_bytes"])},
            "value_definition": "reads and result records resident in HBM: no PCIe traffic inside the timed region (dartgpu_set_result_location); "
                                "value_with_result_copy = the same with the records copied to page-locked host memory every step (round 1's `value`); "
                                "e2e = host buffers both ways",
            "value_with_result_copy": {"value": world * n_reads * args.steps / (ms_res_copy * 1e-3), "unit": "reads/s", "ms_per_step": ms_res_copy / args.steps,
                                       "d2h_bytes_per_step": int(st_copy["d2h_bytes"])},
            "gpu_launches": int(st["kernel_launches"]) * args.steps, "contexts_per_gpu": CONTEXTS, "batches_per_step": PARTS, "host_threads_per_gpu": 1,
            "host_sync": os.environ.get("DARTGPU_SYNC", "block (one sleeping wait per batch)"), "host_cores": os.cpu_count(),
            "host_thread_bound_to_gpu_numa_node": bool(numa_bound), "host_link": link,
            "roofline": roof,
            "kernels_ms_per_step": {k: st[k] for k in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_report", "ms_d2h", "ms_host")},
            "work_per_step": {**{k: int(sk[k]) for k in ("ext_steps", "ext_blocks", "search_sector_loads")},
                              **{k: int(st[k]) for k in ("lf_steps", "hits", "seeds", "nw_jobs", "nw_cells", "kmer_jobs", "kmer_window_bases")}},
            "kernels_ms_one_context_alone": {k: alone[k] for k in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_report", "ms_d2h")},
            "nw_gcups": (alone["nw_cells"] / (alone["ms_nw"] * 1e-3) / 1e9) if alone["ms_nw"] > 0 else None,
            "seed_gbs": rf["algorithmic_gbs"],
        }
        if line["nw_gcups"]:
            ops_per_cell = NW_OPS_PER_CELL
            line["nw_roofline"] = {"bound": "int32", "achieved_gcups": line["nw_gcups"], "int32_ops_per_s_measured": int32_peak,
                                   "ops_per_cell": ops_per_cell, "peak_gcups": int32_peak / ops_per_cell / 1e9,
                                   "frac": line["nw_gcups"] * 1e9 * ops_per_cell / int32_peak,
                                   "note": "whole NW stage of one context alone (shape-class sort + both kernels + tracebacks), cells = sum of m*n; "
                                           "ops_per_cell from the SASS listing profiles/r02_nw_thread_sass.txt"}
    for m in lanes.mappers:
        m.close()
    del lanes, batch
    # ---- the HBM-bound configuration in the same run ----
    human = None
    if os.environ.get("DART_BENCH_HUMAN", "1") != "0" and WORKLOAD == "c2":
        try:
            human = human_leg(args, rank, world, local, capi, barrier, allmax, peaks)
        except Exception as ex:   # reported, never a reason to lose the headline measurement
            human = {"failed": repr(ex)}
    if rank == 0:
        if human is not None:
            line["human_scale"] = human
            if "roofline_hbm" in human:
                line["roofline_hbm"] = human["roofline_hbm"]
        if world == 1:
            try:
                extra = (["-mis", MIS] if MIS else []) + (["-m", "-max_dup", "10000", "-all_sj"] if WORKLOAD == "c5" else [])
                sp = min(REF_SAMPLE_PAIRS, 100_000 if WORKLOAD != "c4" else 40_000)
                r1, r2, e1, e2 = reference_setup(g, idx, sp)
                load = min(time_reference(idx, e1, e2, 2, cores, extra) for _ in range(2))
                t = max(time_reference(idx, r1, r2, 2 * sp, cores, extra) - load, 1e-6)
                line["cpu_baseline"] = {"value": 2 * sp / t, "unit": "reads/s", "cores": cores, "kind": "reference",
                                        "sample": f"{sp} pairs of the same workload, oracle/_ref/dart_ref -t {cores} (FASTQ parse + map + SAM text + write), index-load time subtracted"}
            except Exception as ex:  # the baseline is reported, never a reason to lose the measurement
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": f"failed: {ex}"}
            try:
                line["cpu_baseline_hotpath"] = hotpath_baseline(g, idx, cores)
            except Exception as ex:
                line["cpu_baseline_hotpath"] = {"value": None, "sample": f"failed: {ex}"}
            if WORKLOAD == "c2" and os.environ.get("DART_BENCH_TOOL", "1") != "0":
                try:
                    line["fastq_to_sam"] = tool_level(g, idx, cores)
                except Exception as ex:
                    line["fastq_to_sam"] = {"failed": repr(ex)}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
