#!/usr/bin/env python
"""BASELINE.json's human-sized configs at FULL size on one GPU box: config[2] (3.1 Gbp genome with gene models, spliced
2x101 pairs) and config[3] (same genome, 2x250 pairs with 3 % substitutions + indels, -mis 10).

The reference's own index builder needs hours for 3.1 Gbp, so the index comes from dartgpu_index_build (byte-identical
to bwt_index wherever both run: tests/test_index_build.py).  Parity: the canonical reference (oracle/_ref/dart_canon,
all host cores) and dart_b200_map run on the same files; junctions.tab must be identical and the SAM records must be the
same multiset (tools/samhash.cpp — the reference writes chunks in completion order).

usage: fullsize.py [--scale 1.0] [--pairs2 20000000] [--pairs3 200000] [--dir /dev/shm/dart_full] [--skip-ref]
Writes a JSON summary to gpurun_out/fullsize.json.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dart_b200 import capi, synth  # noqa: E402


def log(*a):
    print(time.strftime("%H:%M:%S"), *a, flush=True)


def write_pairs(g, cfg, n_pairs, r1, r2, chunk=500_000):
    """FASTQ files for n_pairs of the named config, generated in chunks (seeded per chunk)."""
    with open(r1, "wb") as f1, open(r2, "wb") as f2:
        done = 0
        while done < n_pairs:
            k = min(chunk, n_pairs - done)
            seed = (2003 if cfg == 3 else 2004) * 1000 + done // chunk
            if cfg == 3:
                m1, m2 = synth.simulate_pairs(g, k, 101, 0.01, seed=seed, spliced=True, frag_min=202, frag_max=500)
            else:
                m1, m2 = synth.simulate_pairs(g, k, 250, 0.03, seed=seed, frag_mean=600, frag_sd=50, frag_min=500, frag_max=900,
                                              p_ins=0.002, p_del=0.002)
            f1.write(synth.fastq_bytes(m1, 1, first_id=done))
            f2.write(synth.fastq_bytes(m2, 2, first_id=done))
            done += k


def samhash(exe, sam, dump=None):
    out = subprocess.run([exe, sam] + ([dump] if dump else []), check=True, capture_output=True, text=True).stdout.split()
    return int(out[0]), out[1], out[2]


def run_config(tag, g, idx, d, cfg, n_pairs, flags, skip_ref, exe_hash, summary):
    r1, r2 = os.path.join(d, f"{tag}_1.fq"), os.path.join(d, f"{tag}_2.fq")
    t = time.time()
    write_pairs(g, cfg, n_pairs, r1, r2)
    log(tag, "reads written", n_pairs, "pairs in %.0f s" % (time.time() - t))
    res = {"pairs": n_pairs, "flags": flags}
    t = time.time()
    gsam, gj = os.path.join(d, tag + "_gpu.sam"), os.path.join(d, tag + "_gpu.junc")
    p = subprocess.run([os.path.join(ROOT, "dart_b200", "dart_b200_map"), "-i", idx, "-f", r1, "-f2", r2, "-o", gsam, "-j", gj,
                        "-stats"] + flags, capture_output=True, text=True)
    res["gpu_wall_s"] = time.time() - t
    res["gpu_stdout"] = p.stdout.strip().split("\n")[:8]
    res["gpu_stats_tail"] = p.stderr.strip().split("\n")[-3:]
    log(tag, "GPU path rc", p.returncode, "%.0f s" % res["gpu_wall_s"], p.stdout.strip().split("\n")[0] if p.stdout else p.stderr[-500:])
    if p.returncode != 0:
        res["error"] = p.stderr[-2000:]
        summary[tag] = res
        return
    res["gpu_hash"] = samhash(exe_hash, gsam, os.path.join(d, tag + "_gpu.h64"))
    if not skip_ref:
        env = dict(os.environ, MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0")
        rsam, rj = os.path.join(d, tag + "_ref.sam"), os.path.join(d, tag + "_ref.junc")
        cores = os.cpu_count() or 1
        t = time.time()
        subprocess.run([os.path.join(ROOT, "oracle", "_ref", "dart_canon"), "-i", idx, "-f", r1, "-f2", r2, "-t", str(cores), "-o", rsam,
                        "-j", rj] + flags, check=True, stdout=subprocess.DEVNULL, env=env)
        res["ref_wall_s"] = time.time() - t
        res["ref_cores"] = cores
        log(tag, "reference done in %.0f s" % res["ref_wall_s"])
        res["ref_hash"] = samhash(exe_hash, rsam, os.path.join(d, tag + "_ref.h64"))
        res["sam_identical_multiset"] = res["ref_hash"] == res["gpu_hash"]
        res["junctions_identical"] = open(rj, "rb").read() == open(gj, "rb").read()
        res["junction_lines"] = sum(1 for _ in open(gj))
        if not res["sam_identical_multiset"]:
            a = np.fromfile(os.path.join(d, tag + "_gpu.h64"), dtype=np.uint64)
            b = np.fromfile(os.path.join(d, tag + "_ref.h64"), dtype=np.uint64)
            only_gpu = np.setdiff1d(a, b)
            only_ref = np.setdiff1d(b, a)
            res["records_only_gpu"], res["records_only_ref"] = int(len(only_gpu)), int(len(only_ref))
            # show a few differing records (ours are in input order: record i of the GPU file)
            want = set(np.nonzero(np.isin(a, only_gpu[:5]))[0].tolist())
            shown = []
            if want:
                i = 0
                with open(gsam) as f:
                    for line in f:
                        if line.startswith("@"):
                            continue
                        if i in want:
                            shown.append(line.rstrip("\n")[:400])
                        i += 1
                        if i > max(want):
                            break
            names = {s.split("\t")[0] for s in shown}
            ref_shown = []
            if names:
                out = subprocess.run(["grep", "-m", "20", "-F", "-e", *sum([["-e", n] for n in names], [])[1:], rsam], capture_output=True, text=True).stdout
                ref_shown = [l[:400] for l in out.split("\n") if l][:12]
            res["sample_gpu"], res["sample_ref"] = shown, ref_shown
        log(tag, "SAM multiset identical:", res["sam_identical_multiset"], " junctions identical:", res["junctions_identical"])
        for p_ in (rsam,):
            os.remove(p_)
    os.remove(gsam)
    os.remove(r1); os.remove(r2)
    summary[tag] = res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--pairs2", type=int, default=20_000_000)
    ap.add_argument("--pairs3", type=int, default=200_000)
    ap.add_argument("--dir", default="/dev/shm/dart_full")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fullsize.json"))
    args = ap.parse_args()
    os.makedirs(args.dir, exist_ok=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    exe_hash = os.path.join(args.dir, "samhash")
    subprocess.run(["g++", "-O2", "-o", exe_hash, os.path.join(ROOT, "tools", "samhash.cpp")], check=True)
    summary = {"scale": args.scale}
    t = time.time()
    g = synth.config_genome(3, args.scale)
    summary["genome_bp"] = g.total_len
    summary["genes"] = len(g.genes)
    log("genome", g.total_len, "bp,", len(g.genes), "genes in %.0f s" % (time.time() - t))
    idx = os.path.join(args.dir, "idx")
    t = time.time()
    capi.index_build(g, idx)
    summary["index_build_s"] = time.time() - t
    summary["index_bytes"] = {e: os.path.getsize(idx + e) for e in (".bwt", ".sa", ".pac")}
    log("index built on the GPU in %.0f s" % summary["index_build_s"], summary["index_bytes"])
    json.dump(summary, open(args.out, "w"), indent=1)
    if args.pairs2 > 0:
        run_config("config2", g, idx, args.dir, 3, args.pairs2, [], args.skip_ref, exe_hash, summary)
        json.dump(summary, open(args.out, "w"), indent=1)
    if args.pairs3 > 0:
        run_config("config3", g, idx, args.dir, 4, args.pairs3, ["-mis", "10"], args.skip_ref, exe_hash, summary)
        json.dump(summary, open(args.out, "w"), indent=1)
    log("done")


if __name__ == "__main__":
    main()
