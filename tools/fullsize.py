#!/usr/bin/env python
"""BASELINE.json's configs at FULL size on one GPU box, over any number of GPUs: config[2] (3.1 Gbp genome with gene
models, spliced 2x101 pairs), config[3] (same genome, 2x250 pairs with 3 % substitutions + indels, -mis 10) and config[4]
(repeat-rich 4.6 Mbp genome, -m -max_dup 10000 -all_sj).

`--devices 0,1,..` maps over several GPUs (contiguous blocks per GPU in turn, results in input order, junction counts summed).
`--record FILE` stores the reference's fingerprints (multiset hash of the SAM records, junctions.tab digest) so that a later
run on a bigger, more expensive box (`--expect FILE --skip-ref`) checks the N-GPU output against them without paying for the
CPU reference again: the inputs are seeded and identical on every box.

The reference's own index builder needs hours for 3.1 Gbp, so the index comes from dartgpu_index_build (byte-identical
to bwt_index wherever both run: tests/test_index_build.py).  Parity: the canonical reference (oracle/_ref/dart_canon,
all host cores) and dart_b200_map run on the same files; junctions.tab must be identical and the SAM records must be the
same multiset (tools/samhash.cpp — the reference writes chunks in completion order).

usage: fullsize.py [--scale 1.0] [--pairs2 20000000] [--pairs3 200000] [--dir /dev/shm/dart_full] [--skip-ref]
Writes a JSON summary to gpurun_out/fullsize.json.
"""
import argparse
import hashlib
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
            seed = (2003 if cfg == 3 else 2004 if cfg == 4 else 2005) * 1000 + done // chunk
            if cfg == 3:
                m1, m2 = synth.simulate_pairs(g, k, 101, 0.01, seed=seed, spliced=True, frag_min=202, frag_max=500)
            elif cfg == 5:
                m1, m2 = synth.simulate_pairs(g, k, 101, 0.01, seed=seed)
            else:
                m1, m2 = synth.simulate_pairs(g, k, 250, 0.03, seed=seed, frag_mean=600, frag_sd=50, frag_min=500, frag_max=900,
                                              p_ins=0.002, p_del=0.002)
            f1.write(synth.fastq_bytes(m1, 1, first_id=done))
            f2.write(synth.fastq_bytes(m2, 2, first_id=done))
            done += k


def samhash(exe, sam, dump=None):
    out = subprocess.run([exe, sam] + ([dump] if dump else []), check=True, capture_output=True, text=True).stdout.split()
    return int(out[0]), out[1], out[2]


def run_config(tag, g, idx, d, cfg, n_pairs, flags, skip_ref, exe_hash, summary, devices=None, expect=None, record=None):
    r1, r2 = os.path.join(d, f"{tag}_1.fq"), os.path.join(d, f"{tag}_2.fq")
    t = time.time()
    write_pairs(g, cfg, n_pairs, r1, r2)
    log(tag, "reads written", n_pairs, "pairs in %.0f s" % (time.time() - t))
    res = {"pairs": n_pairs, "flags": flags}
    t = time.time()
    gsam, gj = os.path.join(d, tag + "_gpu.sam"), os.path.join(d, tag + "_gpu.junc")
    p = subprocess.run([os.path.join(ROOT, "dart_b200", "dart_b200_map"), "-i", idx, "-f", r1, "-f2", r2, "-o", gsam, "-j", gj] +
                       (["-devices", devices] if devices else []) + flags, capture_output=True, text=True)
    res["gpu_wall_s"] = time.time() - t
    res["devices"] = devices or "0"
    res["gpu_stdout"] = p.stdout.strip().split("\n")[:8]
    log(tag, "GPU path rc", p.returncode, "%.0f s" % res["gpu_wall_s"], p.stdout.strip().split("\n")[0] if p.stdout else p.stderr[-500:])
    if p.returncode != 0:
        res["error"] = p.stderr[-2000:]
        summary[tag] = res
        return
    res["gpu_hash"] = samhash(exe_hash, gsam, os.path.join(d, tag + "_gpu.h64"))
    res["gpu_junctions_md5"] = hashlib.md5(open(gj, "rb").read()).hexdigest()
    res["junction_lines"] = sum(1 for _ in open(gj))
    if expect is not None and tag in expect:
        e = expect[tag]
        res["expected_from"] = e.get("recorded_on", "an earlier run of the reference on the same seeded inputs")
        res["ref_hash"] = e["ref_hash"]; res["ref_junctions_md5"] = e["ref_junctions_md5"]
        res["sam_identical_multiset"] = list(e["ref_hash"]) == list(res["gpu_hash"])
        res["junctions_identical"] = e["ref_junctions_md5"] == res["gpu_junctions_md5"]
        log(tag, "vs recorded reference: SAM multiset identical:", res["sam_identical_multiset"], " junctions identical:", res["junctions_identical"])
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
        res["ref_junctions_md5"] = hashlib.md5(open(rj, "rb").read()).hexdigest()
        if record is not None:
            record[tag] = {"ref_hash": list(res["ref_hash"]), "ref_junctions_md5": res["ref_junctions_md5"], "pairs": n_pairs, "flags": flags,
                           "ref_wall_s": res["ref_wall_s"], "ref_cores": cores, "recorded_on": time.strftime("%Y-%m-%d %H:%M:%S") + " (1-GPU box, dart_canon)"}
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
    ap.add_argument("--pairs5", type=int, default=0, help="config[4]: repeat-rich 4.6 Mbp genome at full size, -m -max_dup 10000 -all_sj")
    ap.add_argument("--dir", default="/dev/shm/dart_full")
    ap.add_argument("--skip-ref", action="store_true")
    ap.add_argument("--devices", default=None)
    ap.add_argument("--record", default=None)
    ap.add_argument("--expect", default=None)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "fullsize.json"))
    args = ap.parse_args()
    os.makedirs(args.dir, exist_ok=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    exe_hash = os.path.join(args.dir, "samhash")
    subprocess.run(["g++", "-O2", "-o", exe_hash, os.path.join(ROOT, "tools", "samhash.cpp")], check=True)
    expect = json.load(open(args.expect)) if args.expect else None
    record = {} if args.record else None
    summary = {"scale": args.scale, "devices": args.devices or "0", "host_cores": os.cpu_count()}
    kw = dict(devices=args.devices, expect=expect, record=record)

    def save():
        json.dump(summary, open(args.out, "w"), indent=1)
        if record is not None:
            json.dump(record, open(args.record, "w"), indent=1)

    if args.pairs5 > 0:
        t = time.time()
        g5 = synth.config_genome(5, 1.0)
        idx5 = os.path.join(args.dir, "idx5")
        capi.index_build(g5, idx5)
        log("config[4] genome", g5.total_len, "bp, index built in %.0f s" % (time.time() - t))
        run_config("config4", g5, idx5, args.dir, 5, args.pairs5, ["-m", "-max_dup", "10000", "-all_sj"], args.skip_ref, exe_hash, summary, **kw)
        save()
        del g5
    if args.pairs2 > 0 or args.pairs3 > 0:
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
        save()
        if args.pairs2 > 0:
            run_config("config2", g, idx, args.dir, 3, args.pairs2, [], args.skip_ref, exe_hash, summary, **kw)
            save()
        if args.pairs3 > 0:
            run_config("config3", g, idx, args.dir, 4, args.pairs3, ["-mis", "10"], args.skip_ref, exe_hash, summary, **kw)
            save()
    log("done")


if __name__ == "__main__":
    main()
