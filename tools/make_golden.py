#!/usr/bin/env python
"""Generate tests/golden/ from the UNMODIFIED reference (run in the build container, where /root/reference
exists and oracle/_ref has been built by `make -C oracle ref`).

The reference ships no tests or golden vectors for the mapping hot path, so the pins are
outputs of the reference itself on a small seeded workload:
  golden/idx.*            a 36 kbp, 2-contig genome with planted gene models, indexed by the reference's bwt_index
  golden/stage.json       per-read IdentifySeedPairs / GenerateAlignmentCandidate results, nw_alignment,
                          GenerateLongestSimplePairsFromFragmentPair and IdentifyBestGappedPartition cases
                          (via oracle/_ref/libdartref.so)
  golden/{se,pe}.sam/.junc  end-to-end output of oracle/_ref/dart_canon -t 1 (-mis 5)
"""
import json
import os
import random
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dart_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def main():
    os.makedirs(OUT, exist_ok=True)
    po.build(ref=True)
    g = synth.random_genome(36000, 2, seed=77)
    synth.add_gene_models(g, genes_per_mbp=400, seed=78, max_intron=3000)
    fa = os.path.join(OUT, "genome.fa")
    synth.write_fasta(fa, g)
    for e in (".bwt", ".sa", ".pac", ".ann", ".amb"):
        if os.path.exists(os.path.join(OUT, "idx" + e)):
            os.remove(os.path.join(OUT, "idx" + e))
    po.build_index(fa, os.path.join(OUT, "idx"))
    se = synth.simulate_single(g, 150, 100, 0.02, seed=79)
    synth.write_fastq(os.path.join(OUT, "se.fq"), se)
    m1, m2 = synth.simulate_pairs(g, 120, 101, 0.01, seed=80, spliced=True, frag_min=202, frag_max=400, frag_mean=260)
    synth.write_fastq(os.path.join(OUT, "pe1.fq"), m1, 1)
    synth.write_fastq(os.path.join(OUT, "pe2.fq"), m2, 2)
    ref = os.path.join(ROOT, "oracle", "_ref", "dart_canon")
    env = dict(os.environ, MALLOC_PERTURB_="255", GLIBC_TUNABLES="glibc.malloc.tcache_count=0")   # zero-filled heap: see tests/conftest.py canonical_env()
    subprocess.run([ref, "-i", OUT + "/idx", "-f", OUT + "/se.fq", "-t", "1", "-mis", "5", "-o", OUT + "/se.sam",
                    "-j", OUT + "/se.junc"], check=True, stdout=subprocess.DEVNULL, env=env)
    subprocess.run([ref, "-i", OUT + "/idx", "-f", OUT + "/pe1.fq", "-f2", OUT + "/pe2.fq", "-t", "1", "-mis", "5",
                    "-o", OUT + "/pe.sam", "-j", OUT + "/pe.junc"], check=True, stdout=subprocess.DEVNULL, env=env)

    R = po.Reference(OUT + "/idx")
    R.set_params(max_mismatch=5)
    A = "ACGT"
    stage = {"seeds": [], "nw": [], "kmer": [], "gapped": []}
    reads = ["".join(A[c] for c in r) for r in list(se[:60]) + list(m1[:40])]
    reads[3] = reads[3][:40] + "N" + reads[3][41:]
    reads[5] = reads[5][:17] + "NN" + reads[5][19:70] + "n" + reads[5][71:]
    for s in reads:
        codes = po.encode(s)
        r, gp, ln = R.seeds(codes)
        cs, cp, cn, sr, sg, sl = R.candidates(codes)
        stage["seeds"].append(dict(read=s, rpos=r.tolist(), gpos=gp.tolist(), len=ln.tolist(), cand_score=cs.tolist(),
                                   cand_posdiff=cp.tolist(), cand_nseeds=cn.tolist(), cand_seed_rpos=sr.tolist(),
                                   cand_seed_gpos=sg.tolist()))
    rnd = random.Random(81)

    def rs(n):
        return "".join(rnd.choice(A) for _ in range(n))

    def mut(s, p):
        o = []
        for ch in s:
            x = rnd.random()
            if x < p / 3:
                continue
            if x < 2 * p / 3:
                o.append(rnd.choice(A))
            if x < p:
                o.append(rnd.choice(A))
                continue
            o.append(ch)
        return "".join(o) or "A"

    G = R.G
    for _ in range(160):
        n = rnd.randint(1, 90)
        gpos = rnd.randint(0, 2 * G - n - 1)
        s2 = R.ref_chars(gpos, n).decode()
        s1 = mut(s2, rnd.choice([0.0, 0.05, 0.2, 0.5])) if rnd.random() < 0.8 else rs(rnd.randint(1, 70))
        if rnd.random() < 0.1:
            q = rnd.randrange(len(s1)); s1 = s1[:q] + "N" + s1[q + 1:]
        a, b = R.nw(s1.encode(), s2.encode())
        stage["nw"].append(dict(s1=s1, gpos=gpos, n=n, a=a.decode(), b=b.decode()))
    for _ in range(80):
        L2 = rnd.randint(30, 4000)
        gpos = rnd.randint(0, 2 * G - L2 - 1)
        win = R.ref_chars(gpos, L2).decode()
        L1 = rnd.randint(21, 100)
        if rnd.random() < 0.7 and L2 > L1 + 2:
            p = rnd.randint(0, L2 - L1 - 1)
            f1 = mut(win[p:p + L1], rnd.choice([0, 0.03, 0.1]))
        else:
            f1 = rs(L1)
        if rnd.random() < 0.15 and len(f1) > 10:
            q = rnd.randrange(len(f1)); f1 = f1[:q] + rnd.choice("NnR") + f1[q + 1:]
        stage["kmer"].append(dict(f1=f1, gpos=gpos, glen=L2, out=list(R.kmer_pair(f1.encode(), win.encode()))))
    # gapped partitions: a read made of two genome blocks separated by a small deletion in the read
    for _ in range(60):
        gl = rnd.randint(1000, G - 1000)
        left = R.ref_chars(gl, 40).decode()
        gap_len = rnd.randint(1, 25)
        gap = mut(R.ref_chars(gl + 40, gap_len).decode(), rnd.choice([0, 0.1, 0.3]))
        jump = rnd.randint(30, 400)
        right = R.ref_chars(gl + 40 + len(gap) + jump, 40).decode()
        seq = left + gap + right
        rg = len(gap)
        args = dict(seq=seq, rgaps=rg, l_rpos=0, l_rlen=40, l_gpos=gl, l_glen=40, r_rpos=40 + rg, r_gpos=gl + 40 + rg + jump)
        out = R.gapped_partition(seq.encode(), rg, 0, 40, gl, 40, 40 + rg, gl + 40 + rg + jump)
        stage["gapped"].append(dict(args=args, max_mismatch=5, out=list(out)))
    with open(os.path.join(OUT, "stage.json"), "w") as f:
        json.dump(stage, f)
    os.remove(fa)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
