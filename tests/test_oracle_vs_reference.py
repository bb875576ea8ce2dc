"""Pin the CPU oracle against the UNMODIFIED reference (oracle/_ref/libdartref.so), stage by stage, on freshly
generated workloads. One index per process (the reference keeps it in globals), so everything runs on c3."""
import os
import random

import numpy as np
import pytest

from conftest import need_ref, read_fastq_seqs, workload
from oracle import pyoracle as po


@pytest.fixture(scope="module")
def pair():
    need_ref()
    w = workload("c3")
    O, R = po.Oracle(w["idx"]), po.Reference(w["idx"])
    yield w, O, R
    O.close()


def test_rank_and_locate(pair):
    _, O, R = pair
    rng = np.random.default_rng(7)
    for k in rng.integers(0, 2 * O.G, 1500):
        k = int(k)
        l = min(2 * O.G, k + int(rng.integers(0, 300)))
        a, b = R.occ2(k, l)
        assert (O.rank4(k) == a).all() and (O.rank4(l) == b).all()
        assert O.locate(k) == R.sa(k)


def test_search_seeds_candidates(pair):
    w, O, R = pair
    reads = read_fastq_seqs(w["r1"], 800)
    reads[1] = reads[1][:30] + b"N" + reads[1][31:]
    reads[2] = reads[2][:50] + b"acgtn" + reads[2][55:]
    for s in reads:
        c = po.encode(s)
        for start in (0, 7, 33):
            if c[start] <= 3:
                lo, fo, xo = O.search(c, start, len(c))
                lr, fr, xr = R.search(c, start, len(c))
                assert (lo, fo) == (lr, fr) and (xo == xr).all()
        r1, g1, l1 = O.seeds(c)
        r2, g2, l2 = R.seeds(c)
        assert (r1 == r2).all() and (g1 == g2).all() and (l1 == l2).all()
        cb, cc, cs = O.cluster(len(c), r1, g1, l1)
        rs, rp, rn, sr, sg, sl = R.candidates(c)
        assert (cs == rs).all() and (cc == rn).all()
        flat = np.concatenate([np.arange(b, b + n) for b, n in zip(cb, cc)]) if len(cb) else np.zeros(0, int)
        assert (r1[flat] == sr).all() and (g1[flat] == sg).all()


def test_mirrored_locate_gives_the_reference_hits():
    """The CUDA path locates each match from the reverse-complement half of the bi-interval and mirrors the coordinate
    (p -> 2G - p - len).  The text is its own reverse complement, so the hit SET must equal the reference's LocArr —
    checked here against the reference itself on a repeat-rich index with -max_dup 10000, every search start of many reads."""
    import subprocess, sys, json
    from conftest import ROOT
    code = """
import sys, json, numpy as np
sys.path.insert(0, %r)
from oracle import pyoracle as po
from conftest import workload, read_fastq_seqs
w = workload("c5")
O, R = po.Oracle(w["idx"]), po.Reference(w["idx"])
R.set_params(max_dup=10000)
n = multi = 0
for s in read_fastq_seqs(w["r1"], 400) + [b"ACACACACACACACACACACACACACACACACACACACAC", b"A" * 60]:
    c = po.encode(s)
    for start in range(0, max(1, len(c) - 14), 3):
        if c[start] > 3: continue
        lr, fr, xr = R.search(c, start, len(c))
        lm, fm, xm = O.search_mirrored(c, start, len(c), max_dup=10000)
        assert (lr, fr) == (lm, fm), (s, start)
        assert sorted(xr.tolist()) == sorted(xm.tolist()), (s, start)
        n += 1; multi += fr > 1
print(json.dumps([n, multi]))
""" % os.path.join(ROOT, "tests")
    out = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, cwd=os.path.join(ROOT, "tests"),
                         env=dict(os.environ, PYTHONPATH=ROOT)).stdout.decode().strip().splitlines()[-1]
    n, multi = json.loads(out)
    assert n > 5000 and multi > 100


def _rnd(rng, n):
    return bytes(rng.choice(b"ACGT") for _ in range(n))


def _mut(rng, s, p):
    out = bytearray()
    for ch in s:
        x = rng.random()
        if x < p / 3:
            continue
        if x < 2 * p / 3:
            out.append(rng.choice(b"ACGT"))
        if x < p:
            out.append(rng.choice(b"ACGT"))
            continue
        out.append(ch)
    return bytes(out) or b"A"


def test_nw_random(pair):
    _, O, R = pair
    rng = random.Random(5)
    for _ in range(1500):
        s1 = _rnd(rng, rng.randint(1, 70))
        s2 = _mut(rng, s1, 0.15) if rng.random() < 0.7 else _rnd(rng, rng.randint(1, 70))
        if rng.random() < 0.05:
            s1 = s1[:len(s1) // 2] + b"N" + s1[len(s1) // 2 + 1:]
        assert po.ops_to_strings(s1, s2, O.nw(s1, s2)) == R.nw(s1, s2)


def test_kmer_random(pair):
    _, O, R = pair
    rng = random.Random(6)
    for _ in range(600):
        f1 = _rnd(rng, rng.randint(8, 120))
        mode = rng.random()
        if mode < 0.5:
            f2 = bytearray(_rnd(rng, rng.randint(20, 3000)))
            seg = _mut(rng, f1, 0.05)
            p = rng.randint(0, max(0, len(f2) - len(f1)))
            f2[p:p + len(seg)] = seg
            f2 = bytes(f2)
        elif mode < 0.8:  # low complexity: many equal-PosDiff runs, exercises the carried counter
            unit = _rnd(rng, rng.randint(1, 5))
            f1 = (unit * 60)[:len(f1)]
            f2 = (unit * 300)[:rng.randint(30, 600)]
        else:
            f2 = _rnd(rng, rng.randint(8, 2000))
        if rng.random() < 0.1:
            q = rng.randrange(len(f1))
            f1 = f1[:q] + rng.choice([b"N", b"n", b"R"]) + f1[q + 1:]
        assert O.kmer_pair(f1, f2) == R.kmer_pair(f1, f2)


def test_gapped_partition_random(pair):
    _, O, R = pair
    rng = random.Random(8)
    R.set_params(max_mismatch=5)
    G = O.G
    for _ in range(300):
        gl = rng.randint(1000, G - 2000)
        left = R.ref_chars(gl, 40)
        gap = _mut(rng, R.ref_chars(gl + 40, rng.randint(1, 30)), rng.choice([0, 0.1, 0.3]))
        jump = rng.randint(30, 400)
        right = R.ref_chars(gl + 40 + len(gap) + jump, 40)
        seq = left + gap + right
        rg = len(gap)
        args = (seq, rg, 0, 40, gl, 40, 40 + rg, gl + 40 + rg + jump)
        assert O.gapped_partition(*args, 5) == R.gapped_partition(*args)
    R.set_params(max_mismatch=0)
