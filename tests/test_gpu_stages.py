"""GPU parity, stage by stage, through the C-ABI (dart_b200/capi.py -> libdartgpu.so), against the CPU oracle
and the committed golden vectors. Bit-exact: every value is an integer."""
import os
import random

import numpy as np
import pytest

from conftest import GOLDEN, read_fastq_seqs, workload
from dart_b200 import capi
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
A = b"ACGT"


def _check_seeds(M, O, reads, max_dup=100):
    res = M.identify_seed_pairs(capi.ReadBatch.from_list(reads))
    so, co = res["seed_off"], res["cand_off"]
    assert len(so) == len(reads) + 1
    nseed = ncand = 0
    for i, s in enumerate(reads):
        c = po.encode(s)
        r, g, l = O.seeds(c, max_dup=max_dup)
        a, b = so[i], so[i + 1]
        assert res["seed_rpos"][a:b].tolist() == r.tolist(), (i, s)
        assert res["seed_gpos"][a:b].tolist() == g.tolist(), (i, s)
        assert res["seed_len"][a:b].tolist() == l.tolist(), (i, s)
        cb, cc, cs = O.cluster(len(c), r, g, l)
        a2, b2 = co[i], co[i + 1]
        assert res["cand_begin"][a2:b2].tolist() == cb.tolist(), (i, s)
        assert res["cand_count"][a2:b2].tolist() == cc.tolist(), (i, s)
        assert res["cand_score"][a2:b2].tolist() == cs.tolist(), (i, s)
        nseed += len(r); ncand += len(cb)
    return nseed, ncand


@pytest.mark.parametrize("name,limit", [("c1", 3000), ("c3", 2500), ("c4", 800)])
def test_seeds_and_candidates_match_oracle(name, limit, monkeypatch):
    # keep the reference's SA sampling (every 32nd entry) so that the device's LF-step counter is the reference's
    monkeypatch.setenv("DARTGPU_SA_SAMPLE", "32")
    w = workload(name)
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"])
    reads = read_fastq_seqs(w["r1"], limit)
    O.reset_counters()
    nseed, ncand = _check_seeds(M, O, reads)
    assert nseed > len(reads) // 2 and ncand > len(reads) // 4
    # the device counts the same algorithmic work the oracle does (these define the roofline bytes)
    st, oc = M.stats(), O.counters()
    assert (st["ext_steps"], st["ext_blocks"], st["lf_steps"], st["hits"]) == \
           (oc["ext_steps"], oc["ext_blocks"], oc["lf_steps_rc"], oc["hits"])
    assert st["read_bases"] == oc["read_bases"]
    M.close(); O.close()


def test_64bit_interval_kernels(monkeypatch):
    """Human-sized texts (2G = 6.2e9) need 64-bit SA intervals; the same kernels are instantiated for both widths.
    Force the wide instantiation on a small index and require identical results."""
    monkeypatch.setenv("DARTGPU_FORCE_IDX64", "1")
    monkeypatch.setenv("DARTGPU_SA_SAMPLE", "32")
    w = workload("c3")
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"])
    reads = read_fastq_seqs(w["r1"], 1500)
    O.reset_counters()
    _check_seeds(M, O, reads)
    st, oc = M.stats(), O.counters()
    assert (st["ext_steps"], st["ext_blocks"], st["lf_steps"], st["hits"]) == (oc["ext_steps"], oc["ext_blocks"], oc["lf_steps_rc"], oc["hits"])
    M.close(); O.close()


@pytest.mark.parametrize("sample,wide", [(None, False), (1, True), (4, False), (8, True)])
def test_resident_suffix_array_densities(sample, wide, monkeypatch):
    """At load the GPU materialises a denser suffix array than the file's every-32nd sampling (the full one by default):
    the seeds must not depend on the density, and the LF steps walked must shrink accordingly (none for the full SA)."""
    if sample is not None:
        monkeypatch.setenv("DARTGPU_SA_SAMPLE", str(sample))
    if wide:
        monkeypatch.setenv("DARTGPU_FORCE_IDX64", "1")
    w = workload("c3")
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"])
    reads = read_fastq_seqs(w["r1"], 1200)
    O.reset_counters()
    _check_seeds(M, O, reads)
    st, oc = M.stats(), O.counters()
    assert (st["ext_steps"], st["ext_blocks"], st["hits"]) == (oc["ext_steps"], oc["ext_blocks"], oc["hits"])
    if sample in (None, 1):
        assert st["lf_steps"] == 0
    else:
        assert 0 < st["lf_steps"] < oc["lf_steps_rc"]
    M.close(); O.close()


@pytest.mark.parametrize("K,name", [("0", "c1"), ("5", "c4"), ("11", "c3"), ("12", "c5")])
def test_search_start_table_sizes(K, name, monkeypatch):
    """The K-mer search-start table (one gather instead of the first K-1 rank steps) must not change a single seed nor
    the algorithmic work counters, whatever K: absent K-mers, reads shorter than K, ambiguous bases fall back."""
    monkeypatch.setenv("DARTGPU_KTAB", K)
    w = workload(name)
    O = po.Oracle(w["idx"])
    md = 10000 if name == "c5" else 100
    M = capi.Mapper(w["idx"], max_dup=md)
    reads = read_fastq_seqs(w["r1"], 800) + [b"ACGTACGTAC", b"ACGTNACGTACGTTTGACCANNNACGATCAGCTAGCTAGCTAGGGATCGATCGACTAGCTAGCTAGCATCGATCAGCTACGNA", b"N" * 40,
                                              b"acgtacgtagctagctagctagctagctagcatcgatcagctagctagcta"]
    O.reset_counters()
    _check_seeds(M, O, reads, max_dup=md)
    st, oc = M.stats(), O.counters()
    assert (st["ext_steps"], st["ext_blocks"], st["hits"]) == (oc["ext_steps"], oc["ext_blocks"], oc["hits"])
    M.close(); O.close()


def test_repeat_rich_genome_with_max_dup_10000():
    w = workload("c5")
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"], max_dup=10000)
    reads = read_fastq_seqs(w["r1"], 1200)
    # a low-complexity read forces very large seed lists through the block-level sort
    reads.append(b"ACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACACAC")
    _check_seeds(M, O, reads, max_dup=10000)
    M.close(); O.close()


def test_edge_reads():
    w = workload("c1")
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"])
    base = read_fastq_seqs(w["r1"], 40)
    reads = list(base)
    reads[0] = reads[0][:13]                                   # shorter than the first search start bound
    reads[1] = reads[1][:14]
    reads[2] = reads[2][:16]
    reads[3] = reads[3][:17]
    reads[4] = b"N" * 50                                       # nothing searchable
    reads[5] = reads[5][:30] + b"N" + reads[5][31:]            # ambiguous base splits the read
    reads[6] = reads[6][:20].lower() + reads[6][20:]           # lower case maps to the same codes
    reads[7] = reads[7][:45] + b"RYK" + reads[7][48:]          # IUPAC codes are ambiguous
    reads[8] = b"A"                                            # single base
    reads[9] = reads[9] + reads[10] + reads[11]                # longer, chimeric read (ragged batch)
    reads[12] = (reads[12] * 11)[:1024]                        # maximum supported length
    reads[13] = b"N" + reads[13][1:]
    reads[14] = reads[14][:-1] + b"N"
    _check_seeds(M, O, reads)
    # empty batch
    res = M.identify_seed_pairs(capi.ReadBatch.from_list([]))
    assert res["seed_off"].tolist() == [0] and res["cand_off"].tolist() == [0]
    # too long
    with pytest.raises(capi.DartGpuError) as e:
        M.identify_seed_pairs(capi.ReadBatch.from_list([b"A" * 1025]))
    assert e.value.code == -5
    M.close(); O.close()


def test_golden_seeds(golden):
    M = capi.Mapper(os.path.join(GOLDEN, "idx"))
    reads = [c["read"].encode() for c in golden["seeds"]]
    res = M.identify_seed_pairs(capi.ReadBatch.from_list(reads))
    for i, case in enumerate(golden["seeds"]):
        a, b = res["seed_off"][i], res["seed_off"][i + 1]
        assert res["seed_rpos"][a:b].tolist() == case["rpos"]
        assert res["seed_gpos"][a:b].tolist() == case["gpos"]
        assert res["seed_len"][a:b].tolist() == case["len"]
        a2, b2 = res["cand_off"][i], res["cand_off"][i + 1]
        assert res["cand_score"][a2:b2].tolist() == case["cand_score"]
        assert res["cand_count"][a2:b2].tolist() == case["cand_nseeds"]
    M.close()


def _mut(rng, s, p):
    out = bytearray()
    for ch in s:
        x = rng.random()
        if x < p / 3:
            continue
        if x < 2 * p / 3:
            out.append(rng.choice(A))
        if x < p:
            out.append(rng.choice(A))
            continue
        out.append(ch)
    return bytes(out)


def _win(O, gpos, n):
    return bytes(A[c] for c in O.ref_codes(gpos, n))


def test_nw_matches_oracle():
    w = workload("c3")
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"])
    rng = random.Random(11)
    frags, jobs, expect = bytearray(), [], []
    def add(s1, gpos, n):
        jobs.append((len(frags), len(s1), gpos, n))
        frags.extend(s1)
        expect.append(O.nw(s1, _win(O, gpos, n)).tolist())
    G = O.G
    for _ in range(3000):                                     # the common case: tiny jobs around a mismatch / small indel
        n = rng.randint(1, 12)
        gpos = rng.randint(0, 2 * G - 40)
        s1 = _mut(rng, _win(O, gpos, n), rng.choice([0.0, 0.2, 0.5])) or b"A"
        add(s1, gpos, n)
    for _ in range(600):                                      # up to a full read against twice its length
        n = rng.randint(1, 120)
        gpos = rng.randint(0, 2 * G - 600)
        s1 = _mut(rng, _win(O, gpos, n), rng.choice([0.02, 0.1, 0.3])) if rng.random() < 0.8 else bytes(rng.choice(A) for _ in range(rng.randint(1, 90)))
        add(s1 or b"C", gpos, n)
    for _ in range(40):                                       # several 32-row strips, and more than 16 flag words per row
        m = rng.randint(33, 260)
        gpos = rng.randint(0, 2 * G - 1200)
        s1 = _mut(rng, _win(O, gpos, m), 0.08) or b"G"
        add(s1, gpos, min(2 * len(s1), rng.randint(30, 520)))
    for m, n in [(64, 64), (64, 65), (65, 64), (1, 64), (64, 1), (63, 33), (33, 63), (32, 32), (17, 48), (64, 17)]:   # class boundaries
        for rate in (0.0, 0.1, 0.4):
            gpos = rng.randint(0, 2 * G - 200)
            s1 = (_mut(rng, _win(O, gpos, max(m, n)), rate) * 2)[:m]
            add(s1, gpos, n)
    add(b"ACGTN", 1000, 5); add(b"N", 5000, 1); add(b"", 7000, 4); add(b"ACG", 9000, 0); add(b"", 100, 0)
    got = M.nw_alignment(bytes(frags), jobs)
    assert len(got) == len(jobs)
    for k, (g, e) in enumerate(zip(got, expect)):
        assert g.tolist() == e, (k, jobs[k])
    st = M.stats()
    assert st["nw_jobs"] == len(jobs) and st["nw_cells"] == sum(j[1] * j[3] for j in jobs)
    M.close(); O.close()


def test_nw_golden(golden):
    M = capi.Mapper(os.path.join(GOLDEN, "idx"))
    frags, jobs = bytearray(), []
    for c in golden["nw"]:
        jobs.append((len(frags), len(c["s1"]), c["gpos"], c["n"]))
        frags.extend(c["s1"].encode())
    got = M.nw_alignment(bytes(frags), jobs)
    O = po.Oracle(os.path.join(GOLDEN, "idx"))
    for c, ops in zip(golden["nw"], got):
        a, b = po.ops_to_strings(c["s1"].encode(), _win(O, c["gpos"], c["n"]), ops)
        assert (a.decode(), b.decode()) == (c["a"], c["b"])
    M.close(); O.close()


@pytest.mark.parametrize("cap", [None, "40"])
def test_kmer_matches_oracle(cap, monkeypatch):
    """Both k-mer paths: the grid-wide scan that keeps matches as records, and (cap=40: nearly every job overflows its
    record capacity; N / IUPAC symbols; the 300 kb windows when their capacity would exceed the limit) the ring kernel."""
    if cap:
        monkeypatch.setenv("DARTGPU_KMER_CAP", cap)
    w = workload("c5")                                        # repeat-rich: many equal-PosDiff runs
    O = po.Oracle(w["idx"])
    M = capi.Mapper(w["idx"])
    rng = random.Random(12)
    G = O.G
    frags, jobs, expect = bytearray(), [], []
    def add(f1, gpos, glen):
        jobs.append((len(frags), len(f1), gpos, glen))
        frags.extend(f1)
        expect.append(O.kmer_pair(f1, _win(O, gpos, glen)))
    for it in range(1500):
        glen = rng.choice([rng.randint(26, 300), rng.randint(300, 6000), rng.randint(6000, 60000)])
        if it % 100 == 7:
            glen = rng.randint(200000, 500000)
        gpos = rng.randint(0, 2 * G - glen - 1)
        L1 = rng.randint(21, 101) if it % 10 else rng.randint(101, 250)
        mode = rng.random()
        if mode < 0.7 and glen > L1 + 2:
            p = rng.randint(0, glen - L1 - 1)
            f1 = _mut(rng, _win(O, gpos + p, L1), rng.choice([0, 0.03, 0.1]))
        else:
            f1 = bytes(rng.choice(A) for _ in range(L1))
        if rng.random() < 0.12 and len(f1) > 10:
            q = rng.randrange(len(f1))
            f1 = f1[:q] + rng.choice([b"N", b"n", b"R", b"NN"]) + f1[q + 1:]
        add(f1 or b"A", gpos, glen)
    add(b"ACGTACG", 100, 500); add(b"ACGTACGT", 100, 7); add(b"ACGTACGTAC", 100, 8); add(b"", 5, 100)
    got = M.kmer_reseed(bytes(frags), jobs)
    found = 0
    for k, (h, e) in enumerate(zip(got, expect)):
        assert (int(h["rpos"]), int(h["gpos"]), int(h["len"])) == e, (k, jobs[k])
        found += e[2] > 0
    assert found > 500
    M.close(); O.close()


def test_kmer_golden(golden):
    M = capi.Mapper(os.path.join(GOLDEN, "idx"))
    frags, jobs = bytearray(), []
    for c in golden["kmer"]:
        jobs.append((len(frags), len(c["f1"]), c["gpos"], c["glen"]))
        frags.extend(c["f1"].encode())
    got = M.kmer_reseed(bytes(frags), jobs)
    for c, h in zip(golden["kmer"], got):
        assert [int(h["rpos"]), int(h["gpos"]), int(h["len"])] == c["out"]
    M.close()
