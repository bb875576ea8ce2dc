"""The CPU oracle against the committed golden vectors (generated from the reference by tools/make_golden.py)."""
import numpy as np

from oracle import pyoracle as po


def test_seeds_and_candidates(golden, golden_oracle):
    O = golden_oracle
    assert len(golden["seeds"]) >= 100
    for case in golden["seeds"]:
        codes = po.encode(case["read"])
        r, g, l = O.seeds(codes)
        assert r.tolist() == case["rpos"] and g.tolist() == case["gpos"] and l.tolist() == case["len"]
        cb, cc, cs = O.cluster(len(codes), r, g, l)
        assert cs.tolist() == case["cand_score"] and cc.tolist() == case["cand_nseeds"]
        flat = np.concatenate([np.arange(b, b + n) for b, n in zip(cb, cc)]) if len(cb) else np.zeros(0, int)
        assert r[flat].tolist() == case["cand_seed_rpos"] and g[flat].tolist() == case["cand_seed_gpos"]
        pd = [max(int(g[b] - r[b]), 0) for b in cb]
        assert pd == case["cand_posdiff"]


def test_reads_with_ambiguous_bases_are_covered(golden):
    assert any("N" in c["read"] for c in golden["seeds"]) and any("n" in c["read"] for c in golden["seeds"])


def test_nw(golden, golden_oracle):
    O = golden_oracle
    A = b"ACGT"
    for case in golden["nw"]:
        s2 = bytes(A[c] for c in O.ref_codes(case["gpos"], case["n"]))
        s1 = case["s1"].encode()
        a, b = po.ops_to_strings(s1, s2, O.nw(s1, s2))
        assert (a.decode(), b.decode()) == (case["a"], case["b"])


def test_kmer(golden, golden_oracle):
    O = golden_oracle
    A = b"ACGT"
    found = 0
    for case in golden["kmer"]:
        win = bytes(A[c] for c in O.ref_codes(case["gpos"], case["glen"]))
        out = O.kmer_pair(case["f1"].encode(), win)
        assert list(out) == case["out"]
        found += out[2] > 0
    assert found > 20


def test_gapped_partition(golden, golden_oracle):
    O = golden_oracle
    nz = 0
    for case in golden["gapped"]:
        a = case["args"]
        out = O.gapped_partition(a["seq"].encode(), a["rgaps"], a["l_rpos"], a["l_rlen"], a["l_gpos"], a["l_glen"],
                                 a["r_rpos"], a["r_gpos"], case["max_mismatch"])
        assert list(out) == case["out"]
        nz += out[1] > 0 or out[2] > 0
    assert nz > 5
