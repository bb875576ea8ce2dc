"""Index construction: the files dartgpu_index_build writes must be byte-identical to the reference's own bwt_index
output (the BWT of a text is unique). CPU part: the numpy restatement of the file formats (oracle/index_format.py) is
pinned against the committed files the reference's builder wrote (tests/golden/idx.*)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, WORKLOADS, workload
from dart_b200 import synth
from oracle import index_format as fmt

EXTS = (".bwt", ".sa", ".pac", ".ann", ".amb")


def _golden_genome():
    codes = fmt.read_pac(os.path.join(GOLDEN, "idx.pac"))
    names, seqs = [], []
    with open(os.path.join(GOLDEN, "idx.ann")) as f:
        lines = f.read().split("\n")
    n = int(lines[0].split()[1])
    for i in range(n):
        names.append(lines[1 + 2 * i].split()[1])
        off, ln, _ = (int(x) for x in lines[2 + 2 * i].split())
        seqs.append(codes[off:off + ln].copy())
    return synth.Genome(names, seqs)


def test_format_restatement_matches_reference_files():
    g = _golden_genome()
    out = fmt.index_files(np.concatenate(g.seqs))
    for e in (".bwt", ".sa", ".pac"):
        assert out[e] == open(os.path.join(GOLDEN, "idx" + e), "rb").read(), e


@pytest.mark.parametrize("kind", ["random", "contigs", "repeats", "tiny"])
def test_format_restatement_matches_reference_builder(kind, tmp_path):
    """The numpy restatement against the reference's own bwt_index on freshly generated small genomes: multi-contig
    (sequence boundaries are invisible to the BWT), repeat-rich (long ties in the suffix sort), and a text shorter than one
    Occ block."""
    from conftest import need_ref
    from oracle import pyoracle as po
    need_ref()
    if kind == "random":
        g = synth.random_genome(30011, 1, seed=7)
    elif kind == "contigs":
        g = synth.random_genome(26000, 5, seed=8)
    elif kind == "repeats":
        g = synth.random_genome(40000, 2, seed=9)
        synth.add_segmental_duplications(g, 0.4, seed=9, seg_min=300, seg_max=3000)
    else:
        g = synth.random_genome(57, 1, seed=10)
        g.seqs = [s[:57] for s in g.seqs]
    fa = str(tmp_path / "g.fa")
    synth.write_fasta(fa, g)
    po.build_index(fa, str(tmp_path / "ref"))
    out = fmt.index_files(np.concatenate(g.seqs))
    for e in (".bwt", ".sa", ".pac"):
        assert out[e] == open(str(tmp_path / "ref") + e, "rb").read(), (kind, e)
    synth.write_index_meta(str(tmp_path / "mine"), g)
    for e in (".pac", ".ann", ".amb"):
        assert open(str(tmp_path / "mine") + e, "rb").read() == open(str(tmp_path / "ref") + e, "rb").read(), (kind, e)


def test_meta_files_match_reference(tmp_path):
    g = _golden_genome()
    synth.write_index_meta(str(tmp_path / "idx"), g)
    for e in (".pac", ".ann", ".amb"):
        assert open(tmp_path / ("idx" + e), "rb").read() == open(os.path.join(GOLDEN, "idx" + e), "rb").read(), e


@pytest.mark.gpu
@pytest.mark.parametrize("per_pass", [0, 9000])
def test_gpu_builder_matches_golden_index(tmp_path, per_pass):
    from dart_b200 import capi
    g = _golden_genome()
    capi.index_build(g, str(tmp_path / "idx"), max_suffixes_per_pass=per_pass)
    for e in EXTS:
        assert open(tmp_path / ("idx" + e), "rb").read() == open(os.path.join(GOLDEN, "idx" + e), "rb").read(), e


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "c3", "c5"])
def test_gpu_builder_matches_reference_builder(name, tmp_path):
    """Random, multi-contig with gene models, and repeat-rich (30 % segmental duplications: long ties) genomes."""
    from dart_b200 import capi
    w = workload(name)                      # its index was written by the reference's bwt_index
    cfg, scale, _ = WORKLOADS[name]
    g = synth.config_genome(cfg, scale)
    capi.index_build(g, str(tmp_path / "idx"), max_suffixes_per_pass=0 if name != "c3" else 700_000)
    for e in EXTS:
        a, b = open(tmp_path / ("idx" + e), "rb").read(), open(w["idx"] + e, "rb").read()
        assert len(a) == len(b) and a == b, e


@pytest.mark.gpu
@pytest.mark.parametrize("length,contigs,per_pass", [(57, 1, 0), (1000, 3, 0), (1000, 3, 64), (40000, 2, 5000)])
def test_gpu_builder_matches_format_oracle_on_small_texts(length, contigs, per_pass, tmp_path):
    """Edge sizes against the numpy restatement: a text shorter than one Occ block, many tiny sort passes, and a
    repeat-rich text whose ties go through the exact tie-break."""
    from dart_b200 import capi
    g = synth.random_genome(length, contigs, seed=20 + length)
    if contigs == 1:
        g.seqs = [s[:length] for s in g.seqs]
    if length >= 40000:
        synth.add_segmental_duplications(g, 0.4, seed=9, seg_min=300, seg_max=3000)
    capi.index_build(g, str(tmp_path / "idx"), max_suffixes_per_pass=per_pass)
    out = fmt.index_files(np.concatenate(g.seqs))
    for e in (".bwt", ".sa", ".pac"):
        assert open(tmp_path / ("idx" + e), "rb").read() == out[e], e
