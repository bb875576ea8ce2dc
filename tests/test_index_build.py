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
