"""The device orchestration logic (dart_b200/csrc/report_logic.cuh — the code the report kernels run one thread per
candidate) compiled for the host and driven by the oracle in place of the CUDA kernels (tests/host/logic_harness.cpp):
its SAM and junctions must be byte-identical to the canonical reference. This proves the phase-split restatement of
GenMappingReport / pairing / flags / MAPQ on a box without a GPU."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, need_ref, run_reference, workload

HARNESS = os.path.join(ROOT, "tests", "host", "logic_harness")


@pytest.fixture(scope="module")
def harness():
    subprocess.run(["g++", "-O2", "-std=c++17", "-fopenmp", "-Wno-attributes", "-I/usr/local/cuda/include",
                    os.path.join(ROOT, "tests", "host", "logic_harness.cpp"), os.path.join(ROOT, "oracle", "dart_oracle.cpp"),
                    "-o", HARNESS], check=True)
    return HARNESS


def _records(p):
    return [l for l in open(p) if not l.startswith("@")]


def _run(harness, w, extra, tag):
    sam, junc = os.path.join(w["dir"], tag + ".sam"), os.path.join(w["dir"], tag + ".junc")
    cmd = [harness, "-i", w["idx"], "-f", w["r1"]] + (["-f2", w["r2"]] if w["r2"] else []) + ["-o", sam, "-j", junc]
    subprocess.run(cmd + list(w["flags"]) + list(extra), check=True, stdout=subprocess.DEVNULL)
    return sam, junc


@pytest.mark.parametrize("name,extra", [("c1", ()), ("c2", ("-mis", "5")), ("c3", ("-mis", "5")), ("c3", ()), ("c4", ()),
                                        ("c5", ()), ("c5", ("-mis", "5"))])
def test_logic_matches_reference(harness, name, extra):
    need_ref()
    w = workload(name)
    tag = "mis5" if extra else "named"
    rs, rj = run_reference(w, "dart_canon", 1, extra, tag="ref_" + tag)
    hs, hj = _run(harness, w, extra, "logic_" + tag)
    a, b = _records(hs), _records(rs)
    bad = [(x, y) for x, y in zip(a, b) if x != y]
    assert len(a) == len(b) and not bad, f"{len(bad)} records differ; first:\n{bad[0][0]}{bad[0][1]}" if bad else "record count"
    assert open(hj).read() == open(rj).read()


def test_logic_golden(harness, tmp_path):
    for tag, r1, r2 in (("se", "se.fq", None), ("pe", "pe1.fq", "pe2.fq")):
        w = dict(dir=str(tmp_path), idx=GOLDEN + "/idx", r1=os.path.join(GOLDEN, r1), r2=os.path.join(GOLDEN, r2) if r2 else None,
                 flags=["-mis", "5"])
        hs, hj = _run(harness, w, (), "logic_" + tag)
        assert _records(hs) == _records(os.path.join(GOLDEN, tag + ".sam"))
        assert open(hj).read() == open(os.path.join(GOLDEN, tag + ".junc")).read()


def test_rank_arithmetic_on_the_host():
    """dart_b200/csrc/rank.cuh (the popcount arithmetic of the search / locate kernels) against a symbol-by-symbol count."""
    exe = os.path.join(ROOT, "tests", "host", "rank_test")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "host", "rank_test.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert "rank ok" in out
