import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from dart_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WORK = os.environ.get("DART_TEST_DIR", "/tmp/dart_b200_tests")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session", autouse=True)
def _build_checkers():
    po.build(ref=True)  # `make ref` is a no-op where /root/reference is absent (prebuilt oracle/_ref is used)


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN, "stage.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def golden_oracle():
    o = po.Oracle(os.path.join(GOLDEN, "idx"))
    yield o
    o.close()


def need_ref():
    if not po.have_reference():
        pytest.skip("oracle/_ref is not built (no /root/reference here and no prebuilt binaries)")


# name -> (config id, genome scale, number of reads/pairs)
WORKLOADS = {
    "c1": (1, 0.10, 4000),      # 460 kbp, SE 100 bp
    "c2": (2, 0.10, 3000),      # 460 kbp, PE 2x101
    "c3": (3, 0.0008, 3000),    # 2.5 Mbp in 24 contigs with gene models, spliced PE 2x101
    "c4": (4, 0.0008, 1200),    # same genome, PE 2x250 3% + indels, -mis 10
    "c5": (5, 0.20, 2000),      # 920 kbp repeat-rich, -m -max_dup 10000 -all_sj
}


def workload(name):
    """Materialise (once per machine) a scaled BASELINE.json config: FASTA, FASTQ, index. Returns a dict."""
    need_ref()
    cfg, scale, n = WORKLOADS[name]
    d = os.path.join(WORK, name)
    stamp = os.path.join(d, "ready")
    if not os.path.exists(stamp):
        w = synth.materialise(cfg, d, n, scale)
        po.build_index(w["fasta"], os.path.join(d, "idx"))
        with open(os.path.join(d, "flags.json"), "w") as f:
            json.dump(w["flags"], f)
        open(stamp, "w").close()
    with open(os.path.join(d, "flags.json")) as f:
        flags = json.load(f)
    r2 = os.path.join(d, "r2.fq")
    return dict(dir=d, idx=os.path.join(d, "idx"), r1=os.path.join(d, "r1.fq"), r2=r2 if os.path.exists(r2) else None,
                flags=flags, cfg=cfg)


def read_fastq_seqs(path, limit=None):
    out = []
    with open(path, "rb") as f:
        for i, line in enumerate(f):
            if i % 4 == 1:
                out.append(line.rstrip(b"\n"))
                if limit and len(out) >= limit:
                    break
    return out


def run_reference(w, binary="dart_canon", threads=1, extra=(), tag="ref"):
    """Run a reference binary of oracle/_ref on a workload; returns (sam path, junction path)."""
    need_ref()
    sam = os.path.join(w["dir"], f"{tag}.sam")
    junc = os.path.join(w["dir"], f"{tag}.junc")
    cmd = [os.path.join(po.REF_DIR, binary), "-i", w["idx"], "-f", w["r1"]]
    if w["r2"]:
        cmd += ["-f2", w["r2"]]
    cmd += ["-t", str(threads), "-o", sam, "-j", junc] + list(w["flags"]) + list(extra)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, env=canonical_env())
    return sam, junc


def canonical_env():
    """The canonical oracle = the reference with every uninitialised object zero-filled: the stack
    object `ReadItem_t read` by the build flags of oracle/Makefile (dart_canon), heap objects (`new AlignmentReport_t[]`,
    whose iFrag / coor.bDir the reference prints without ever assigning them for secondary -m reports) by glibc's
    MALLOC_PERTURB_=255, which fills fresh allocations with ~0xFF = 0x00."""
    env = dict(os.environ)
    env["MALLOC_PERTURB_"] = "255"
    env["GLIBC_TUNABLES"] = "glibc.malloc.tcache_count=0"   # tcache hits bypass the perturbation
    return env


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
