"""The reference binaries built by oracle/Makefile: determinism of the canonicalised build and
agreement of the in-process taps with the stock binary."""
import os
import subprocess

from conftest import GOLDEN, canonical_env, need_ref, read_fastq_seqs
from oracle import pyoracle as po


def _records(path):
    return [l for l in open(path) if not l.startswith("@")]


def test_canon_matches_stock_on_mapped_reads_and_golden(tmp_path):
    need_ref()
    out = {}
    for b in ("dart_ref", "dart_canon"):
        sam, junc = str(tmp_path / f"{b}.sam"), str(tmp_path / f"{b}.junc")
        subprocess.run([os.path.join(po.REF_DIR, b), "-i", GOLDEN + "/idx", "-f", GOLDEN + "/pe1.fq", "-f2", GOLDEN + "/pe2.fq",
                        "-t", "1", "-mis", "5", "-o", sam, "-j", junc], check=True, stdout=subprocess.DEVNULL,
                       env=canonical_env() if b == "dart_canon" else None)
        out[b] = (_records(sam), open(junc).read())
    assert out["dart_canon"][0] == _records(GOLDEN + "/pe.sam")      # the committed golden SAM is reproducible
    assert out["dart_canon"][1] == open(GOLDEN + "/pe.junc").read()
    assert out["dart_ref"][1] == out["dart_canon"][1]
    # the stock binary may differ only in the FLAG of pairs that contain an unmapped read (uninitialised sub_score)
    for a, b in zip(out["dart_ref"][0], out["dart_canon"][0]):
        fa, fb = a.split("\t"), b.split("\t")
        assert fa[:1] + fa[2:] == fb[:1] + fb[2:]


def test_taps_replay_readmapping(tmp_path):
    """oracle/ref_taps.cpp's per-read driver reproduces the canonical binary's SAM records (it replays ReadMapping's
    body). Runs in its own process: the reference keeps one index per process in globals."""
    need_ref()
    import sys
    from conftest import ROOT
    out = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_replay.py"), GOLDEN + "/idx", GOLDEN + "/pe1.fq",
                          GOLDEN + "/pe2.fq", "--mis", "5"], check=True, capture_output=True).stdout.decode()
    lines = [l for l in out.splitlines(True) if "\t" in l and not l.startswith("Load")]
    assert lines == _records(GOLDEN + "/pe.sam")


def test_integration_patch_applies_and_the_patched_binary_is_built(tmp_path):
    """integration/dart_gpu.patch against the reference's own sources: applies cleanly (when /root/reference is here), touches
    nothing but Mapping.cpp / main.cpp / makefile, and oracle/Makefile has produced oracle/_ref/dart_gpu linked against
    libdartgpu.so.  Without a GPU the binary must refuse to map (no CPU fallback behind the boundary)."""
    import shutil
    from conftest import ROOT
    need_ref()
    patch = os.path.join(ROOT, "integration", "dart_gpu.patch")
    files = [l.split()[1] for l in open(patch) if l.startswith("+++ ")]
    assert sorted(files) == ["b/src/Mapping.cpp", "b/src/main.cpp", "b/src/makefile"]
    added = [l for l in open(patch) if l.startswith("+") and not l.startswith("+++")]
    assert len(added) <= 16                                   # the binding lives in dart_gpu_glue.cpp, not in the patch
    ref_src = "/root/reference/src"
    if os.path.isdir(ref_src):
        os.makedirs(tmp_path / "src")
        for f in ("Mapping.cpp", "main.cpp", "makefile"):
            shutil.copy(os.path.join(ref_src, f), tmp_path / "src" / f)
        r = subprocess.run(["patch", "-p1", "--binary", "--dry-run", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
    exe = os.path.join(po.REF_DIR, "dart_gpu")
    assert os.path.exists(exe), "oracle/_ref/dart_gpu is not built (make -C oracle dart_gpu)"
    ldd = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libdartgpu.so" in ldd and "not found" not in ldd
    from conftest import have_gpu
    if not have_gpu():
        r = subprocess.run([exe, "-i", GOLDEN + "/idx", "-f", GOLDEN + "/se.fq", "-o", str(tmp_path / "x.sam"), "-j", str(tmp_path / "x.junc")],
                           capture_output=True, text=True)
        assert r.returncode != 0 and "no CPU fallback" in (r.stdout + r.stderr)
