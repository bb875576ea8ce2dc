"""End-to-end parity: dart_b200_map (reader + SAM writer stand-in over libdartgpu.so) against the canonicalised
reference binary (oracle/_ref/dart_canon -t 1) on scaled versions of all five BASELINE.json
configs, with the flags each config names and with -mis 5. SAM records and junctions.tab must be
byte-identical."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT, run_reference, workload

pytestmark = pytest.mark.gpu
TOOL = os.path.join(ROOT, "dart_b200", "dart_b200_map")


def _records(path):
    return [l for l in open(path) if not l.startswith("@")]


def _headers(path):
    return [l for l in open(path) if l.startswith("@")]


def _run_gpu(w, extra=(), tag="gpu", more=()):
    sam, junc = os.path.join(w["dir"], f"{tag}.sam"), os.path.join(w["dir"], f"{tag}.junc")
    cmd = [TOOL, "-i", w["idx"], "-f", w["r1"]] + (["-f2", w["r2"]] if w["r2"] else [])
    cmd += ["-o", sam, "-j", junc] + list(w["flags"]) + list(extra) + list(more)
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return sam, junc


def _compare(a, b, ja, jb):
    ra, rb = _records(a), _records(b)
    assert len(ra) == len(rb)
    bad = [(x, y) for x, y in zip(ra, rb) if x != y]
    assert not bad, f"{len(bad)} of {len(ra)} SAM records differ; first:\n{bad[0][0]}{bad[0][1]}"
    assert _headers(a) == _headers(b)
    assert open(ja).read() == open(jb).read()


@pytest.mark.parametrize("name", ["c1", "c2", "c3", "c4", "c5"])
@pytest.mark.parametrize("extra", [(), ("-mis", "5")], ids=["as-named", "mis5"])
def test_sam_and_junctions_identical(name, extra):
    w = workload(name)
    if name == "c4" and extra:
        pytest.skip("config 4 already names -mis 10")
    tag = "mis5" if extra else "named"
    rs, rj = run_reference(w, "dart_canon", 1, extra, tag="ref_" + tag)
    gs, gj = _run_gpu(w, extra, tag="gpu_" + tag)
    _compare(gs, rs, gj, rj)


def test_result_is_independent_of_batching_and_host_threads():
    w = workload("c3")
    rs, rj = run_reference(w, "dart_canon", 1, ("-mis", "5"), tag="ref_mis5")
    gs, gj = _run_gpu(w, ("-mis", "5"), tag="gpu_small_batches", more=("-batch", "500", "-t", "3"))
    _compare(gs, rs, gj, rj)


def test_two_gpus_give_the_same_records():
    """Reads shard over the GPUs of a box (contiguous ranges, no collective); the second GPU's index is replicated from the
    first over NVLink (cudaMemcpyPeer) instead of a second upload.  Same records, same junctions."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    w = workload("c3")
    rs, rj = run_reference(w, "dart_canon", 1, ("-mis", "5"), tag="ref_mis5")
    gs, gj = _run_gpu(w, ("-mis", "5"), tag="gpu_two_devices", more=("-devices", "0,1", "-batch", "700"))
    _compare(gs, rs, gj, rj)


def test_golden_sam():
    """The committed golden SAM (generated in the build container from the reference) — no oracle/_ref needed."""
    w = dict(dir=os.environ.get("DART_TEST_DIR", "/tmp/dart_b200_tests"), idx=GOLDEN + "/idx", flags=["-mis", "5"])
    os.makedirs(w["dir"], exist_ok=True)
    for tag, r1, r2 in (("se", "se.fq", None), ("pe", "pe1.fq", "pe2.fq")):
        w.update(r1=os.path.join(GOLDEN, r1), r2=os.path.join(GOLDEN, r2) if r2 else None)
        gs, gj = _run_gpu(w, tag="golden_" + tag)
        _compare(gs, os.path.join(GOLDEN, tag + ".sam"), gj, os.path.join(GOLDEN, tag + ".junc"))


@pytest.mark.parametrize("cfg,n,extra", [(1, 100_000, ()), (2, 1_000_000, ()), (2, 200_000, ("-mis", "5"))],
                         ids=["config0-100k-SE", "config1-1M-PE", "config1-200k-PE-mis5"])
def test_full_size_baseline_configs(cfg, n, extra, tmp_path_factory):
    """BASELINE.json's own sizes for the two configs that fit a GPU-box call: config[0] (100 k SE x 100 bp) and config[1]
    (1 M pairs 2x101) on the 4.6 Mbp genome — every SAM record and junctions.tab against the canonical reference."""
    import hashlib
    from conftest import WORK, need_ref
    from dart_b200 import synth
    from oracle import pyoracle as po
    need_ref()
    d = os.path.join(WORK, f"full_c{cfg}_{n}")
    os.makedirs(d, exist_ok=True)
    g = synth.config_genome(cfg)
    fa = os.path.join(d, "genome.fa")
    if not os.path.exists(os.path.join(d, "idx.bwt")):
        synth.write_fasta(fa, g)
        po.build_index(fa, os.path.join(d, "idx"))
    m1, m2, flags = synth.config_reads(cfg, g, n)
    synth.write_fastq(os.path.join(d, "r1.fq"), m1, 1 if m2 is not None else None)
    if m2 is not None:
        synth.write_fastq(os.path.join(d, "r2.fq"), m2, 2)
    w = dict(dir=d, idx=os.path.join(d, "idx"), r1=os.path.join(d, "r1.fq"), r2=os.path.join(d, "r2.fq") if m2 is not None else None,
             flags=flags)
    rs, rj = run_reference(w, "dart_canon", 1, extra, tag="ref")
    gs, gj = _run_gpu(w, extra, tag="gpu")
    md5 = lambda p: hashlib.md5(open(p, "rb").read()).hexdigest()  # noqa: E731
    if md5(gs) != md5(rs):
        _compare(gs, rs, gj, rj)          # pinpoints the first differing record
    assert md5(gj) == md5(rj)


def test_sharding_over_contexts_and_gpus():
    """The N-GPU path (contiguous read range per device entry, results merged in input order, junction counts summed by key
    like UpdateGlobalSJMap, /root/reference/src/Mapping.cpp:567-577, :676-678) over EVERY visible GPU — and, so that the
    shard / merge code runs on a one-GPU box too, over three contexts of device 0 and over an uneven mix."""
    import torch
    w = workload("c5")                              # -m -max_dup 10000 -all_sj: multi-hit reports + junction sums
    rs, rj = run_reference(w, "dart_canon", 1, (), tag="ref_named")
    n = torch.cuda.device_count()
    lists = ["0,0,0", ",".join(str(i) for i in range(n)), ",".join(str(i % n) for i in range(n + 3))]
    for k, devs in enumerate(lists):
        gs, gj = _run_gpu(w, (), tag=f"gpu_shard{k}", more=("-devices", devs, "-batch", "300"))
        _compare(gs, rs, gj, rj)
    w = workload("c3")
    rs, rj = run_reference(w, "dart_canon", 1, ("-mis", "5"), tag="ref_mis5")
    gs, gj = _run_gpu(w, ("-mis", "5"), tag="gpu_shard_c3", more=("-devices", lists[2], "-batch", "500"))
    _compare(gs, rs, gj, rj)


def test_pool_overflow_is_retried_not_truncated(monkeypatch):
    """Device pools are sized by guesses (the counts only exist on the device): a batch that overflows one must be re-run
    with a bigger pool and give the very same records.  DARTGPU_CAP_SHRINK makes every first guess 1000x too small."""
    w = workload("c4")
    rs, rj = run_reference(w, "dart_canon", 1, (), tag="ref_named")
    monkeypatch.setenv("DARTGPU_CAP_SHRINK", "1000")
    gs, gj = _run_gpu(w, (), tag="gpu_shrunk")
    _compare(gs, rs, gj, rj)


def _run_patched(w, extra=(), tag="dart_gpu", threads=1, env_extra=None):
    """oracle/_ref/dart_gpu = the reference's own binary (its CLI, reader, SAM writer) with integration/dart_gpu.patch applied:
    ReadMapping()'s per-read loop runs in libdartgpu.so."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dart_gpu")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dart_gpu is not built")
    sam, junc = os.path.join(w["dir"], f"{tag}.sam"), os.path.join(w["dir"], f"{tag}.junc")
    cmd = [exe, "-i", w["idx"], "-f", w["r1"]] + (["-f2", w["r2"]] if w["r2"] else [])
    cmd += ["-t", str(threads), "-o", sam, "-j", junc] + list(w["flags"]) + list(extra)
    env = dict(os.environ)
    env.update(env_extra or {})
    subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL, env=env)
    return sam, junc


@pytest.mark.parametrize("name,extra", [("c1", ()), ("c2", ()), ("c3", ("-mis", "5")), ("c5", ())])
def test_patched_reference_binary_is_byte_identical(name, extra):
    """The boundary as an artefact: integration/dart_gpu.patch + dart_gpu_glue.cpp built into the reference
    (/root/reference/src/Mapping.cpp:598-640, main.cpp:220-223).  Output of `dart_gpu -t 1` == `dart_canon -t 1`, byte for byte,
    with batches of 8000 reads so that several GPU calls and chunk boundaries are exercised."""
    w = workload(name)
    tag = "mis5" if extra else "named"
    rs, rj = run_reference(w, "dart_canon", 1, extra, tag="ref_" + tag)
    gs, gj = _run_patched(w, extra, tag="patched_" + tag, env_extra={"DART_GPU_BATCH": "8000"})
    _compare(gs, rs, gj, rj)


def test_patched_reference_binary_with_worker_threads():
    """-t 3: three reference worker threads, three GPU contexts sharing one resident index; chunks are written in completion
    order, so the records are compared as a multiset."""
    w = workload("c3")
    rs, rj = run_reference(w, "dart_canon", 1, ("-mis", "5"), tag="ref_mis5")
    gs, gj = _run_patched(w, ("-mis", "5"), tag="patched_t3", threads=3, env_extra={"DART_GPU_BATCH": "8000", "DART_GPU_DEVICES": "0,0"})
    assert sorted(_records(gs)) == sorted(_records(rs))
    assert open(gj).read() == open(rj).read()


def test_host_reader_path_still_matches():
    """-hostpath: the stand-in host reader / formatter over dartgpu_map_reads (the path FASTA input takes)."""
    w = workload("c3")
    rs, rj = run_reference(w, "dart_canon", 1, ("-mis", "5"), tag="ref_mis5")
    gs, gj = _run_gpu(w, ("-mis", "5"), tag="gpu_hostpath", more=("-hostpath", "-batch", "700"))
    _compare(gs, rs, gj, rj)


def test_device_ingest_and_sam_text_on_awkward_fastq(tmp_path):
    """GPU-side FASTQ parsing and SAM text against the reference's reader / writer (GetData.cpp:77-179, Mapping.cpp:208-369) on
    records that exercise their corners: lower-case and ambiguous bases in both mates (mate 2 is reverse-complemented with
    GetComplementaryBase: anything but ACGT/acgt becomes N; a forward-strand mate-2 record prints the complement of THAT),
    reads of different lengths, headers with '/', blanks, tabs and leading '@@', a last line without its newline."""
    import random
    need = os.path.join(ROOT, "oracle", "_ref", "dart_canon")
    if not os.path.exists(need):
        pytest.skip("oracle/_ref is not built")
    rng = random.Random(7)

    def recs(path):
        L = open(path).read().split("\n")
        return [L[i:i + 4] for i in range(0, len(L) - 3, 4)]
    r1, r2 = recs(os.path.join(GOLDEN, "pe1.fq")), recs(os.path.join(GOLDEN, "pe2.fq"))
    for k, (a, b) in enumerate(zip(r1, r2)):
        for r in (a, b):
            s = list(r[1])
            mode = (k + (r is b)) % 6
            if mode == 0:
                for _ in range(3):
                    p = rng.randrange(len(s)); s[p] = s[p].lower()
            elif mode == 1:
                s[rng.randrange(len(s))] = "N"
            elif mode == 2:
                s[rng.randrange(len(s))] = rng.choice("RYKMn")
            elif mode == 3:
                cut = rng.randrange(40, len(s)); s = s[:cut]; r[3] = r[3][:cut]
            r[1] = "".join(s)
        base = a[0].split("/")[0].split(" ")[0]
        style = k % 5
        if style == 1:
            a[0], b[0] = base + " first mate", base + " second mate"
        elif style == 2:
            a[0], b[0] = base + "\tx", base + "\ty"
        elif style == 3:
            a[0], b[0] = "@" + base + "/1", "@" + base + "/2"
    d = str(tmp_path)
    for name, rr in (("a1.fq", r1), ("a2.fq", r2)):
        with open(os.path.join(d, name), "w") as f:
            f.write("\n".join("\n".join(r) for r in rr))        # no newline behind the last quality line
    w = dict(dir=d, idx=GOLDEN + "/idx", r1=os.path.join(d, "a1.fq"), r2=os.path.join(d, "a2.fq"), flags=["-mis", "5"])
    for extra, tag in (((), "plain"), (("-m",), "multi")):
        rs, rj = run_reference(w, "dart_canon", 1, extra, tag="ref_" + tag)
        gs, gj = _run_gpu(w, extra, tag="gpu_" + tag, more=("-batch", "64"))
        _compare(gs, rs, gj, rj)
    w1 = dict(w, r2=None)
    rs, rj = run_reference(w1, "dart_canon", 1, (), tag="ref_se")
    gs, gj = _run_gpu(w1, (), tag="gpu_se", more=("-batch", "50"))
    _compare(gs, rs, gj, rj)
