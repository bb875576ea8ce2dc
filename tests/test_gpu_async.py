"""dartgpu_submit / dartgpu_wait: one host thread keeps several batches in flight (one per context); the results are the
ones the synchronous call gives, whatever the interleaving."""
import numpy as np
import pytest

from conftest import read_fastq_seqs, workload
from dart_b200 import capi
from dart_b200.shard import shard_bounds

pytestmark = pytest.mark.gpu


def _batch(w):
    r1, r2 = read_fastq_seqs(w["r1"]), read_fastq_seqs(w["r2"])
    comp = bytes.maketrans(b"ACGTacgt", b"TGCATGCA")
    seqs = []
    for a, b in zip(r1, r2):
        seqs += [a, b.translate(comp)[::-1]]
    return capi.ReadBatch.from_list(seqs)


def _decoded(r):
    """Records as a consumer sees them: the slices of the CIGAR-text and junction pools are laid out in the order the warps
    arrive (every report carries its cigar_off, every junction record its read), so compare contents, not pool layouts."""
    rep = r["reports"]
    cig = [r["cigars"][o:o + l] for o, l in zip(rep["cigar_off"].tolist(), rep["cigar_len"].tolist())]
    fields = [rep[n].tolist() for n in rep.dtype.names if n not in ("cigar_off", "reserved")]
    return r["reads"].tobytes(), fields, cig, sorted(r["junctions"].tolist())


def _same(a, b):
    assert _decoded(a) == _decoded(b)


def test_in_flight_batches_equal_synchronous_calls():
    w = workload("c3")
    batch = _batch(w)
    K = 3
    ms = [capi.Mapper(w["idx"], device=0, pair_end=1, max_mismatch=5) for _ in range(K)]
    b = shard_bounds(batch.n, 7, True)
    subs = []
    for lo, hi in zip(b[:-1], b[1:]):
        off = batch.offsets[lo:hi + 1]
        subs.append(capi.ReadBatch(batch.bases[off[0]:off[-1]].copy(), (off - off[0]).copy()))
    want = [ms[0].map_reads(sb) for sb in subs]
    got = [None] * len(subs)
    pending = {}
    for i, sb in enumerate(subs):               # round-robin, a context is waited for only when its next batch is due
        k = i % K
        if k in pending:
            got[pending[k]] = ms[k].wait()
        ms[k].submit(sb)
        pending[k] = i
    for k, i in pending.items():
        got[i] = ms[k].wait()
    for a, g in zip(want, got):
        _same(a, g)
    # pinned inputs, resident re-submission and an empty batch
    ms[1].upload_reads(subs[2])
    ms[1].submit(None)
    _same(want[2], ms[1].wait())
    ms[2].submit(subs[4].pin())
    _same(want[4], ms[2].wait())
    empty = capi.ReadBatch(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    ms[0].submit(empty)
    assert len(ms[0].wait()["reads"]) == 0
    # misuse is an error, not a hang
    ms[0].submit(subs[0])
    with pytest.raises(capi.DartGpuError):
        ms[0].submit(subs[1])
    ms[0].wait()
    with pytest.raises(capi.DartGpuError):
        ms[0].wait()
    st = ms[0].stats()
    assert st["kernel_launches"] > 20 and st["ms_search"] > 0
    for m in ms:
        m.close()


def test_results_can_stay_on_the_device():
    """dartgpu_set_result_location(ctx, 1): the records stay in HBM (what bench.py's `value` times); copied back by hand they
    are the records the default mode delivers."""
    from cuda import cudart
    w = workload("c2")
    batch = _batch(w)
    m = capi.Mapper(w["idx"], device=0, pair_end=1)
    want = m.map_reads(batch)
    m.results_on_device(True)
    m.submit(batch)
    got = m.wait(copy="device")
    assert (got["n_reads"], got["n_reports"], got["n_cigar_bytes"], got["n_junctions"]) == \
           (len(want["reads"]), len(want["reports"]), len(want["cigars"]), len(want["junctions"]))

    def fetch(ptr, dtype, n):
        a = np.empty(n, dtype=dtype)
        if n:
            err, = cudart.cudaMemcpy(a.ctypes.data, ptr, a.nbytes, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost)
            assert err == cudart.cudaError_t.cudaSuccess
        return a
    back = dict(reads=fetch(got["dev_reads"], capi.READ_RESULT, got["n_reads"]), reports=fetch(got["dev_reports"], capi.REPORT, got["n_reports"]),
                cigars=fetch(got["dev_cigars"], np.uint8, got["n_cigar_bytes"]).tobytes(),
                junctions=fetch(got["dev_junctions"], capi.JUNCTION, got["n_junctions"]))
    _same(want, back)
    m.results_on_device(False)
    _same(want, m.map_reads(batch))
    m.close()
