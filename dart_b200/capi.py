"""ctypes binding of include/dartgpu.h — the host-side mirror of the reference's per-read interface.

The names follow the reference functions each call replaces (src/structure.h:192-233):

    Mapper.identify_seed_pairs(reads)      IdentifySeedPairs + GenerateAlignmentCandidate, batched
    Mapper.kmer_reseed(bases, jobs)        GenerateLongestSimplePairsFromFragmentPair, batched
    Mapper.nw_alignment(bases, jobs)       nw_alignment, batched
    Mapper.map_reads(reads)                the body of ReadMapping()'s per-read loop

There is no CPU path here: if libdartgpu.so is missing or no CUDA device is usable, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DARTGPU_LIB", os.path.join(HERE, "libdartgpu.so"))   # DARTGPU_LIB: A/B runs of two builds


class DartGpuError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dartgpu error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("max_gaps", C.c_int32), ("max_intron", C.c_int32), ("min_intron", C.c_int32),
                ("max_mismatch", C.c_int32), ("max_dup", C.c_uint32), ("multi_hit", C.c_int32),
                ("pair_end", C.c_int32), ("all_sj", C.c_int32), ("unique", C.c_int32), ("host_threads", C.c_int32)]


class _Reads(C.Structure):
    _fields_ = [("n_reads", C.c_int32), ("bases", C.c_void_p), ("offsets", C.c_void_p)]


class _Seeds(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("seed_off", "seed_gpos", "seed_rpos", "seed_len",
                                           "cand_off", "cand_begin", "cand_count", "cand_score")]


class _NwResult(C.Structure):
    _fields_ = [("op_off", C.c_void_p), ("ops", C.c_void_p)]


class _MapResult(C.Structure):
    _fields_ = [("reads", C.c_void_p), ("n_reads", C.c_int32), ("reports", C.c_void_p), ("n_reports", C.c_int64),
                ("cigars", C.c_void_p), ("n_cigar_bytes", C.c_int64), ("junctions", C.c_void_p), ("n_junctions", C.c_int64)]


class _FastqBlock(C.Structure):
    _fields_ = [("text1", C.c_void_p), ("len1", C.c_int64), ("text2", C.c_void_p), ("len2", C.c_int64), ("n_records", C.c_int32),
                ("fastq", C.c_int32), ("max_read_len", C.c_int32), ("reserved", C.c_int32)]


class _SamResult(C.Structure):
    _fields_ = [("sam", C.c_void_p), ("n_bytes", C.c_int64), ("n_reads", C.c_int64), ("n_unmapped", C.c_int64), ("n_unique", C.c_int64),
                ("n_paired", C.c_int64), ("junctions", C.c_void_p), ("n_junctions", C.c_int64)]


class Stats(C.Structure):
    _fields_ = ([(n, C.c_double) for n in ("ms_search", "ms_locate", "ms_sort_cluster", "ms_kmer", "ms_nw", "ms_h2d",
                                           "ms_d2h", "ms_total_device", "ms_host", "ms_report")] +
                [(n, C.c_uint64) for n in ("kernel_launches", "ext_steps", "ext_blocks", "lf_steps", "hits", "seeds",
                                           "read_bases", "nw_jobs", "nw_cells", "kmer_jobs", "kmer_window_bases",
                                           "kmer_read_bases", "h2d_bytes", "d2h_bytes", "search_sector_loads")] +
                [("ms_submit", C.c_double)])

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


KMER_JOB = np.dtype([("frag_off", "<i8"), ("frag_len", "<i4"), ("glen", "<i4"), ("gpos", "<i8")])
KMER_HIT = np.dtype([("rpos", "<i4"), ("gpos", "<i4"), ("len", "<i4")])
NW_JOB = np.dtype([("frag_off", "<i8"), ("m", "<i4"), ("n", "<i4"), ("gpos", "<i8")])
READ_RESULT = np.dtype([("report_off", "<i8"), ("n_reports", "<i4"), ("best", "<i4"), ("score", "<i2"), ("sub_score", "<i2"),
                        ("mis_num", "<i2"), ("mapq", "u1"), ("reserved", "u1")])
REPORT = np.dtype([("pos", "<i8"), ("cigar_off", "<i4"), ("flag", "<i4"), ("paired_idx", "<i4"), ("chr_idx", "<i4"),
                   ("aln_score", "<i2"), ("cigar_len", "<i2"), ("sj_type", "i1"), ("dir", "u1"), ("reserved", "u1", (2,))])
JUNCTION = np.dtype([("g1", "<i8"), ("g2", "<i8"), ("type", "<i4"), ("read", "<i4")])

# every symbol include/dartgpu.h declares (tests check the library exports all of them)
EXPORTS = ["dartgpu_default_params", "dartgpu_create", "dartgpu_create_from_files", "dartgpu_destroy",
           "dartgpu_set_params", "dartgpu_last_error", "dartgpu_genome_size", "dartgpu_num_sequences",
           "dartgpu_sequence_name", "dartgpu_sequence_length", "dartgpu_set_stream", "dartgpu_seed_and_cluster",
           "dartgpu_kmer_reseed", "dartgpu_nw_align", "dartgpu_map_reads", "dartgpu_get_stats",
           "dartgpu_upload_reads", "dartgpu_seed_and_cluster_resident", "dartgpu_synchronize",
           "dartgpu_map_reads_resident", "dartgpu_index_build", "dartgpu_measure_int32_peak", "dartgpu_measure_l2_peak",
           "dartgpu_submit", "dartgpu_submit_resident", "dartgpu_wait", "dartgpu_submit_fastq", "dartgpu_wait_sam",
           "dartgpu_fastq_cut", "dartgpu_alloc_pinned", "dartgpu_free_pinned", "dartgpu_set_result_location",
           "dartgpu_bind_host_thread"]


def load_library() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise DartGpuError(-1, f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                               "there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    L.dartgpu_default_params.argtypes = [C.POINTER(Params)]
    L.dartgpu_create_from_files.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.POINTER(Params)]
    L.dartgpu_destroy.argtypes = [C.c_void_p]
    L.dartgpu_index_build.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_char_p, C.c_uint64]
    L.dartgpu_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.dartgpu_last_error.restype = C.c_char_p
    L.dartgpu_last_error.argtypes = [C.c_void_p]
    L.dartgpu_genome_size.restype = C.c_int64
    L.dartgpu_genome_size.argtypes = [C.c_void_p]
    L.dartgpu_num_sequences.argtypes = [C.c_void_p]
    L.dartgpu_sequence_name.restype = C.c_char_p
    L.dartgpu_sequence_name.argtypes = [C.c_void_p, C.c_int]
    L.dartgpu_sequence_length.restype = C.c_int64
    L.dartgpu_sequence_length.argtypes = [C.c_void_p, C.c_int]
    L.dartgpu_set_stream.argtypes = [C.c_void_p, C.c_void_p]
    L.dartgpu_set_result_location.argtypes = [C.c_void_p, C.c_int]
    L.dartgpu_bind_host_thread.argtypes = [C.c_int]
    L.dartgpu_seed_and_cluster.argtypes = [C.c_void_p, C.POINTER(_Reads), C.POINTER(_Seeds)]
    L.dartgpu_kmer_reseed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]
    L.dartgpu_nw_align.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32, C.POINTER(_NwResult)]
    L.dartgpu_map_reads.argtypes = [C.c_void_p, C.POINTER(_Reads), C.POINTER(_MapResult)]
    L.dartgpu_map_reads_resident.argtypes = [C.c_void_p, C.POINTER(_Reads), C.POINTER(_MapResult)]
    L.dartgpu_submit.argtypes = [C.c_void_p, C.POINTER(_Reads)]
    L.dartgpu_submit_resident.argtypes = [C.c_void_p]
    L.dartgpu_wait.argtypes = [C.c_void_p, C.POINTER(_MapResult)]
    L.dartgpu_submit_fastq.argtypes = [C.c_void_p, C.POINTER(_FastqBlock)]
    L.dartgpu_wait_sam.argtypes = [C.c_void_p, C.POINTER(_SamResult)]
    L.dartgpu_fastq_cut.restype = C.c_int64
    L.dartgpu_fastq_cut.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_int32)]
    L.dartgpu_measure_int32_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
    L.dartgpu_measure_l2_peak.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_double)]
    L.dartgpu_get_stats.argtypes = [C.c_void_p, C.POINTER(Stats)]
    L.dartgpu_upload_reads.argtypes = [C.c_void_p, C.POINTER(_Reads)]
    L.dartgpu_seed_and_cluster_resident.argtypes = [C.c_void_p]
    L.dartgpu_synchronize.argtypes = [C.c_void_p]
    return L


def _view(ptr, dtype, n):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (int(n) * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(n))


@dataclass
class ReadBatch:
    """Concatenated ASCII bases + offsets, as ReadItem_t.seq holds them (mate 2 already reverse-complemented)."""
    bases: np.ndarray    # uint8
    offsets: np.ndarray  # int64, n+1

    @staticmethod
    def from_list(seqs) -> "ReadBatch":
        lens = np.fromiter((len(s) for s in seqs), dtype=np.int64, count=len(seqs))
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        bases = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if len(seqs) else np.zeros(0, np.uint8)
        return ReadBatch(bases, off)

    @staticmethod
    def from_codes(codes: np.ndarray) -> "ReadBatch":
        """(n, L) array of codes 0..3 -> batch of equal-length reads."""
        n, L = codes.shape
        bases = np.frombuffer(b"ACGT", dtype=np.uint8)[codes].reshape(-1).copy()
        return ReadBatch(bases, np.arange(n + 1, dtype=np.int64) * L)

    @property
    def n(self) -> int:
        return len(self.offsets) - 1

    def pin(self) -> "ReadBatch":
        """The same batch in page-locked host memory (torch's pinned allocator): dartgpu_map_reads hands page-locked
        caller buffers straight to the DMA engine instead of staging them through its own pinned buffer."""
        import torch
        tb = torch.from_numpy(np.ascontiguousarray(self.bases, dtype=np.uint8)).pin_memory()
        to = torch.from_numpy(np.ascontiguousarray(self.offsets, dtype=np.int64)).pin_memory()
        out = ReadBatch(tb.numpy(), to.numpy())
        out._keep = (tb, to)
        return out

    def _c(self) -> _Reads:
        self.bases = np.ascontiguousarray(self.bases, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(self.offsets, dtype=np.int64)
        return _Reads(self.n, self.bases.ctypes.data, self.offsets.ctypes.data)


class Mapper:
    """One context = one host thread on one GPU, index resident in that GPU's HBM."""

    def __init__(self, index_prefix: str, device: int = 0, **params):
        self.L = load_library()
        self.params = Params()
        self.L.dartgpu_default_params(C.byref(self.params))
        for k, v in params.items():
            setattr(self.params, k, v)
        self.h = C.c_void_p()
        rc = self.L.dartgpu_create_from_files(C.byref(self.h), device, index_prefix.encode(), C.byref(self.params))
        if rc != 0:
            raise DartGpuError(rc, self.L.dartgpu_last_error(None).decode())
        self.G = self.L.dartgpu_genome_size(self.h)

    def close(self):
        if getattr(self, "h", None):
            self.L.dartgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise DartGpuError(rc, self.L.dartgpu_last_error(self.h).decode())

    def set_params(self, **params):
        for k, v in params.items():
            setattr(self.params, k, v)
        self._check(self.L.dartgpu_set_params(self.h, C.byref(self.params)))

    def results_on_device(self, on: bool):
        """Leave the records in HBM (device pointers in the result) instead of copying them to the host."""
        self._check(self.L.dartgpu_set_result_location(self.h, int(bool(on))))

    def set_stream(self, cuda_stream_ptr: int):
        self._check(self.L.dartgpu_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sequences(self):
        return [(self.L.dartgpu_sequence_name(self.h, i).decode(), self.L.dartgpu_sequence_length(self.h, i))
                for i in range(self.L.dartgpu_num_sequences(self.h))]

    def stats(self) -> dict:
        s = Stats()
        self.L.dartgpu_get_stats(self.h, C.byref(s))
        return s.as_dict()

    def int32_peak(self) -> float:
        """Measured INT32 add/max operations per second of this GPU (the NW kernels' roofline denominator)."""
        v = C.c_double(0)
        self._check(self.L.dartgpu_measure_int32_peak(self.h, C.byref(v)))
        return float(v.value)

    def l2_peak(self, table_bytes: int = 16 << 20) -> float:
        """Measured bytes/s of random one-sector gathers from an L2-resident table (k_search's roof on a small index)."""
        v = C.c_double(0)
        self._check(self.L.dartgpu_measure_l2_peak(self.h, int(table_bytes), C.byref(v)))
        return float(v.value)

    # ---- IdentifySeedPairs + GenerateAlignmentCandidate ----
    def identify_seed_pairs(self, reads: ReadBatch) -> dict:
        out = _Seeds()
        self._check(self.L.dartgpu_seed_and_cluster(self.h, C.byref(reads._c()), C.byref(out)))
        n = reads.n
        seed_off = _view(out.seed_off, np.int64, n + 1).copy()
        cand_off = _view(out.cand_off, np.int64, n + 1).copy()
        ns, nc = int(seed_off[-1]), int(cand_off[-1])
        return dict(seed_off=seed_off, cand_off=cand_off,
                    seed_gpos=_view(out.seed_gpos, np.int64, ns).copy(), seed_rpos=_view(out.seed_rpos, np.int32, ns).copy(),
                    seed_len=_view(out.seed_len, np.int32, ns).copy(), cand_begin=_view(out.cand_begin, np.int32, nc).copy(),
                    cand_count=_view(out.cand_count, np.int32, nc).copy(), cand_score=_view(out.cand_score, np.int32, nc).copy())

    def upload_reads(self, reads: ReadBatch):
        self._check(self.L.dartgpu_upload_reads(self.h, C.byref(reads._c())))

    def seed_resident(self):
        self._check(self.L.dartgpu_seed_and_cluster_resident(self.h))

    def synchronize(self):
        self._check(self.L.dartgpu_synchronize(self.h))

    # ---- GenerateLongestSimplePairsFromFragmentPair ----
    def kmer_reseed(self, bases: bytes, jobs) -> np.ndarray:
        """jobs: iterable of (frag_off, frag_len, gpos, glen). Returns KMER_HIT records (rpos, gpos, len)."""
        arr = np.array([(o, l, gl, g) for (o, l, g, gl) in jobs], dtype=KMER_JOB)
        b = np.frombuffer(bases, dtype=np.uint8)
        out = C.c_void_p()
        self._check(self.L.dartgpu_kmer_reseed(self.h, b.ctypes.data if len(b) else None, len(b),
                                               arr.ctypes.data if len(arr) else None, len(arr), C.byref(out)))
        return _view(out.value, KMER_HIT, len(arr)).copy()

    # ---- nw_alignment ----
    def nw_alignment(self, bases: bytes, jobs) -> list:
        """jobs: iterable of (frag_off, m, gpos, n). Returns one uint8 op array per job (0 both, 1 gap in read, 2 gap in genome)."""
        arr = np.array([(o, m, n, g) for (o, m, g, n) in jobs], dtype=NW_JOB)
        b = np.frombuffer(bases, dtype=np.uint8)
        out = _NwResult()
        self._check(self.L.dartgpu_nw_align(self.h, b.ctypes.data if len(b) else None, len(b),
                                            arr.ctypes.data if len(arr) else None, len(arr), C.byref(out)))
        off = _view(out.op_off, np.int64, len(arr) + 1).copy()
        ops = _view(out.ops, np.uint8, int(off[-1]) if len(off) else 0).copy()
        return [ops[off[i]:off[i + 1]] for i in range(len(arr))]

    # ---- the per-read loop body of ReadMapping ----
    def submit(self, reads: ReadBatch = None):
        """Enqueue a batch and return at once (reads=None: the batch uploaded with upload_reads). One batch per context."""
        if reads is None:
            self._check(self.L.dartgpu_submit_resident(self.h))
        else:
            self._held = reads      # page-locked caller buffers are read by the DMA engine until wait()
            self._check(self.L.dartgpu_submit(self.h, C.byref(reads._c())))

    def wait(self, copy: bool = True) -> dict:
        out = _MapResult()
        self._check(self.L.dartgpu_wait(self.h, C.byref(out)))
        return self._result(out, copy)

    # ---- FASTQ text in, SAM text out (device-side ingest and formatting) ----
    def submit_fastq(self, text1, text2=None, n_records=None, max_read_len=0):
        """text1 / text2: bytes-like or uint8 arrays (page-locked arrays are DMA'd as they are). Returns at once."""
        a1 = np.frombuffer(text1, dtype=np.uint8) if not isinstance(text1, np.ndarray) else text1
        a2 = None if text2 is None else (np.frombuffer(text2, dtype=np.uint8) if not isinstance(text2, np.ndarray) else text2)
        if n_records is None:
            k = C.c_int32(0)
            used = self.L.dartgpu_fastq_cut(a1.ctypes.data, len(a1), 0, C.byref(k))
            assert used == len(a1), "text1 does not end at a record boundary"
            n_records = k.value
        self._held = (a1, a2)
        b = _FastqBlock(a1.ctypes.data if len(a1) else None, len(a1), a2.ctypes.data if a2 is not None and len(a2) else None,
                        len(a2) if a2 is not None else 0, n_records, 1, max_read_len, 0)
        self._check(self.L.dartgpu_submit_fastq(self.h, C.byref(b)))

    def wait_sam(self) -> dict:
        out = _SamResult()
        self._check(self.L.dartgpu_wait_sam(self.h, C.byref(out)))
        return dict(sam=_view(out.sam, np.uint8, out.n_bytes).tobytes(), n_reads=out.n_reads, n_unmapped=out.n_unmapped,
                    n_unique=out.n_unique, n_paired=out.n_paired, junctions=_view(out.junctions, JUNCTION, out.n_junctions).copy())

    def map_reads(self, reads: ReadBatch, resident: bool = False, copy: bool = True) -> dict:
        out = _MapResult()
        fn = self.L.dartgpu_map_reads_resident if resident else self.L.dartgpu_map_reads
        self._check(fn(self.h, C.byref(reads._c()), C.byref(out)))
        return self._result(out, copy)

    @staticmethod
    def _result(out, copy):
        if copy == "device":   # results left in HBM: counts only
            return dict(n_reads=out.n_reads, n_reports=out.n_reports, n_cigar_bytes=out.n_cigar_bytes, n_junctions=out.n_junctions,
                        dev_reads=out.reads, dev_reports=out.reports, dev_cigars=out.cigars, dev_junctions=out.junctions)
        if not copy:  # views into context-owned memory, valid until the next call
            return dict(reads=_view(out.reads, READ_RESULT, out.n_reads), reports=_view(out.reports, REPORT, out.n_reports),
                        n_cigar_bytes=out.n_cigar_bytes, junctions=_view(out.junctions, JUNCTION, out.n_junctions))
        return dict(reads=_view(out.reads, READ_RESULT, out.n_reads).copy(),
                    reports=_view(out.reports, REPORT, out.n_reports).copy(),
                    cigars=_view(out.cigars, np.uint8, out.n_cigar_bytes).tobytes(),
                    junctions=_view(out.junctions, JUNCTION, out.n_junctions).copy())


def bind_host_thread(device: int) -> bool:
    """Pin this thread to the CPUs of the GPU's NUMA node (before creating Mappers: their pinned buffers follow)."""
    return load_library().dartgpu_bind_host_thread(int(device)) == 0


def index_build(genome, prefix: str, device: int = 0, max_suffixes_per_pass: int = 0) -> None:
    """`dart index` / bwt_index on the GPU: writes <prefix>.pac/.ann/.amb (the packing bns_fasta2bntseq does,
    src/BWT_Index/bntseq.c:59-89, :158-211 — host side, linear) and <prefix>.bwt/.sa (dartgpu_index_build).
    `genome` is a dart_b200.synth.Genome (names + uint8 code arrays, ACGT only)."""
    from . import synth
    pac = synth.write_index_meta(prefix, genome)
    L = load_library()
    buf = np.ascontiguousarray(pac)
    rc = L.dartgpu_index_build(device, buf.ctypes.data, int(genome.total_len), prefix.encode(), int(max_suffixes_per_pass))
    if rc != 0:
        raise DartGpuError(rc, (L.dartgpu_last_error(None) or b"").decode())
