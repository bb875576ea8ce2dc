"""TEST INFRASTRUCTURE — ctypes bindings for the two checkers. Imported only by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.

  Oracle     oracle/liboracle.so       the CPU restatement (oracle/dart_oracle.cpp)
  Reference  oracle/_ref/libdartref.so the UNMODIFIED reference behind oracle/ref_taps.cpp
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

# nst_nt4_table (/root/reference/src/BWT_Index/bntseq.c:40-57) as the oracle restates it
NT4 = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate("ACGT"):
    NT4[ord(_c)] = _i
    NT4[ord(_c.lower())] = _i
NT4[ord("-")] = 5


def encode(seq: bytes | str) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode()
    return NT4[np.frombuffer(seq, dtype=np.uint8)]


def build(ref: bool = True) -> None:
    """(Re)build the checkers. `make ref` is a no-op where /root/reference is absent."""
    subprocess.run(["make", "-s", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True)


def have_reference() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libdartref.so"))


def build_index(fasta: str, prefix: str) -> None:
    """Run the reference's own index builder (oracle/_ref/bwt_index) unless the index exists."""
    if all(os.path.exists(prefix + e) for e in (".bwt", ".sa", ".pac", ".ann", ".amb")):
        return
    subprocess.run([os.path.join(REF_DIR, "bwt_index"), fasta, prefix], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


_u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "searches", "ext_steps", "ext_blocks", "lf_steps", "hits", "seeds", "read_bases",
        "nw_calls", "nw_cells", "kmer_calls", "kmer_window_bases", "kmer_read_bases", "lf_steps_rc")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class Oracle:
    def __init__(self, prefix: str):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        L = self.L = C.CDLL(path)
        L.or_load.restype = C.c_void_p
        L.or_load.argtypes = [C.c_char_p]
        L.or_free.argtypes = [C.c_void_p]
        L.or_genome_size.restype = C.c_int64
        L.or_genome_size.argtypes = [C.c_void_p]
        L.or_counters_get.argtypes = [C.c_void_p, C.POINTER(Counters)]
        L.or_counters_reset.argtypes = [C.c_void_p]
        L.or_ref_codes.argtypes = [C.c_void_p, C.c_int64, C.c_int, _u8p]
        L.or_rank4.argtypes = [C.c_void_p, C.c_uint64, _u64p]
        L.or_locate.restype = C.c_uint64
        L.or_locate.argtypes = [C.c_void_p, C.c_uint64]
        L.or_search.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), _u64p, C.c_int]
        L.or_search_mirrored.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), _u64p, C.c_int]
        L.or_seed_read.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, _i32p, _i64p, _i32p, C.c_int]
        L.or_cluster_read.argtypes = [C.c_void_p, C.c_int, C.c_int, _i32p, _i64p, _i32p, C.c_int, C.c_int,
                                      _i32p, _i32p, _i32p, C.c_int]
        L.or_kmer_pair.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, _i64p]
        L.or_nw.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int, C.c_char_p, _u8p]
        L.or_gapped_partition.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int,
                                          C.c_int, C.c_int64, C.c_int, _i32p]
        self.h = L.or_load(prefix.encode())
        if not self.h:
            raise RuntimeError(f"oracle: cannot load index {prefix}")
        self.G = L.or_genome_size(self.h)

    def close(self):
        if self.h:
            self.L.or_free(self.h)
            self.h = None

    def counters(self) -> dict:
        c = Counters()
        self.L.or_counters_get(self.h, C.byref(c))
        return c.as_dict()

    def reset_counters(self):
        self.L.or_counters_reset(self.h)

    def ref_codes(self, pos: int, n: int) -> np.ndarray:
        out = np.empty(n, dtype=np.uint8)
        self.L.or_ref_codes(self.h, pos, n, out)
        return out

    def rank4(self, k: int) -> np.ndarray:
        out = np.zeros(4, dtype=np.uint64)
        self.L.or_rank4(self.h, C.c_uint64(k & 0xFFFFFFFFFFFFFFFF), out)
        return out

    def locate(self, k: int) -> int:
        return int(self.L.or_locate(self.h, k))

    def search(self, codes: np.ndarray, start: int, stop: int, max_dup: int = 100):
        ln = C.c_int(0)
        locs = np.zeros(max(max_dup, 1), dtype=np.uint64)
        f = self.L.or_search(self.h, np.ascontiguousarray(codes, np.uint8), start, stop, max_dup, C.byref(ln), locs, len(locs))
        return ln.value, f, locs[:f].copy()

    def search_mirrored(self, codes: np.ndarray, start: int, stop: int, max_dup: int = 100):
        ln = C.c_int(0)
        locs = np.zeros(max(max_dup, 1), dtype=np.uint64)
        f = self.L.or_search_mirrored(self.h, np.ascontiguousarray(codes, np.uint8), start, stop, max_dup, C.byref(ln), locs, len(locs))
        return ln.value, f, locs[:f].copy()

    def seeds(self, codes: np.ndarray, max_dup: int = 100, cap: int = 1 << 16):
        r = np.empty(cap, np.int32); g = np.empty(cap, np.int64); l = np.empty(cap, np.int32)
        n = self.L.or_seed_read(self.h, np.ascontiguousarray(codes, np.uint8), len(codes), max_dup, r, g, l, cap)
        assert n <= cap
        return r[:n].copy(), g[:n].copy(), l[:n].copy()

    def cluster(self, rlen, r, g, l, max_gaps=5, max_intron=500000, cap: int = 1 << 14):
        b = np.empty(cap, np.int32); c = np.empty(cap, np.int32); s = np.empty(cap, np.int32)
        n = self.L.or_cluster_read(self.h, rlen, len(r), np.ascontiguousarray(r, np.int32),
                                   np.ascontiguousarray(g, np.int64), np.ascontiguousarray(l, np.int32),
                                   max_gaps, max_intron, b, c, s, cap)
        assert n <= cap
        return b[:n].copy(), c[:n].copy(), s[:n].copy()

    def kmer_pair(self, f1: bytes, f2: bytes):
        out = np.zeros(3, np.int64)
        self.L.or_kmer_pair(self.h, len(f1), f1, len(f2), f2, out)
        return tuple(int(x) for x in out)

    def nw(self, s1: bytes, s2: bytes) -> np.ndarray:
        ops = np.zeros(len(s1) + len(s2) + 2, np.uint8)
        k = self.L.or_nw(self.h, len(s1), s1, len(s2), s2, ops)
        return ops[:k].copy()

    def gapped_partition(self, seq: bytes, rgaps, l_rpos, l_rlen, l_gpos, l_glen, r_rpos, r_gpos, max_mismatch):
        out = np.zeros(3, np.int32)
        self.L.or_gapped_partition(self.h, seq, rgaps, l_rpos, l_rlen, l_gpos, l_glen, r_rpos, r_gpos, max_mismatch, out)
        return tuple(int(x) for x in out)


def ops_to_strings(s1: bytes, s2: bytes, ops) -> tuple[bytes, bytes]:
    a, b, i, j = bytearray(), bytearray(), 0, 0
    for o in ops:
        if o == 0:
            a.append(s1[i]); b.append(s2[j]); i += 1; j += 1
        elif o == 1:
            a.append(45); b.append(s2[j]); j += 1
        else:
            a.append(s1[i]); b.append(45); i += 1
    return bytes(a), bytes(b)


_REFERENCE_PREFIX = None
_KEEP = []


class Reference:
    """The reference's own functions (oracle/_ref/libdartref.so). One index per process: the reference keeps the
    index, ChrLocMap and its parameters in globals that are never cleared."""

    def __init__(self, prefix: str, threads: int = 1):
        global _REFERENCE_PREFIX
        if _REFERENCE_PREFIX not in (None, prefix):
            raise RuntimeError(f"libdartref.so already holds {_REFERENCE_PREFIX}; use a fresh process for {prefix}")
        reload_needed = _REFERENCE_PREFIX is None
        _REFERENCE_PREFIX = prefix
        L = self.L = C.CDLL(os.path.join(REF_DIR, "libdartref.so"))
        L.ref_load.argtypes = [C.c_char_p, C.c_int]
        L.ref_genome_size.restype = C.c_int64
        L.ref_sa.restype = C.c_uint64
        L.ref_sa.argtypes = [C.c_uint64]
        L.ref_2occ4.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u64p]
        L.ref_bwt_search.argtypes = [_u8p, C.c_int, C.c_int, C.POINTER(C.c_int), _u64p, C.c_int]
        L.ref_identify_seed_pairs.argtypes = [C.c_int, _u8p, _i32p, _i64p, _i32p, C.c_int]
        L.ref_candidates.argtypes = [C.c_int, _u8p, _i32p, _i64p, _i32p, C.c_int, _i32p, _i64p, _i32p, C.c_int,
                                     C.POINTER(C.c_int)]
        L.ref_kmer_pair.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, _i64p]
        L.ref_nw.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int]
        L.ref_gapped_partition.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int,
                                           C.c_int64, _i32p]
        L.ref_sj_new.restype = C.c_void_p
        L.ref_sj_free.argtypes = [C.c_void_p]
        L.ref_sj_size.argtypes = [C.c_void_p]
        L.ref_sj_dump.argtypes = [C.c_void_p, _i64p, _i64p, _i32p, _i32p]
        L.ref_map_single.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ref_map_pair.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_char_p,
                                   C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_void_p, C.c_char_p, C.c_int]
        L.ref_sequence.restype = C.c_void_p
        L.ref_hotpath.restype = C.c_double
        L.ref_hotpath.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
        self._prefix = prefix.encode()  # the library keeps the pointer
        if reload_needed:
            _KEEP.append(self._prefix)
            rc = L.ref_load(self._prefix, threads)
            if rc != 0:
                raise RuntimeError(f"reference: cannot load index {prefix} (rc={rc})")
        self.G = L.ref_genome_size()
        self.params = dict(max_mismatch=0, max_dup=100, max_intron=500000, min_intron=5, multi_hit=0, pair_end=0,
                           all_sj=0, unique=0)

    def set_params(self, **kw):
        self.params.update(kw)
        p = self.params
        self.L.ref_set_params(p["max_mismatch"], p["max_dup"], p["max_intron"], p["min_intron"], p["multi_hit"],
                              p["pair_end"], p["all_sj"], p["unique"])

    def ref_chars(self, pos: int, n: int) -> bytes:
        return C.string_at(self.L.ref_sequence() + pos, n)

    def occ2(self, k: int, l: int):
        a = np.zeros(4, np.uint64); b = np.zeros(4, np.uint64)
        self.L.ref_2occ4(k, l, a, b)
        return a, b

    def sa(self, k: int) -> int:
        return int(self.L.ref_sa(k))

    def search(self, codes, start, stop):
        ln = C.c_int(0)
        locs = np.zeros(10000, np.uint64)
        f = self.L.ref_bwt_search(np.ascontiguousarray(codes, np.uint8), start, stop, C.byref(ln), locs, len(locs))
        return ln.value, f, locs[:f].copy()

    def seeds(self, codes, cap: int = 1 << 16):
        r = np.empty(cap, np.int32); g = np.empty(cap, np.int64); l = np.empty(cap, np.int32)
        n = self.L.ref_identify_seed_pairs(len(codes), np.ascontiguousarray(codes, np.uint8), r, g, l, cap)
        assert n <= cap
        return r[:n].copy(), g[:n].copy(), l[:n].copy()

    def candidates(self, codes, cap_c: int = 1 << 14, cap_s: int = 1 << 16):
        cs = np.empty(cap_c, np.int32); cp = np.empty(cap_c, np.int64); cn = np.empty(cap_c, np.int32)
        r = np.empty(cap_s, np.int32); g = np.empty(cap_s, np.int64); l = np.empty(cap_s, np.int32)
        ns = C.c_int(0)
        nc = self.L.ref_candidates(len(codes), np.ascontiguousarray(codes, np.uint8), cs, cp, cn, cap_c, r, g, l, cap_s,
                                   C.byref(ns))
        assert nc <= cap_c and ns.value <= cap_s
        return cs[:nc].copy(), cp[:nc].copy(), cn[:nc].copy(), r[:ns.value].copy(), g[:ns.value].copy(), l[:ns.value].copy()

    def kmer_pair(self, f1: bytes, f2: bytes):
        out = np.zeros(3, np.int64)
        self.L.ref_kmer_pair(len(f1), f1, len(f2), f2, out)
        return tuple(int(x) for x in out)

    def nw(self, s1: bytes, s2: bytes):
        cap = len(s1) + len(s2) + 2
        o1 = C.create_string_buffer(cap); o2 = C.create_string_buffer(cap)
        self.L.ref_nw(len(s1), s1, len(s2), s2, o1, o2, cap)
        return o1.value, o2.value

    def gapped_partition(self, seq: bytes, rgaps, l_rpos, l_rlen, l_gpos, l_glen, r_rpos, r_gpos):
        out = np.zeros(3, np.int32)
        self.L.ref_gapped_partition(seq, rgaps, l_rpos, l_rlen, l_gpos, l_glen, r_rpos, r_gpos, out)
        return tuple(int(x) for x in out)

    def hotpath(self, bases, offsets, paired: int, threads: int) -> float:
        """Seconds the reference's per-read functions need for a pre-parsed batch on `threads` threads (no parse / format / IO)."""
        b = np.ascontiguousarray(bases, np.uint8); o = np.ascontiguousarray(offsets, np.int64)
        m = C.c_long(0)
        sec = self.L.ref_hotpath(b.ctypes.data, o.ctypes.data, len(o) - 1, int(paired), int(threads), C.byref(m))
        self.hotpath_mapped = m.value
        return float(sec)

    def map_single(self, name: bytes, seq: bytes, qual: bytes, sj=None) -> bytes:
        out = C.create_string_buffer(1 << 20)
        n = self.L.ref_map_single(name, seq, len(seq), qual, sj, out, len(out))
        assert n >= 0
        return out.raw[:n]

    def map_pair(self, name1, seq1, qual1, name2, seq2, qual2, sj=None) -> bytes:
        out = C.create_string_buffer(1 << 22)
        n = self.L.ref_map_pair(name1, seq1, len(seq1), qual1, name2, seq2, len(seq2), qual2, sj, out, len(out))
        assert n >= 0
        return out.raw[:n]
