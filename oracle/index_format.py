"""TEST INFRASTRUCTURE (oracle): a numpy restatement of the index files the reference's builder writes, for small genomes.

Follows /root/reference/src/BWT_Index/bwtindex.c:77-148 (bwa_idx_build: pack fwd+revcomp, BWT, Occ interleave
bwt_bwtupdate_core :53-75, SA sampling every 32 via bwt_cal_sa bwt.c:101-123) and the dump formats
bwt.c:174-196 (.bwt, .sa) and bntseq.c:192-205 (.pac).  The suffix array here is a plain prefix-doubling sort, nothing
like BWT-SW: the BWT of a text is unique, so any correct construction must give byte-identical files.  Pinned against
the files the reference's own bwt_index wrote (tests/golden/idx.*, tests/test_index_build.py).
Only tests may import this module; the product's builder is dart_b200/csrc/index_build.cu.
"""
import numpy as np


def suffix_array(text: np.ndarray) -> np.ndarray:
    """SA of text + '$' ($ smallest), n+1 entries, by prefix doubling."""
    n = len(text)
    rank = np.concatenate([text.astype(np.int64) + 1, [0]])
    sa = np.argsort(rank, kind="stable")
    k = 1
    while True:
        r2 = np.zeros(n + 1, dtype=np.int64)
        r2[: n + 1 - k] = rank[k:] + 1 if k <= n else 0
        order = np.lexsort((r2, rank))
        a, b = rank[order], r2[order]
        new = np.zeros(n + 1, dtype=np.int64)
        new[order] = np.concatenate([[0], np.cumsum((a[1:] != a[:-1]) | (b[1:] != b[:-1]))])
        rank, sa = new, order
        if rank.max() == n:
            return sa
        k *= 2


def pac_bytes(codes_fwd: np.ndarray) -> bytes:
    """.pac: 2 bits per base, first base in the top bits; then (a zero byte if l%4==0) and the byte l%4."""
    l = len(codes_fwd)
    pad = (-l) % 4
    c = np.concatenate([codes_fwd.astype(np.uint8), np.zeros(pad, np.uint8)]).reshape(-1, 4)
    packed = (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8).tobytes()
    return packed + (b"\0" if l % 4 == 0 else b"") + bytes([l % 4])


def index_files(codes_fwd: np.ndarray, sa_intv: int = 32) -> dict:
    """{'.bwt': bytes, '.sa': bytes, '.pac': bytes} for a forward strand given as codes 0..3."""
    fwd = codes_fwd.astype(np.uint8)
    text = np.concatenate([fwd, 3 - fwd[::-1]])
    n = len(text)
    sa = suffix_array(text)
    primary = int(np.nonzero(sa == 0)[0][0])
    rows = np.delete(sa, primary)                       # the BWT string has no symbol for the row of suffix 0
    bwt = text[rows - 1]
    counts = np.bincount(text, minlength=4)
    L2 = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint64)
    # 16 symbols per u32, first symbol in the top bits
    pad = (-n) % 16
    sym = np.concatenate([bwt, np.zeros(pad, np.uint8)]).astype(np.uint32).reshape(-1, 16)
    words = np.zeros(len(sym), dtype=np.uint32)
    for i in range(16):
        words |= sym[:, i] << np.uint32(30 - 2 * i)
    out = []
    cum = np.zeros(4, dtype=np.uint64)
    for b in range(0, n, 128):
        out.append(cum.copy().view(np.uint32))
        out.append(words[b // 16:(min(b + 128, n) + 15) // 16])
        cum += np.bincount(bwt[b:b + 128], minlength=4).astype(np.uint64)
    out.append(cum.copy().view(np.uint32))
    body = np.concatenate(out).astype("<u4").tobytes()
    head = np.array([primary], dtype="<u8").tobytes() + L2[1:].astype("<u8").tobytes()
    n_sa = (n + sa_intv) // sa_intv
    sampled = sa[::sa_intv][:n_sa].astype("<u8")
    sa_file = head + np.array([sa_intv, n], dtype="<u8").tobytes() + sampled[1:].tobytes()
    return {".bwt": head + body, ".sa": sa_file, ".pac": pac_bytes(fwd)}


def read_pac(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.uint8)
    l = (len(raw) - 2) * 4 + int(raw[-1]) if raw[-1] != 0 else (len(raw) - 2) * 4
    if raw[-1] == 0:
        l = (len(raw) - 2) * 4
    b = raw[: (l + 3) // 4]
    c = np.stack([(b >> 6) & 3, (b >> 4) & 3, (b >> 2) & 3, b & 3], axis=1).reshape(-1)
    return c[:l].astype(np.uint8)
