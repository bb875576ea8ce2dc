"""Seeded synthetic genomes and reads of the shapes BASELINE.json names.

There is no network and the reference ships no data, so every workload is generated:
  * random genomes (1..n contigs), optionally repeat-rich (segmental duplications, config 5);
  * gene models with canonical splice motifs planted on both strands (SURVEY.md F6 — the
    reference only records a junction when GT/AG, CT/AC, GC/AG or CT/GC is found,
    /root/reference/src/AlignmentCandidates.cpp:732-756);
  * single-end / paired-end reads with substitutions and short indels, written as FASTQ with
    fixed-width names so that a million records are produced with numpy only.

Bases are handled as uint8 codes 0..3 = A,C,G,T (the order of nst_nt4_table,
/root/reference/src/BWT_Index/bntseq.c:40-57).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)


def revcomp_codes(a: np.ndarray) -> np.ndarray:
    return (3 - a)[..., ::-1]


# ----------------------------------------------------------------------------------------------
# genomes
# ----------------------------------------------------------------------------------------------
@dataclass
class Genome:
    names: list[str]
    seqs: list[np.ndarray]  # uint8 codes
    genes: list["Gene"] = field(default_factory=list)

    @property
    def total_len(self) -> int:
        return int(sum(len(s) for s in self.seqs))


@dataclass
class Gene:
    contig: int
    strand: int  # +1 / -1
    exons: list[tuple[int, int]]  # [start, end) on the contig, ascending


def random_genome(total_len: int, n_contigs: int = 1, seed: int = 1001) -> Genome:
    rng = np.random.default_rng(seed)
    if n_contigs == 1:
        lens = [total_len]
    else:
        w = rng.uniform(0.5, 1.5, n_contigs)
        lens = np.maximum(1000, (w / w.sum() * total_len).astype(np.int64)).tolist()
    seqs = [rng.integers(0, 4, size=int(n), dtype=np.uint8) for n in lens]
    names = [f"chr{i + 1}" for i in range(n_contigs)]
    return Genome(names, seqs)


def add_segmental_duplications(g: Genome, fraction: float = 0.30, seed: int = 1005,
                               seg_min: int = 5000, seg_max: int = 50000) -> None:
    """Overwrite `fraction` of each contig with diverged (0.2-1 %) copies of other segments."""
    rng = np.random.default_rng(seed)
    for s in g.seqs:
        n = len(s)
        covered, target = 0, int(n * fraction)
        hi = min(seg_max, max(seg_min + 1, n // 8))
        lo = min(seg_min, hi - 1)
        while covered < target:
            L = int(rng.integers(lo, hi))
            src = int(rng.integers(0, n - L))
            dst = int(rng.integers(0, n - L))
            seg = s[src:src + L].copy()
            if rng.random() < 0.5:
                seg = revcomp_codes(seg).copy()
            div = rng.uniform(0.002, 0.01)
            m = rng.random(L) < div
            seg[m] = (seg[m] + rng.integers(1, 4, size=int(m.sum()), dtype=np.uint8)) & 3
            s[dst:dst + L] = seg
            covered += L


def add_gene_models(g: Genome, genes_per_mbp: float = 8.0, seed: int = 1003,
                    max_intron: int = 500000, min_intron: int = 70) -> None:
    """Lay out non-overlapping genes; plant GT..AG (some GC..AG) on + genes and CT..AC (CT..GC) on - genes."""
    rng = np.random.default_rng(seed)
    for ci, s in enumerate(g.seqs):
        n = len(s)
        pos = 1000
        n_target = max(1, int(n / 1e6 * genes_per_mbp))
        made = 0
        while made < n_target and pos < n - 5000:
            n_ex = int(rng.integers(2, 9))
            strand = 1 if rng.random() < 0.5 else -1
            exons = []
            p = pos
            ok = True
            for e in range(n_ex):
                el = int(rng.integers(50, 501))
                if p + el + 100 >= n:
                    ok = False
                    break
                exons.append((p, p + el))
                p += el
                if e + 1 < n_ex:
                    cap = min(max_intron, max(min_intron + 1, (n - p) // (n_ex - e)))
                    il = int(np.exp(rng.uniform(np.log(min_intron), np.log(cap))))
                    il = max(min_intron, min(il, n - p - 600))
                    if il < min_intron:
                        ok = False
                        break
                    gc = rng.random() < 0.05
                    if strand > 0:
                        s[p:p + 2] = (2, 1) if gc else (2, 3)       # GC / GT donor
                        s[p + il - 2:p + il] = (0, 2)               # AG acceptor
                    else:
                        s[p:p + 2] = (1, 3)                         # CT  (revcomp of AG)
                        s[p + il - 2:p + il] = (2, 1) if gc else (0, 1)  # GC / AC (revcomp of GC / GT)
                    p += il
            if ok and len(exons) >= 2:
                g.genes.append(Gene(ci, strand, exons))
                made += 1
            pos = p + int(rng.integers(500, 5000))


def write_fasta(path: str, g: Genome, width: int = 80) -> None:
    with open(path, "wb") as f:
        for name, s in zip(g.names, g.seqs):
            f.write(b">" + name.encode() + b"\n")
            chars = ALPHABET[s]
            full = (len(chars) // width) * width
            if full:
                block = np.empty((full // width, width + 1), dtype=np.uint8)
                block[:, :width] = chars[:full].reshape(-1, width)
                block[:, width] = 10
                f.write(block.tobytes())
            if full < len(chars):
                f.write(chars[full:].tobytes() + b"\n")


def pack_2bit(codes: np.ndarray) -> np.ndarray:
    """4 bases per byte, first base in the top bits (_set_pac, src/BWT_Index/bntseq.c)."""
    pad = (-len(codes)) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, np.uint8)])
    c = codes.reshape(-1, 4)
    return (c[:, 0] << 6 | c[:, 1] << 4 | c[:, 2] << 2 | c[:, 3]).astype(np.uint8)


def write_index_meta(prefix: str, g: Genome) -> np.ndarray:
    """<prefix>.pac / .ann / .amb as the reference's bns_fasta2bntseq(for_only=1) + bns_dump write them
    (/root/reference/src/BWT_Index/bntseq.c:59-89, :192-205) for an ACGT-only genome. Returns the packed forward strand."""
    l = g.total_len
    # contigs are packed back to back: pack the concatenation in slices that start on a multiple of 4 bases
    packed = np.empty((l + 3) // 4 + 1, dtype=np.uint8)
    flat = np.concatenate(g.seqs) if len(g.seqs) > 1 else g.seqs[0]
    step = 1 << 28
    for a in range(0, l, step):
        b = min(l, a + step)
        packed[a // 4:(b + 3) // 4] = pack_2bit(flat[a:b])
    body = packed[:(l + 3) // 4]
    with open(prefix + ".pac", "wb") as f:
        f.write(body.tobytes())
        if l % 4 == 0:
            f.write(b"\0")
        f.write(bytes([l % 4]))
    with open(prefix + ".ann", "w") as f:
        f.write(f"{l} {len(g.seqs)} 11\n")
        off = 0
        for name, s in zip(g.names, g.seqs):
            f.write(f"0 {name} (null)\n{off} {len(s)} 0\n")
            off += len(s)
    with open(prefix + ".amb", "w") as f:
        f.write(f"{l} {len(g.seqs)} 0\n")
    packed[(l + 3) // 4] = 0
    return packed


# ----------------------------------------------------------------------------------------------
# reads
# ----------------------------------------------------------------------------------------------
def _mutate(reads: np.ndarray, p_sub: float, rng) -> np.ndarray:
    if p_sub > 0:
        m = rng.random(reads.shape) < p_sub
        k = int(m.sum())
        reads[m] = (reads[m] + rng.integers(1, 4, size=k, dtype=np.uint8)) & 3
    return reads


def _apply_indels(frags: np.ndarray, L: int, p_ins: float, p_del: float, rng) -> np.ndarray:
    """frags: (n, L+pad) source windows; returns (n, L) reads with 1-3 bp insertions/deletions."""
    n = frags.shape[0]
    out = frags[:, :L].copy()
    ev = rng.random((n, L)) < (p_ins + p_del)
    rows = np.nonzero(ev.any(axis=1))[0]
    for r in rows:
        src = frags[r]
        res = []
        i = 0
        cols = set(np.nonzero(ev[r])[0].tolist())
        while len(res) < L and i < len(src):
            if len(res) in cols:
                cols.discard(len(res))
                k = int(rng.integers(1, 4))
                if rng.random() < p_ins / (p_ins + p_del):
                    res.extend(rng.integers(0, 4, size=k).tolist())
                    continue
                i += k
                continue
            res.append(int(src[i]))
            i += 1
        res = (res + [0] * L)[:L]
        out[r] = np.asarray(res, dtype=np.uint8)
    return out


def _concat_source(g: Genome, spliced: bool):
    """Return (flat source codes, unit offsets, unit lengths). Units are contigs or transcripts."""
    if not spliced:
        units = g.seqs
    else:
        units = []
        for gene in g.genes:
            s = g.seqs[gene.contig]
            units.append(np.concatenate([s[a:b] for a, b in gene.exons]))
    lens = np.array([len(u) for u in units], dtype=np.int64)
    offs = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return np.concatenate(units), offs, lens


def _sample_fragments(flat, offs, lens, n, frag_len, rng):
    """Pick n windows of per-read length frag_len[i] fully inside one unit. Returns flat start indices."""
    ok = lens >= frag_len.max() + 8
    if not ok.any():
        raise ValueError("no source unit is long enough for the requested fragments")
    w = np.where(ok, lens, 0).astype(np.float64)
    unit = rng.choice(len(lens), size=n, p=w / w.sum())
    room = lens[unit] - frag_len - 4
    start = offs[unit] + (rng.random(n) * room).astype(np.int64)
    return start


def simulate_single(g: Genome, n: int, L: int = 100, p_sub: float = 0.01, seed: int = 2001,
                    spliced: bool = False, p_ins: float = 0.0, p_del: float = 0.0) -> np.ndarray:
    """(n, L) uint8 codes. Uniform start, strand 50/50."""
    rng = np.random.default_rng(seed)
    flat, offs, lens = _concat_source(g, spliced)
    pad = 16 if (p_ins or p_del) else 0
    fl = np.full(n, L + pad, dtype=np.int64)
    st = _sample_fragments(flat, offs, lens, n, fl, rng)
    fr = flat[st[:, None] + np.arange(L + pad)[None, :]]
    reads = _apply_indels(fr, L, p_ins, p_del, rng) if pad else fr
    rc = rng.random(n) < 0.5
    reads[rc] = revcomp_codes(reads[rc])
    return _mutate(reads, p_sub, rng)


def simulate_pairs(g: Genome, n: int, L: int = 101, p_sub: float = 0.01, seed: int = 2002,
                   frag_mean: float = 300.0, frag_sd: float = 30.0, frag_min: int | None = None,
                   frag_max: int = 500, spliced: bool = False, p_ins: float = 0.0, p_del: float = 0.0):
    """FR pairs: mate1 = fragment[:L], mate2 = revcomp(fragment)[:L]; fragment strand 50/50."""
    rng = np.random.default_rng(seed)
    flat, offs, lens = _concat_source(g, spliced)
    frag_min = frag_min or 2 * L
    frag_max = max(frag_max, frag_min + 1)
    pad = 16 if (p_ins or p_del) else 0
    fl = np.clip(rng.normal(frag_mean, frag_sd, n).round().astype(np.int64), frag_min, frag_max)
    st = _sample_fragments(flat, offs, lens, n, fl + pad, rng)
    ar = np.arange(L + pad)[None, :]
    left = flat[st[:, None] + ar]                                # fragment[:L+pad]
    right_fwd = flat[(st + fl - L - pad)[:, None] + ar]          # fragment[-(L+pad):]
    right = revcomp_codes(right_fwd)                             # reads inward from the fragment end
    if pad:
        m1 = _apply_indels(left, L, p_ins, p_del, rng)
        m2 = _apply_indels(right, L, p_ins, p_del, rng)
    else:
        m1, m2 = left.copy(), right.copy()
    swap = rng.random(n) < 0.5
    m1[swap], m2[swap] = m2[swap].copy(), m1[swap].copy()
    return _mutate(m1, p_sub, rng), _mutate(m2, p_sub, rng)


def fastq_bytes(reads: np.ndarray, mate: int | None = None, first_id: int = 0) -> bytes:
    """Fixed-width records: @r%08d[/m] \n seq \n + \n IIII.. \n — vectorised."""
    n, L = reads.shape
    ids = np.arange(first_id, first_id + n)
    digits = np.empty((n, 8), dtype=np.uint8)
    x = ids.copy()
    for k in range(7, -1, -1):
        digits[:, k] = 48 + (x % 10)
        x //= 10
    suffix = b"" if mate is None else b"/" + str(mate).encode()
    head = 2 + 8 + len(suffix) + 1
    rec = np.empty((n, head + L + 3 + L + 1), dtype=np.uint8)
    rec[:, 0] = ord("@")
    rec[:, 1] = ord("r")
    rec[:, 2:10] = digits
    for i, ch in enumerate(suffix):
        rec[:, 10 + i] = ch
    rec[:, head - 1] = 10
    rec[:, head:head + L] = ALPHABET[reads]
    rec[:, head + L:head + L + 3] = np.frombuffer(b"\n+\n", dtype=np.uint8)
    rec[:, head + L + 3:head + 2 * L + 3] = ord("I")
    rec[:, -1] = 10
    return rec.tobytes()


def write_fastq(path: str, reads: np.ndarray, mate: int | None = None) -> None:
    with open(path, "wb") as f:
        f.write(fastq_bytes(reads, mate))


# ----------------------------------------------------------------------------------------------
# named workloads (scaled by `scale` for tests; scale=1.0 is the BASELINE.json size)
# ----------------------------------------------------------------------------------------------
def config_genome(cfg: int, scale: float = 1.0) -> Genome:
    if cfg in (1, 2):
        return random_genome(int(4_600_000 * scale), 1, seed=1001)
    if cfg in (3, 4):
        g = random_genome(int(3_100_000_000 * scale), 24, seed=1003)
        add_gene_models(g, seed=1003)
        return g
    if cfg == 5:
        g = random_genome(int(4_600_000 * scale), 4, seed=1005)
        add_segmental_duplications(g, 0.30, seed=1005)
        return g
    raise ValueError(cfg)


def config_reads(cfg: int, g: Genome, n: int):
    """Returns (mate1, mate2 or None, extra dart flags)."""
    if cfg == 1:
        return simulate_single(g, n, 100, 0.01, seed=2001), None, []
    if cfg == 2:
        m1, m2 = simulate_pairs(g, n, 101, 0.01, seed=2002)
        return m1, m2, []
    if cfg == 3:
        m1, m2 = simulate_pairs(g, n, 101, 0.01, seed=2003, spliced=True, frag_min=202, frag_max=500)
        return m1, m2, []
    if cfg == 4:
        m1, m2 = simulate_pairs(g, n, 250, 0.03, seed=2004, frag_mean=600, frag_sd=50, frag_min=500,
                                frag_max=900, p_ins=0.002, p_del=0.002)
        return m1, m2, ["-mis", "10"]
    if cfg == 5:
        m1, m2 = simulate_pairs(g, n, 101, 0.01, seed=2005)
        return m1, m2, ["-m", "-max_dup", "10000", "-all_sj"]
    raise ValueError(cfg)


def materialise(cfg: int, outdir: str, n: int, scale: float = 1.0) -> dict:
    """Write genome FASTA + FASTQ for a config; returns paths and flags."""
    os.makedirs(outdir, exist_ok=True)
    g = config_genome(cfg, scale)
    fa = os.path.join(outdir, "genome.fa")
    write_fasta(fa, g)
    m1, m2, flags = config_reads(cfg, g, n)
    r1 = os.path.join(outdir, "r1.fq")
    write_fastq(r1, m1, 1 if m2 is not None else None)
    r2 = None
    if m2 is not None:
        r2 = os.path.join(outdir, "r2.fq")
        write_fastq(r2, m2, 2)
    return {"fasta": fa, "r1": r1, "r2": r2, "flags": flags, "genome": g}
