// Read ingest and SAM text on the device.
//
// Ingest — replaces GetNextEntry / GetNextChunk (/root/reference/src/GetData.cpp:77-179) for FASTQ input: the host only
// preads raw file blocks into page-locked memory and counts newlines to cut the blocks at record boundaries; the GPU finds
// the line starts (count / prefix sum / write), cuts the header the way IdentifyHeaderBegPos / IdentifyHeaderEndPos do
// (GetData.cpp:55-75), and encodes the bases — mate 2 reverse-complemented on the fly with GetComplementaryBase's table
// (/root/reference/src/tools.cpp:3-17: anything but ACGT/acgt becomes 'N'), as GetNextChunk leaves it (GetData.cpp:157-168).
//
// SAM text — replaces OutputPairedAlignments / OutputSingledAlignments (/root/reference/src/Mapping.cpp:208-369): complete
// lines (name, FLAG, RNAME, POS, MAPQ, CIGAR, RNEXT/PNEXT/TLEN, SEQ, QUAL, NM/AS/XS [XS:A]) are written into one byte pool
// in input order; the host only fwrite()s the pool.  Two passes: one thread per read (pair) measures its lines, a prefix sum
// over the reads places them, one WARP per read (pair) writes them — lane 0 the numeric fields, all lanes SEQ and QUAL.
// SEQ/QUAL orientation (Mapping.cpp:243-249, :313-319): the reference keeps mate 2 flipped in memory, prints the stored
// string for (mate 1, forward) and (mate 2, reverse) and GetComplementarySeq() of it otherwise; complementing twice turns
// lower-case and ambiguous symbols into upper-case ACGT / N, which is reproduced here ("normalised" mode).
#include "context.h"

namespace dartgpu {

namespace {

constexpr unsigned FULLS = 0xffffffffu;

struct SamDev {            // everything the SAM kernels read, on the device
    int n_reads, paired, multi_hit, unique;
    const dartgpu_read_result *rr; const dartgpu_report *rep; const char *cigars;
    const uint8_t *text;
    const int64_t *seq_pos, *name_pos, *qual_pos; const int32_t *rlen, *name_len, *qual_len;
    const char *chr_names; const int32_t *chr_name_off;
};
#define SAM_UNMAPPED_MID "*\t0\t0\t*\t*\t0\t0\t"
#define SAM_UNMAPPED_TAIL "\tAS:i:0\tXS:i:0\n"

// ---------------------------------------------------------------------------------------------------
// ingest
// ---------------------------------------------------------------------------------------------------
constexpr int NL_CHUNK = 64;     // bytes per thread

__global__ void k_nl_count(const uint8_t *__restrict__ text, int64_t len, uint32_t *cnt, int64_t n_chunks)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= n_chunks; c += (int64_t)gridDim.x * blockDim.x) {
        if (c == n_chunks) { cnt[c] = 0; continue; }
        const int64_t p0 = c * NL_CHUNK;
        uint32_t k = 0;
        if (p0 + NL_CHUNK <= len) {
            const uint4 *w = reinterpret_cast<const uint4 *>(text + p0);
#pragma unroll
            for (int i = 0; i < NL_CHUNK / 16; i++) {
                const uint4 x = w[i];
                k += __popc(__vcmpeq4(x.x, 0x0A0A0A0Au)) + __popc(__vcmpeq4(x.y, 0x0A0A0A0Au)) + __popc(__vcmpeq4(x.z, 0x0A0A0A0Au)) +
                     __popc(__vcmpeq4(x.w, 0x0A0A0A0Au));
            }
            k >>= 3;
        } else for (int64_t p = p0; p < len; p++) k += text[p] == '\n';
        cnt[c] = k;
    }
}

// line_start[j] = first byte of line j (line_start[0] = 0 is written by the caller's thread 0), up to max_lines entries
__global__ void k_nl_write(const uint8_t *__restrict__ text, int64_t len, const int64_t *__restrict__ off, int64_t n_chunks,
                           int64_t *line_start, int64_t max_lines)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (int64_t)gridDim.x * blockDim.x) {
        if (c == 0) line_start[0] = 0;
        int64_t j = off[c] + 1;
        if (off[c + 1] == off[c]) continue;
        const int64_t p0 = c * NL_CHUNK, p1 = min(p0 + NL_CHUNK, len);
        for (int64_t p = p0; p < p1; p++)
            if (text[p] == '\n') { if (j <= max_lines) line_start[j] = p + 1; j++; }
    }
}

struct FqLayout {
    const uint8_t *text; int64_t base2;                  // file 2 starts at text + base2
    const int64_t *ls1, *ls2;                            // line starts of the two files (ls2 relative to base2)
    int n_reads; int two_files; int flip_odd;
    int64_t *seq_pos, *name_pos, *qual_pos; int32_t *rlen, *name_len, *qual_len; uint32_t *padded;
    int32_t *err, *max_rlen; unsigned long long *bases; const int64_t *nl_total1, *nl_total2; int64_t want1, want2;
};

__global__ void k_fq_records(FqLayout F)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (*F.nl_total1 != F.want1 || (F.two_files && *F.nl_total2 != F.want2)) atomicOr(F.err, ERR_FASTQ_LINES);
    }
    int mx = 0;
    unsigned long long nb = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= F.n_reads; i += gridDim.x * blockDim.x) {
        if (i == F.n_reads) { F.padded[i] = 0; continue; }
        const bool second = F.two_files && (i & 1);
        const int rec = F.two_files ? i >> 1 : i;
        const int64_t *ls = second ? F.ls2 : F.ls1;
        const int64_t fb = second ? F.base2 : 0;
        const uint8_t *t = F.text + fb;
        const int64_t h0 = ls[4 * rec], s0 = ls[4 * rec + 1], q0 = ls[4 * rec + 3], e0 = ls[4 * rec + 4];
        // header: getline length includes the newline (GetData.cpp:86-90)
        const int hl = (int)(s0 - h0);
        int p1 = hl - 1, p2 = hl - 1;
        for (int k = 1; k < hl; k++) if (t[h0 + k] != '>' && t[h0 + k] != '@') { p1 = k; break; }
        for (int k = 1; k < hl; k++) { const uint8_t ch = t[h0 + k]; if (ch == ' ' || ch == '/' || ch == '\t') { p2 = k; break; } }
        F.name_pos[i] = fb + h0 + p1; F.name_len[i] = p2 > p1 ? p2 - p1 : 0;
        // sequence: rlen = getline length - 1 (GetData.cpp:94-101)
        int rl = (int)(ls[4 * rec + 2] - s0) - 1;
        if (rl < 0) rl = 0;
        if (rl > DARTGPU_MAX_RLEN) { atomicOr(F.err, ERR_READ_TOO_LONG); rl = DARTGPU_MAX_RLEN; }
        F.seq_pos[i] = fb + s0; F.rlen[i] = rl;
        F.qual_pos[i] = fb + q0; F.qual_len[i] = min(rl, (int)(e0 - q0));      // strncpy(qual, line, rlen): the line may be shorter
        F.padded[i] = (uint32_t)((rl + 15) & ~15);
        mx = max(mx, rl); nb += (unsigned long long)rl;
    }
    mx = __reduce_max_sync(FULLS, mx);
    for (int d = 16; d > 0; d >>= 1) nb += __shfl_xor_sync(FULLS, nb, d);
    if ((threadIdx.x & 31) == 0) { if (mx) atomicMax(F.max_rlen, mx); if (nb) atomicAdd(F.bases, nb); }
}

// GetComplementaryBase (tools.cpp:3-17) on raw characters
__device__ __forceinline__ uint8_t comp_char(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    case 'T': case 't': return 'A';
    default: return 'N';
    }
}
__device__ __forceinline__ uint8_t code_of_char(uint8_t c)
{
    switch (c) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
    case 'a': return 8; case 'c': return 9; case 'g': return 10; case 't': return 11;
    case 'N': return CODE_N;
    default: return CODE_OTHER;
    }
}

// codes + the search kernel's packed view straight from the file text; 8 lanes per read, 16 bases per lane and iteration
__global__ void k_encode_fastq(const uint8_t *__restrict__ text, const int64_t *__restrict__ seq_pos, const int32_t *__restrict__ rlen,
                               const int64_t *__restrict__ dev_off, int n, int flip_odd, uint8_t *codes, uint2 *packed)
{
    const int gl = threadIdx.x & 7;
    const int ngroups = (gridDim.x * blockDim.x) >> 3;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n; r += ngroups) {
        const uint8_t *src = text + seq_pos[r];
        const int rl = rlen[r];
        const bool flip = flip_odd && (r & 1);
        const int chunks = (rl + 15) >> 4;
        const int64_t d0 = dev_off[r];
        uint4 *dst = reinterpret_cast<uint4 *>(codes + d0);
        for (int ch = gl; ch < chunks; ch += 8) {
            uint32_t w[4] = {0, 0, 0, 0}, two = 0, amb = 0;
#pragma unroll
            for (int k = 0; k < 16; k++) {
                const int p = ch * 16 + k;
                uint32_t c = 4u;
                if (p < rl) c = flip ? code_of_char(comp_char(src[rl - 1 - p])) : code_of_char(src[p]);
                w[k >> 2] |= c << (8 * (k & 3));
                two |= (c & 3u) << (2 * k);
                amb |= ((c >> 2) & 1u) << k;
            }
            dst[ch] = make_uint4(w[0], w[1], w[2], w[3]);
            packed[(d0 >> 4) + ch] = make_uint2(two, amb);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// SAM text
// ---------------------------------------------------------------------------------------------------
struct Line {              // one output line
    int read;              // read index in the batch: name, SEQ, QUAL
    int mapped;
    int flag, chr, mapq, cigar_off, cigar_len, has_mate, dist, seq_mode, nm, as, xs, xs_a;
    long long pos, mpos;
};
enum { SEQ_AS_IS = 0, SEQ_FLIPPED = 1, SEQ_NORMALISED = 2 };

__device__ __forceinline__ int dec_len(long long v) { int d = v < 0 ? 2 : 1; if (v < 0) v = -v; while (v >= 10) { v /= 10; d++; } return d; }
__device__ __forceinline__ char *put_dec(char *o, long long v)
{
    if (v < 0) { *o++ = '-'; v = -v; }
    char tmp[20]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *o++ = tmp[--n];
    return o;
}
__device__ __forceinline__ char *put_str(char *o, const char *s) { while (*s) *o++ = *s++; return o; }
__device__ __forceinline__ int str_len(const char *s) { int n = 0; while (s[n]) n++; return n; }

// bytes of the line before SEQ (incl. the tab in front of SEQ) and after QUAL (incl. the newline)
__device__ int line_head_len(const SamDev &S, const Line &L)
{
    int n = S.name_len[L.read] + 1 + dec_len(L.flag) + 1;
    if (!L.mapped) return n + (int)sizeof(SAM_UNMAPPED_MID) - 1;
    n += S.chr_name_off[L.chr + 1] - S.chr_name_off[L.chr] + 1 + dec_len(L.pos) + 1 + dec_len(L.mapq) + 1 + L.cigar_len + 1;
    if (L.has_mate) n += 2 + dec_len(L.mpos) + 1 + dec_len(L.dist) + 1;   // "=\t" pos "\t" dist "\t"
    else n += 6;                                                     // "*\t0\t0\t"
    return n;
}
__device__ int line_tail_len(const Line &L)
{
    if (!L.mapped) return (int)sizeof(SAM_UNMAPPED_TAIL) - 1;
    return 6 + dec_len(L.nm) + 6 + dec_len(L.as) + 6 + dec_len(L.xs) + (L.xs_a ? 7 : 0) + 1;
}
__device__ char *line_head_put(const SamDev &S, const Line &L, char *o)
{
    const uint8_t *nm = S.text + S.name_pos[L.read];
    for (int k = 0; k < S.name_len[L.read]; k++) *o++ = (char)nm[k];
    *o++ = '\t'; o = put_dec(o, L.flag); *o++ = '\t';
    if (!L.mapped) return put_str(o, SAM_UNMAPPED_MID);
    const char *cn = S.chr_names + S.chr_name_off[L.chr];
    for (int k = 0, e = S.chr_name_off[L.chr + 1] - S.chr_name_off[L.chr]; k < e; k++) *o++ = cn[k];
    *o++ = '\t'; o = put_dec(o, L.pos); *o++ = '\t'; o = put_dec(o, L.mapq); *o++ = '\t';
    for (int k = 0; k < L.cigar_len; k++) *o++ = S.cigars[L.cigar_off + k];
    *o++ = '\t';
    if (L.has_mate) { *o++ = '='; *o++ = '\t'; o = put_dec(o, L.mpos); *o++ = '\t'; o = put_dec(o, L.dist); *o++ = '\t'; }
    else o = put_str(o, "*\t0\t0\t");
    return o;
}
__device__ char *line_tail_put(const Line &L, char *o)
{
    if (!L.mapped) return put_str(o, SAM_UNMAPPED_TAIL);
    o = put_str(o, "\tNM:i:"); o = put_dec(o, L.nm); o = put_str(o, "\tAS:i:"); o = put_dec(o, L.as);
    o = put_str(o, "\tXS:i:"); o = put_dec(o, L.xs);
    if (L.xs_a) o = put_str(o, L.xs_a == 1 ? " XS:A:+" : " XS:A:-");
    *o++ = '\n';
    return o;
}

// The lines of one read (single-end) or pair, in the reference's order.  `sink(Line)` is called once per line by every
// thread that runs this (one thread in the measuring pass, the 32 lanes of a warp in the writing pass).
template <class Sink>
__device__ void for_each_line(const SamDev &S, int u, Sink &sink, int &n_unmapped, int &n_unique, int &n_paired)
{
    auto xs_of = [](int sj, bool second) { return sj == -1 ? 0 : ((sj == 0 || sj == 2) != second) ? 1 : 2; };
    if (!S.paired) {
        const dartgpu_read_result rr = S.rr[u];
        Line L{}; L.read = u;
        if (rr.score == 0) { n_unmapped++; L.mapped = 0; L.flag = S.rep[rr.report_off].flag; L.seq_mode = SEQ_AS_IS; sink(L); return; }
        if (S.unique && rr.mapq <= 3) return;
        if (rr.mapq == 50) n_unique++;
        for (int i = rr.best; i < rr.n_reports; i++) {
            const dartgpu_report p = S.rep[rr.report_off + i];
            if (p.aln_score == rr.score) {
                L.mapped = 1; L.flag = p.flag; L.chr = p.chr_idx; L.pos = p.pos; L.mapq = rr.mapq; L.cigar_off = p.cigar_off; L.cigar_len = p.cigar_len;
                L.has_mate = 0; L.seq_mode = p.dir ? SEQ_AS_IS : SEQ_FLIPPED; L.nm = rr.mis_num; L.as = rr.score; L.xs = rr.sub_score;
                L.xs_a = xs_of(p.sj_type, false);
                sink(L);
                if (!S.multi_hit) break;
            }
        }
        return;
    }
    const int g1 = 2 * u, g2 = g1 + 1;
    const dartgpu_read_result r1 = S.rr[g1], r2 = S.rr[g2];
    const int L1 = S.rlen[g1], L2 = S.rlen[g2];
    {
        Line L{}; L.read = g1;
        if (r1.score == 0) { n_unmapped++; L.mapped = 0; L.flag = S.rep[r1.report_off].flag; L.seq_mode = SEQ_AS_IS; sink(L); }
        else if (!S.unique || r1.mapq > 3) {
            if (r1.mapq == 50) n_unique++;
            for (int i = r1.best; i < r1.n_reports; i++) {
                const dartgpu_report p = S.rep[r1.report_off + i];
                if (p.aln_score > 0) {
                    const int j = p.paired_idx;
                    bool mate = false; dartgpu_report m{};
                    if (j != -1) { m = S.rep[r2.report_off + j]; mate = m.aln_score > 0; }
                    L.mapped = 1; L.flag = p.flag; L.chr = p.chr_idx; L.pos = p.pos; L.mapq = r1.mapq; L.cigar_off = p.cigar_off; L.cigar_len = p.cigar_len;
                    L.has_mate = mate; L.mpos = mate ? m.pos : 0;
                    L.dist = mate ? (int)(m.pos - p.pos + (p.dir ? L2 : 0 - L1)) : 0;
                    if (mate && i == r1.best) n_paired += 2;
                    L.seq_mode = p.dir ? SEQ_AS_IS : SEQ_FLIPPED; L.nm = r1.mis_num; L.as = r1.score; L.xs = r1.sub_score; L.xs_a = xs_of(p.sj_type, false);
                    sink(L);
                }
                if (!S.multi_hit) break;
            }
        }
    }
    {
        Line L{}; L.read = g2;
        // mate 2 is stored flipped: the stored string (what a reverse-strand report and an unmapped read print) is the
        // reverse complement of the file's; a forward-strand report prints its complement back: the file's, normalised
        if (r2.score == 0) { n_unmapped++; L.mapped = 0; L.flag = S.rep[r2.report_off].flag; L.seq_mode = SEQ_FLIPPED; sink(L); }
        else if (!S.unique || r2.mapq > 3) {
            if (r2.mapq == 50) n_unique++;
            for (int j = r2.best; j < r2.n_reports; j++) {
                const dartgpu_report p = S.rep[r2.report_off + j];
                if (p.aln_score > 0) {
                    const int i = p.paired_idx;
                    bool mate = false; dartgpu_report m{};
                    if (i != -1) { m = S.rep[r1.report_off + i]; mate = m.aln_score > 0; }
                    L.mapped = 1; L.flag = p.flag; L.chr = p.chr_idx; L.pos = p.pos; L.mapq = r2.mapq; L.cigar_off = p.cigar_off; L.cigar_len = p.cigar_len;
                    L.has_mate = mate; L.mpos = mate ? m.pos : 0;
                    L.dist = mate ? 0 - (int)(p.pos - m.pos + (m.dir ? L2 : 0 - L1)) : 0;
                    L.seq_mode = p.dir ? SEQ_NORMALISED : SEQ_FLIPPED; L.nm = r2.mis_num; L.as = r2.score; L.xs = r2.sub_score; L.xs_a = xs_of(p.sj_type, true);
                    sink(L);
                }
                if (!S.multi_hit) break;
            }
        }
    }
}

struct MeasureSink {
    const SamDev &S; unsigned long long bytes = 0;
    __device__ explicit MeasureSink(const SamDev &s) : S(s) {}
    __device__ void operator()(const Line &L) { bytes += (unsigned long long)(line_head_len(S, L) + 2 * S.rlen[L.read] + 1 + line_tail_len(L)); }
};

__global__ void k_sam_measure(SamDev S, int n_units, uint32_t *unit_bytes, BatchCtl *ctl)
{
    if (ctl->abort) return;
    int un = 0, uq = 0, pr = 0;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u <= n_units; u += gridDim.x * blockDim.x) {
        if (u == n_units) { unit_bytes[u] = 0; continue; }
        MeasureSink sink(S);
        for_each_line(S, u, sink, un, uq, pr);
        unit_bytes[u] = (uint32_t)sink.bytes;
    }
    un = __reduce_add_sync(FULLS, un); uq = __reduce_add_sync(FULLS, uq); pr = __reduce_add_sync(FULLS, pr);
    if ((threadIdx.x & 31) == 0) {
        if (un) atomicAdd(&ctl->sam_counts[0], (unsigned long long)un);
        if (uq) atomicAdd(&ctl->sam_counts[1], (unsigned long long)uq);
        if (pr) atomicAdd(&ctl->sam_counts[2], (unsigned long long)pr);
    }
}

__global__ void k_ctl_sam(BatchCtl *ctl, const int64_t *unit_off, int n_units, long long cap)
{
    if (ctl->abort) return;
    const long long t = unit_off[n_units];
    ctl->sam_bytes = t;
    if (t > cap) atomicOr(&ctl->abort, CAP_SAM);
}

struct WriteSink {
    const SamDev &S; char *out; int lane;
    __device__ WriteSink(const SamDev &s, char *o, int l) : S(s), out(o), lane(l) {}
    __device__ void operator()(const Line &L)
    {
        const int hl = line_head_len(S, L), rl = S.rlen[L.read], ql = S.qual_len[L.read], tl = line_tail_len(L);
        if (lane == 0) { line_head_put(S, L, out); out[hl + rl] = '\t'; line_tail_put(L, out + hl + 2 * rl + 1); }
        const uint8_t *sq = S.text + S.seq_pos[L.read], *qq = S.text + S.qual_pos[L.read];
        char *so = out + hl, *qo = out + hl + rl + 1;
        for (int k = lane; k < rl; k += 32) {
            uint8_t b, q;
            if (L.seq_mode == SEQ_AS_IS) { b = sq[k]; q = k < ql ? qq[k] : 0; }
            else if (L.seq_mode == SEQ_FLIPPED) { b = comp_char(sq[rl - 1 - k]); q = rl - 1 - k < ql ? qq[rl - 1 - k] : 0; }
            else { b = comp_char(comp_char(sq[k])); q = k < ql ? qq[k] : 0; }
            so[k] = (char)b; qo[k] = (char)q;
        }
        out += hl + 2 * rl + 1 + tl;
    }
};

__global__ void __launch_bounds__(128) k_sam_write(SamDev S, int n_units, const int64_t *__restrict__ unit_off, char *pool, const BatchCtl *ctl)
{
    if (ctl->abort) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; u < n_units; u += nwarps) {
        if (unit_off[u + 1] == unit_off[u]) continue;
        WriteSink sink(S, pool + unit_off[u], lane);
        int a = 0, b = 0, c = 0;
        for_each_line(S, u, sink, a, b, c);
    }
}

} // namespace

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static inline int grid_cap(int64_t want, int per_sm) { const int64_t cap = (int64_t)sm_count() * per_sm; return (int)std::max<int64_t>(1, std::min(want, cap)); }

// line starts of one file block already on the device
static void enqueue_line_starts(dartgpu_ctx *c, const uint8_t *text, int64_t len, int64_t n_lines, DevBuf<int64_t> &ls, DevBuf<uint32_t> &cnt,
                                DevBuf<int64_t> &off)
{
    cudaStream_t st = c->stream;
    const int64_t n_chunks = (len + NL_CHUNK - 1) / NL_CHUNK;
    cnt.reserve(n_chunks + 2); off.reserve(n_chunks + 2); ls.reserve(n_lines + 2);
    if (n_chunks >= INT32_MAX) throw std::make_pair(DARTGPU_ERR_ARG, std::string("FASTQ block too large"));
    k_nl_count<<<grid_cap((n_chunks + 1 + 255) / 256, 16), 256, 0, st>>>(text, len, cnt.p, n_chunks);
    const size_t tmp = scan_tmp_bytes((int)n_chunks);
    c->d_scan_tmp.reserve(tmp + 256);
    launch_scan_u32_to_i64(cnt.p, off.p, (int)n_chunks, c->d_scan_tmp.p, tmp, st);
    k_nl_write<<<grid_cap((n_chunks + 255) / 256, 16), 256, 0, st>>>(text, len, off.p, n_chunks, ls.p, n_lines);
    c->stats.kernel_launches += 4;
}

// dartgpu_submit_fastq: the read batch comes as raw FASTQ text (one or two file blocks); everything per read happens here
void upload_fastq(dartgpu_ctx *c, const dartgpu_fastq_block *b)
{
    if (!b->fastq) throw std::make_pair(DARTGPU_ERR_ARG, std::string("GPU ingest handles FASTQ records only"));
    if (b->n_records < 0 || b->len1 < 0 || (b->n_records > 0 && !b->text1) || (b->text2 && b->len2 < 0))
        throw std::make_pair(DARTGPU_ERR_ARG, std::string("bad FASTQ block"));
    const bool two = b->text2 != nullptr;
    const int64_t n64 = (int64_t)b->n_records * (two ? 2 : 1);
    if (n64 > INT32_MAX / 8) throw std::make_pair(DARTGPU_ERR_ARG, std::string("FASTQ block holds too many records"));
    const int n = (int)n64;
    const bool paired = c->prm.pair_end != 0;
    if (paired && (n & 1)) throw std::make_pair(DARTGPU_ERR_ARG, std::string("paired-end batch with an odd number of reads"));
    FastqDev &F = c->fq;
    cudaStream_t st = c->stream;
    c->n_reads = n; c->from_fastq = true;
    c->max_rlen = 0;
    if (n == 0) return;
    const int64_t base2 = (b->len1 + 64 + 63) & ~(int64_t)63;
    const int64_t total = base2 + (two ? b->len2 : 0) + 64;
    F.text.reserve(total);
    // The longest read sizes scratch (search records, NW rows, 8-mer tables) but is only known once the GPU has parsed the
    // text: assume the caller's hint, else the first record's length, never less than any read this context has seen; the
    // device compares (k_ctl_ingest) and a longer read costs one more attempt with the right bound (CAP_RLEN).
    int hint = b->max_read_len;
    if (hint <= 0) {
        const char *nl = (const char *)memchr(b->text1, '\n', (size_t)b->len1);
        const char *nl2 = nl ? (const char *)memchr(nl + 1, '\n', (size_t)(b->text1 + b->len1 - nl - 1)) : nullptr;
        hint = nl2 ? (int)(nl2 - nl - 1) : 0;
    }
    F.rlen_seen = std::max(F.rlen_seen, std::min(hint, DARTGPU_MAX_RLEN));
    c->max_rlen = F.rlen_seen;
    c->cap_rec = std::max(1, (c->max_rlen + 15) / 16);
    DG_CUDA(cudaEventRecord(c->ev[0], st));
    auto pinned = [](const void *p) {
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
        return at.type == cudaMemoryTypeHost;
    };
    auto up = [&](const char *src, int64_t len, int64_t at) {
        if (len == 0) return;
        if (pinned(src)) { DG_CUDA(cudaMemcpyAsync(F.text.p + at, src, (size_t)len, cudaMemcpyHostToDevice, st)); return; }
        c->h_raw.reserve(total);
        memcpy(c->h_raw.p + at, src, (size_t)len);
        DG_CUDA(cudaMemcpyAsync(F.text.p + at, c->h_raw.p + at, (size_t)len, cudaMemcpyHostToDevice, st));
    };
    up(b->text1, b->len1, 0);
    if (two) up(b->text2, b->len2, base2);
    c->stats.h2d_bytes += b->len1 + (two ? b->len2 : 0);
    const int64_t lines1 = 4ll * b->n_records, lines2 = two ? lines1 : 0;
    enqueue_line_starts(c, F.text.p, b->len1, lines1, F.ls1, F.cnt1, F.off1);
    if (two) enqueue_line_starts(c, F.text.p + base2, b->len2, lines2, F.ls2, F.cnt2, F.off2);
    F.seq_pos.reserve(n + 1); F.name_pos.reserve(n + 1); F.qual_pos.reserve(n + 1); F.name_len.reserve(n + 1); F.qual_len.reserve(n + 1);
    c->d_rlen.reserve(n + 1); c->d_padded.reserve(n + 2); c->d_dev_off.reserve(n + 2);
    const int64_t n_bases_bound = b->len1 + (two ? b->len2 : 0);
    const int64_t code_bytes = n_bases_bound / 2 + 15ll * n + 64;      // sequences are less than half of a FASTQ file
    c->d_codes.reserve(code_bytes); c->d_packed.reserve(code_bytes / 16 + 4);
    c->n_code_bytes = code_bytes;
    FqLayout L{};
    L.text = F.text.p; L.base2 = base2; L.ls1 = F.ls1.p; L.ls2 = two ? F.ls2.p : F.ls1.p; L.n_reads = n; L.two_files = two;
    L.flip_odd = paired;
    L.seq_pos = F.seq_pos.p; L.name_pos = F.name_pos.p; L.qual_pos = F.qual_pos.p; L.rlen = c->d_rlen.p; L.name_len = F.name_len.p; L.qual_len = F.qual_len.p;
    L.padded = c->d_padded.p; L.err = &c->d_ctl.p->ingest_err; L.max_rlen = &c->d_ctl.p->ingest_max_rlen; L.bases = &c->d_ctl.p->ingest_bases;
    const int64_t nc1 = (b->len1 + NL_CHUNK - 1) / NL_CHUNK, nc2 = two ? (b->len2 + NL_CHUNK - 1) / NL_CHUNK : 0;
    L.nl_total1 = F.off1.p + nc1; L.nl_total2 = two ? F.off2.p + nc2 : F.off1.p + nc1; L.want1 = lines1; L.want2 = lines2;
    launch_zero(&c->d_ctl.p->ingest_err, sizeof(BatchCtl) - BATCHCTL_RESET_BYTES, st);   // this part survives the per-attempt reset of the control block
    k_fq_records<<<grid_cap((n + 1 + 255) / 256, 8), 256, 0, st>>>(L);
    const size_t tmp = scan_tmp_bytes(n);
    c->d_scan_tmp.reserve(tmp + 256);
    launch_scan_u32_to_i64(c->d_padded.p, c->d_dev_off.p, n, c->d_scan_tmp.p, tmp, st);
    k_encode_fastq<<<grid_cap(((int64_t)n * 8 + 255) / 256, 16), 256, 0, st>>>(F.text.p, F.seq_pos.p, c->d_rlen.p, c->d_dev_off.p, n, paired ? 1 : 0,
                                                                             c->d_codes.p, c->d_packed.p);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[1], st));
    c->stats.kernel_launches += 5;
    c->stats.read_bases = 0;            // known on the device only
}

// SAM lines of the batch into the context's byte pool + the copy of the predicted size (enqueue_pipeline calls this)
void enqueue_sam(dartgpu_ctx *c, const dartgpu_read_result *rr, const dartgpu_report *rep, const char *cigars)
{
    FastqDev &F = c->fq;
    cudaStream_t st = c->stream;
    const int n = c->n_reads, paired = c->prm.pair_end != 0, units = paired ? n / 2 : n;
    SharedIndex &X = *c->shared;
    SamDev S{};
    S.n_reads = n; S.paired = paired; S.multi_hit = c->prm.multi_hit; S.unique = c->prm.unique;
    S.rr = rr; S.rep = rep; S.cigars = cigars; S.text = F.text.p;
    S.seq_pos = F.seq_pos.p; S.name_pos = F.name_pos.p; S.qual_pos = F.qual_pos.p; S.rlen = c->d_rlen.p; S.name_len = F.name_len.p; S.qual_len = F.qual_len.p;
    S.chr_names = X.d_chr_names.p; S.chr_name_off = X.d_chr_name_off.p;
    F.unit_bytes.reserve(units + 2); F.unit_off.reserve(units + 2); F.sam.reserve(c->caps.sam + 1);
    k_sam_measure<<<grid_cap((units + 1 + 127) / 128, 16), 128, 0, st>>>(S, units, F.unit_bytes.p, c->d_ctl.p);
    const size_t tmp = scan_tmp_bytes(units);
    c->d_scan_tmp.reserve(tmp + 256);
    launch_scan_u32_to_i64(F.unit_bytes.p, F.unit_off.p, units, c->d_scan_tmp.p, tmp, st);
    k_ctl_sam<<<1, 1, 0, st>>>(c->d_ctl.p, F.unit_off.p, units, c->caps.sam);
    k_sam_write<<<grid_cap(((int64_t)units * 32 + 127) / 128, 16), 128, 0, st>>>(S, units, F.unit_off.p, F.sam.p, c->d_ctl.p);
    DG_CUDA(cudaGetLastError());
    c->stats.kernel_launches += 5;
}

} // namespace dartgpu
