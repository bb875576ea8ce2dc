// One-time GPU-side re-layout of the BWA-format index into the HBM-resident tables the kernels read.
// File formats: /root/reference/src/bwt_index.cpp:15-35 (.sa), :102-121 (.bwt), :193-212 (.pac decode);
// block interleave written by /root/reference/src/BWT_Index/bwtindex.c:53-75.
#include "dartgpu_internal.h"
#include "rank.cuh"

#include <atomic>

namespace dartgpu {

int sm_count()
{   // per device: a process drives several GPUs from several threads (round-1 advice: function-local statics keyed to
    // the first device were a data race and sized grids for the wrong GPU)
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int v = cache[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 1; }
        cache[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

__global__ void k_set_i32(int32_t *dst, int32_t v) { *dst = v; }
void launch_set_i32(int32_t *dst, int32_t v, cudaStream_t st) { k_set_i32<<<1, 1, 0, st>>>(dst, v); }

__global__ void k_zero(uint32_t *p, size_t n_words)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (size_t)gridDim.x * blockDim.x) p[i] = 0u;
}
void launch_zero(void *p, size_t bytes, cudaStream_t st)
{
    const size_t n = bytes / 4;
    if (!n) return;
    size_t want = (n + 255) / 256;
    const int cap = sm_count() * 8;
    k_zero<<<(int)(want < (size_t)cap ? want : (size_t)cap), 256, 0, st>>>(reinterpret_cast<uint32_t *>(p), n);
}

__global__ void k_ctl_check(BatchCtl *ctl, long long *dst, const int64_t *src, int64_t cap, int bit)
{
    const long long v = *src;
    *dst = v;
    if (v > cap) atomicOr(&ctl->abort, bit);
}
void launch_ctl_check(BatchCtl *ctl, long long *dst, const int64_t *src, int64_t cap, int bit, cudaStream_t st)
{
    k_ctl_check<<<1, 1, 0, st>>>(ctl, dst, src, cap, bit);
}

// BWA block (128 symbols) = 8 words of counts (4 x u64, little endian) + 8 words of symbols (16 per word, first symbol
// in the top bits).  Output: two Occ32 blocks (rank.cuh) per BWA block; the second one's counts include the first half.
__global__ void k_relayout_occ32(const uint32_t *__restrict__ w, uint64_t n_words, Occ32 *__restrict__ occ, uint64_t n_blocks32)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t <= n_blocks32; t += (uint64_t)gridDim.x * blockDim.x) {   // one guard block past the end
        const uint64_t B = t >> 1;
        const int h = (int)(t & 1);
        auto word = [&](uint64_t i) -> uint32_t { return i < n_words ? w[i] : 0u; };
        const uint64_t base = B * 16;
        uint64_t cnt[4];
        for (int c = 0; c < 4; c++) cnt[c] = (uint64_t)word(base + 2 * c) | (uint64_t)word(base + 2 * c + 1) << 32;
        uint64_t lo[2] = {0, 0}, hi[2] = {0, 0};
        for (int half = 0; half <= h; half++)
            for (int j = 0; j < 4; j++) {
                const uint32_t x = word(base + 8 + 4 * half + j);
                for (int i = 0; i < 16; i++) {
                    const uint32_t sym = (x >> (30 - 2 * i)) & 3u;
                    lo[half] |= (uint64_t)(sym & 1u) << (16 * j + i);
                    hi[half] |= (uint64_t)(sym >> 1) << (16 * j + i);
                }
            }
        if (h) {
            cnt[0] += __popcll(~hi[0] & ~lo[0]); cnt[1] += __popcll(~hi[0] & lo[0]);
            cnt[2] += __popcll(hi[0] & ~lo[0]);  cnt[3] += __popcll(hi[0] & lo[0]);
        }
        Occ32 o;
        for (int c = 0; c < 4; c++) o.cnt[c] = (uint32_t)cnt[c];
        o.lo = lo[h]; o.hi = hi[h];
        occ[t] = o;
    }
}

void launch_relayout_occ32(const uint32_t *bwt_words, uint64_t n_words, Occ32 *occ, uint64_t n_blocks32, cudaStream_t st)
{
    uint64_t want = (n_blocks32 + 1 + 255) / 256;
    int grid = (int)(want < sm_count() * 16 ? want : sm_count() * 16);
    if (grid < 1) grid = 1;
    k_relayout_occ32<<<grid, 256, 0, st>>>(bwt_words, n_words, occ, n_blocks32);
}

// Search-start table, level by level: entry i of level j+1 = one backward step (k_search's) from entry (i mod 4^j) of
// level j with read base b = i >> 2j, i.e. BWT symbol c = 3 - b.
__global__ void k_ktab_level(DevIndex ix, int j, const KmerStart *__restrict__ prev, KmerStart *__restrict__ next)
{
    const uint64_t n_next = 1ull << (2 * (j + 1)), pmask = (1ull << (2 * j)) - 1;
    const char *occ = reinterpret_cast<const char *>(ix.occ32);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_next; i += (uint64_t)gridDim.x * blockDim.x) {
        const int b = (int)(i >> (2 * j));
        KmerStart e;
        if (j == 0) { e.x1 = ix.L2[3 - b] + 1; e.x2 = (uint32_t)(ix.L2[b + 1] - ix.L2[b]); e.splits = 0; next[i] = e; continue; }
        e = prev[i & pmask];
        if (e.x2 != 0) {
            const int c = 3 - b;
            const uint64_t k = e.x1 - 1, l = k + e.x2;
            const uint64_t kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);
            const Occ32 Bk = *reinterpret_cast<const Occ32 *>(occ + (kk >> 6) * 32), Bl = *reinterpret_cast<const Occ32 *>(occ + (ll >> 6) * 32);
            const uint32_t ok = Bk.cnt[c] + occ32_eq_upto(Bk.lo, Bk.hi, c, (uint32_t)kk & 63u);
            const uint32_t ol = Bl.cnt[c] + occ32_eq_upto(Bl.lo, Bl.hi, c, (uint32_t)ll & 63u);
            e.splits += (uint32_t)(((kk ^ ll) >> 7) != 0);
            e.x2 = ol - ok;
            e.x1 = ix.L2[c] + 1 + ok;
        }
        next[i] = e;
    }
}

void launch_build_ktab(const DevIndex &ix, int K, KmerStart *out, KmerStart *tmp, cudaStream_t st)
{
    // ping-pong so that the last level lands in `out`
    KmerStart *bufs[2] = {(K & 1) ? out : tmp, (K & 1) ? tmp : out};
    for (int j = 0; j < K; j++) {
        const uint64_t n_next = 1ull << (2 * (j + 1));
        uint64_t want = (n_next + 255) / 256;
        int grid = (int)(want < sm_count() * 16 ? want : sm_count() * 16);
        k_ktab_level<<<grid, 256, 0, st>>>(ix, j, bufs[(j + 1) & 1], bufs[j & 1]);
    }
}

// One pass of the LF mapping over the whole text: every entry of the file's sampled SA starts a walker that writes
// SA[k] = v, steps k -> LF(k), v -> v-1, and stops at the next sampled index (which another walker owns).  Together the
// walkers visit every SA index exactly once; entries at multiples of 2^shift are kept.  bwt_invPsi:
// /root/reference/src/bwt_search.cpp:119-125.
template <typename SaT>
__global__ void k_sa_densify(DevIndex ix, const uint64_t *__restrict__ sa_file, uint64_t sa_intv, uint64_t n_sa_file, SaT *__restrict__ out)
{
    __shared__ uint64_t s_L2[4];
    if (threadIdx.x == 0) { s_L2[0] = ix.L2[0]; s_L2[1] = ix.L2[1]; s_L2[2] = ix.L2[2]; s_L2[3] = ix.L2[3]; }
    __syncthreads();
    const char *occ = reinterpret_cast<const char *>(ix.occ32);
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; j < n_sa_file; j += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t k = j * sa_intv;
        uint64_t v = j == 0 ? ix.seq_len : sa_file[j];     // SA[0] is the empty suffix at position seq_len (kept as -1 in the file)
        for (;;) {
            if ((k & ix.sa_mask) == 0) out[k >> ix.sa_shift] = (SaT)v;
            if (k == ix.primary) k = 0;
            else {
                const uint64_t kk = k - (k > ix.primary);
                const char *blk = occ + (kk >> 6) * 32;
                const ulonglong2 pl = __ldg(reinterpret_cast<const ulonglong2 *>(blk + 16));
                const uint32_t t = (uint32_t)kk & 63u;
                const int c = occ32_symbol(pl.x, pl.y, t);
                k = s_L2[c] + __ldg(reinterpret_cast<const uint32_t *>(blk) + c) + occ32_eq_upto(pl.x, pl.y, c, t);
            }
            v--;
            if ((k & (sa_intv - 1)) == 0) break;
        }
    }
}

void launch_sa_densify(const DevIndex &ix, const uint64_t *sa_file, uint64_t sa_intv, uint64_t n_sa_file, void *out, cudaStream_t st)
{
    uint64_t want = (n_sa_file + 255) / 256;
    int grid = (int)(want < sm_count() * 32 ? want : sm_count() * 32);
    if (grid < 1) grid = 1;
    if (ix.sa_wide) k_sa_densify<uint64_t><<<grid, 256, 0, st>>>(ix, sa_file, sa_intv, n_sa_file, (uint64_t *)out);
    else k_sa_densify<uint32_t><<<grid, 256, 0, st>>>(ix, sa_file, sa_intv, n_sa_file, (uint32_t *)out);
}

// RefSequence over both strands, 2 bits per base: position p < G is the .pac base, position p >= G is the
// complement of position 2G-1-p (/root/reference/src/bwt_index.cpp:199-209).
__device__ __forceinline__ uint32_t pac_base(const uint8_t *pac, int64_t p)
{
    return (pac[p >> 2] >> ((~p & 3) << 1)) & 3;
}

__global__ void k_build_ref2(const uint8_t *__restrict__ pac, uint32_t *__restrict__ ref2, int64_t G, int64_t n_words)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int i = 0; i < 16; i++) {
            int64_t p = w * 16 + i;
            uint32_t c = 0;
            if (p < G) c = pac_base(pac, p);
            else if (p < 2 * G) c = 3 - pac_base(pac, 2 * G - 1 - p);
            v |= c << (30 - 2 * i);
        }
        ref2[w] = v;
    }
}

void launch_build_ref2(const uint8_t *pac, uint32_t *ref2, int64_t G, cudaStream_t st)
{
    int64_t n_words = (2 * G + 15) / 16 + 2; // two guard words so 64-bit window reads never run off the end
    int64_t blocks = (n_words + 255) / 256;
    int grid = (int)(blocks < sm_count() * 16 ? blocks : sm_count() * 16);
    if (grid < 1) grid = 1;
    k_build_ref2<<<grid, 256, 0, st>>>(pac, ref2, G, n_words);
}


// ---------------------------------------------------------------------------------------------------
// read batch: raw ASCII -> device codes, on the device (the host only memcpy's the caller's bases into pinned memory)
// ---------------------------------------------------------------------------------------------------
// padded length of every read (each read starts on a 16-byte boundary so the search kernel stages 16 bases per load)
__global__ void k_read_layout(const int64_t *__restrict__ off, int n, int32_t *rlen, uint32_t *padded)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        if (i == n) { padded[i] = 0; continue; }
        int rl = (int)(off[i + 1] - off[i]);
        rlen[i] = rl;
        padded[i] = (uint32_t)((rl + 15) & ~15);
    }
}

// codes: 0..3 = ACGT, 8..11 = acgt, 5 = 'N', 4 = anything else (dartgpu_internal.h)
__global__ void k_encode_reads(const uint8_t *__restrict__ raw, const int64_t *__restrict__ off, const int64_t *__restrict__ dev_off,
                               int n, uint8_t *codes, uint2 *packed)
{
    __shared__ uint8_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint8_t v = 4;
        switch (i) {
        case 'A': v = 0; break; case 'C': v = 1; break; case 'G': v = 2; break; case 'T': v = 3; break;
        case 'a': v = 8; break; case 'c': v = 9; break; case 'g': v = 10; break; case 't': v = 11; break;
        case 'N': v = 5; break;
        }
        tab[i] = v;
    }
    __syncthreads();
    // 8 lanes per read (a 101-base read has 7 chunks of 16 bases; the first version gave it a whole warp), one chunk per
    // lane: five aligned 32-bit loads cover the 16 unaligned source bytes, funnel shifts realign them.
    const int gl = threadIdx.x & 7;
    const int ngroups = (gridDim.x * blockDim.x) >> 3;
    const int64_t base0 = off[0];
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < n; r += ngroups) {
        const int64_t src = off[r] - base0;
        const int rl = (int)(off[r + 1] - off[r]);
        const int chunks = (rl + 15) >> 4;
        const int64_t d0 = dev_off[r];
        uint4 *dst = reinterpret_cast<uint4 *>(codes + d0);
        for (int ch = gl; ch < chunks; ch += 8) {
            const int64_t s0 = src + 16 * ch;
            const uint32_t *wp = reinterpret_cast<const uint32_t *>(raw + (s0 & ~(int64_t)3));
            const int sh = (int)(s0 & 3) * 8;
            uint32_t in[5];
#pragma unroll
            for (int k = 0; k < 5; k++) in[k] = wp[k];
            uint32_t w[4] = {0, 0, 0, 0};
            uint32_t two = 0, amb = 0;      // the search kernel's view: 2 bits per base + one "not ACGT" bit
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint32_t x = __funnelshift_r(in[q], in[q + 1], sh);
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const int p = ch * 16 + q * 4 + k;
                    const uint32_t c = p < rl ? tab[(x >> (8 * k)) & 0xFFu] : 4u;
                    w[q] |= c << (8 * k);
                    two |= (c & 3u) << (2 * (q * 4 + k));
                    amb |= ((c >> 2) & 1u) << (q * 4 + k);
                }
            }
            dst[ch] = make_uint4(w[0], w[1], w[2], w[3]);
            packed[(d0 >> 4) + ch] = make_uint2(two, amb);
        }
    }
}

// Read-back of a few words (counts, totals) into pinned host memory WITHOUT the copy engine: the host pointer is device-
// accessible (unified addressing), one warp stores through it.  These read-backs gate the next kernel launches of a
// context; as cudaMemcpyAsync they queued on the single D2H copy engine behind other contexts' 40 MB result copies and
// stalled ~2 ms per step (host trace: contexts spent 6.5 ms in a 5 ms compute phase whenever result copies were on).
__global__ void k_small_d2h(uint32_t *__restrict__ host, const uint32_t *__restrict__ dev, int n_words)
{
    for (int i = threadIdx.x; i < n_words; i += blockDim.x) host[i] = dev[i];
    __threadfence_system();
}
void small_d2h(void *host_pinned, const void *dev, size_t bytes, cudaStream_t st)
{
    k_small_d2h<<<1, 32, 0, st>>>(reinterpret_cast<uint32_t *>(host_pinned), reinterpret_cast<const uint32_t *>(dev), (int)(bytes / 4));
}

void launch_read_layout(const int64_t *off, int n, int32_t *rlen, uint32_t *padded, cudaStream_t st)
{
    int grid = (n + 1 + 255) / 256; if (grid > sm_count() * 8) grid = sm_count() * 8;
    k_read_layout<<<grid, 256, 0, st>>>(off, n, rlen, padded);
}
void launch_encode_reads(const uint8_t *raw, const int64_t *off, const int64_t *dev_off, int n, uint8_t *codes, uint2 *packed, cudaStream_t st)
{
    int64_t want = ((int64_t)n * 8 + 255) / 256;
    int grid = (int)(want < sm_count() * 16 ? want : sm_count() * 16); if (grid < 1) grid = 1;
    k_encode_reads<<<grid, 256, 0, st>>>(raw, off, dev_off, n, codes, packed);
}

} // namespace dartgpu
