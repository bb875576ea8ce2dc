// Batched 8-mer re-seeding inside a (read gap, genome window) pair — replaces
// GenerateLongestSimplePairsFromFragmentPair and its helpers CreateKmerVecFromReadSeq / CreateKmerID /
// IdentifyCommonKmers (/root/reference/src/KmerAnalysis.cpp:25-166) bit-exactly.
//
// The reference builds both 8-mer lists, sorts them by id, joins them, sorts all common pairs by
// (PosDiff, rPos) and walks the equal-PosDiff runs with a counter `s` that is only reset when a run is
// accepted (KmerAnalysis.cpp:147-163).  Only three numbers of each run are used: its size, its first rPos
// and its last rPos.  So instead of materialising and sorting the pairs (up to ~500 k window positions):
//   * the read gap's 8-mers (<= ~250) are sorted once in shared memory (block bitonic sort);
//   * the block streams the genome window 1024 positions at a time straight from the HBM-resident 2-bit
//     reference, each thread forming four 8-mers (one funnel shift each, all loads in flight first) and looking
//     them up by four interleaved binary searches;
//   * matches update a shared-memory ring of per-diagonal {count, min rPos, max rPos} with shared-memory
//     atomics; a diagonal is final once the stream has passed it: all warps ballot which diagonals are occupied
//     (almost none are), warp 0 retires the occupied ones in increasing PosDiff order — exactly the order of the
//     reference's sorted walk — carrying `s` and `max_len` in registers.
// One block (8 warps) per job: windows reach 500 kb (MaxIntronSize) and a single warp streaming one is pure
// latency (round-1 profile: 1.7 ms for one 115 kb window).
// Algorithmic traffic: ceil(len2/4) bytes of reference + len1 bytes of read + 12 bytes of result per job.
//
// Quirks kept (only reachable with non-ACGT read symbols): only a literal 'N' breaks an 8-mer; other symbols
// add 4 into the rolling id; the first id after a (re)start is unmasked; after an 'N' restart the rolling
// window is one base late (KmerAnalysis.cpp:52, :60-75).  Those fragments take a sequential path on thread 0.
#include "dartgpu_internal.h"

namespace dartgpu {

constexpr unsigned FULLK = 0xffffffffu;
constexpr int KMER_THREADS = 256;
constexpr int KMER_PPT = 4;                       // window positions per thread and tile
constexpr int KMER_TILE = KMER_THREADS * KMER_PPT;

// nst_nt4_table value of a device read code (0..3 ACGT, 8..11 acgt, 4 other, 5 'N')
__device__ __forceinline__ uint32_t nt4(uint8_t c) { return (c & 4) ? 4u : (uint32_t)(c & 3); }

__device__ __forceinline__ uint32_t genome_kmer(const DevIndex &ix, int64_t p)
{
    const uint32_t *w = ix.ref2 + (p >> 4);
    uint64_t x = (uint64_t)__ldg(w) << 32 | __ldg(w + 1);
    return (uint32_t)(x >> (48 - 2 * (int)(p & 15))) & 0xFFFFu;
}

__global__ void __launch_bounds__(KMER_THREADS)
k_kmer(DevIndex ix, const uint8_t *__restrict__ codes, const KmerJobDev *__restrict__ jobs, int n_jobs,
       int tab_cap, int ring, dartgpu_kmer_hit *out)
{
    extern __shared__ uint32_t smem[];
    uint32_t *tab = smem;                       // tab_cap entries: id << 16 | read position
    uint32_t *cnt = tab + tab_cap;              // ring of per-diagonal aggregates
    uint32_t *rmin = cnt + ring, *rmax = rmin + ring;
    uint32_t *bal = rmax + ring;                // occupancy bitmaps of the diagonals being retired
    uint32_t *bm = bal + ring / 32 + 8;          // 64 Kbit presence bitmap of the read gap's 8-mer ids
    __shared__ int s_nk, s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const KmerJobDev J = jobs[job];
        const int L1 = J.len1, L2 = J.len2;
        int best_r = 0, best_g = 0, best_len = 0;
        __syncthreads();
        if (L1 >= 8 && L2 >= 8 && L1 <= tab_cap) {
            const uint8_t *s1 = codes + J.s1_off;
            // ---- 1. the read gap's 8-mer list ----
            if (tid == 0) { s_bad = 0; s_nk = 0; }
            __syncthreads();
            bool bad = false;
            for (int p = tid; p < L1; p += KMER_THREADS) bad |= (s1[p] & 4) != 0;
            if (bad) s_bad = 1;
            __syncthreads();
            if (!s_bad) {
                for (int p = tid; p < L1 - 7; p += KMER_THREADS) {
                    uint32_t id = 0;
#pragma unroll
                    for (int i = 0; i < 8; i++) id = (id << 2) | (s1[p + i] & 3u);
                    tab[p] = id << 16 | (uint32_t)p;
                }
                if (tid == 0) s_nk = L1 - 7;
            } else if (tid == 0) { // KmerAnalysis.cpp:34-80 step by step; ids above 16 bits can never match the genome
                int nk = 0, tail = 0, count = 0, head;
                while (count < 8 && tail < L1) { if (s1[tail++] != CODE_N) count++; else count = 0; }
                if (count == 8) {
                    head = tail - 8;
                    uint32_t wid = 0;
                    for (int i = head; i < head + 8; i++) wid = (wid << 2) + nt4(s1[i]);
                    if (wid <= 0xFFFFu) tab[nk++] = wid << 16 | (uint32_t)head;
                    for (head += 1; tail < L1; head++, tail++) {
                        if (s1[tail] != CODE_N) {
                            wid = ((wid & 0x3FFFu) << 2) + nt4(s1[tail]);
                            if (wid <= 0xFFFFu) tab[nk++] = wid << 16 | (uint32_t)head;
                        } else {
                            count = 0; tail++;
                            while (count < 8 && tail < L1) { if (s1[tail++] != CODE_N) count++; else count = 0; }
                            if (count != 8) break;
                            head = tail - 8;
                            wid = 0;
                            for (int i = head; i < head + 8; i++) wid = (wid << 2) + nt4(s1[i]);
                            if (wid <= 0xFFFFu) tab[nk++] = wid << 16 | (uint32_t)head;
                        }
                    }
                }
                s_nk = nk;
            }
            __syncthreads();
            const int nk = s_nk;
            if (nk > 0) {
                int NP = 32;
                while (NP < nk) NP <<= 1;
                for (int i = nk + tid; i < NP; i += KMER_THREADS) tab[i] = 0xFFFFFFFFu;
                __syncthreads();
                for (int k = 2; k <= NP; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int i = tid; i < NP; i += KMER_THREADS) {
                            int p = i ^ j;
                            if (p > i) {
                                uint32_t x = tab[i], y = tab[p];
                                bool up = (i & k) == 0;
                                if ((x > y) == up) { tab[i] = y; tab[p] = x; }
                            }
                        }
                        __syncthreads();
                    }
                for (int s = tid; s < ring; s += KMER_THREADS) { cnt[s] = 0; rmin[s] = 0xFFFFFFFFu; rmax[s] = 0; }
                for (int s = tid; s < 2048; s += KMER_THREADS) bm[s] = 0;
                __syncthreads();
                for (int i = tid; i < nk; i += KMER_THREADS) { const uint32_t w = tab[i] >> 16; atomicOr(&bm[w >> 5], 1u << (w & 31)); }
                __syncthreads();

                // ---- 2. stream the window KMER_TILE positions at a time; retire diagonals in increasing PosDiff order ----
                const int ngk = L2 - 7;
                const int dd_end = (ngk - 1) + L1 + 1; // one past the largest diagonal index (dd = gPos - rPos + L1)
                int s_acc = 1, max_len = 0;             // live in warp 0 only
                int fin = 8;
                for (int g0 = 0; g0 < ngk; g0 += KMER_TILE) {
                    // four positions per thread: all eight reference words are in flight before the first lookup.  A 64 Kbit
                    // presence bitmap of the read gap's 8-mers rejects almost every window position with one shared-memory
                    // load (round-1 ncu: binary-searching every position was half of the kernel's instructions).
                    uint32_t wid[KMER_PPT];
#pragma unroll
                    for (int i = 0; i < KMER_PPT; i++) {
                        const int g = g0 + tid + KMER_THREADS * i;
                        wid[i] = g < ngk ? genome_kmer(ix, J.gpos + g) : 0xFFFFFFFFu;
                    }
#pragma unroll
                    for (int i = 0; i < KMER_PPT; i++) {
                        const uint32_t w = wid[i];
                        if (w == 0xFFFFFFFFu || !((bm[w >> 5] >> (w & 31)) & 1u)) continue;
                        const int g = g0 + tid + KMER_THREADS * i;
                        int lo = 0, hi = nk;
                        while (lo < hi) { int mid = (lo + hi) >> 1; if (tab[mid] < (w << 16)) lo = mid + 1; else hi = mid; }
                        for (int e = lo; e < nk && (tab[e] >> 16) == w; e++) {
                            const uint32_t r = tab[e] & 0xFFFFu;
                            const int slot = (g - (int)r + L1) & (ring - 1);
                            atomicAdd(&cnt[slot], 1u); atomicMin(&rmin[slot], r); atomicMax(&rmax[slot], r);
                        }
                    }
                    __syncthreads();
                    const bool last_tile = g0 + KMER_TILE >= ngk;
                    const int fin_end = last_tile ? dd_end : min(g0 + KMER_TILE + 8, dd_end);
                    const int nchunks = (fin_end - fin + 31) >> 5;
                    for (int cb = warp; cb < nchunks; cb += KMER_THREADS / 32) {   // all warps: which 32-diagonal chunks are occupied
                        const int dd = fin + cb * 32 + lane;
                        const unsigned b = __ballot_sync(FULLK, dd < fin_end && cnt[dd & (ring - 1)] > 0);
                        if (lane == 0) bal[cb] = b;
                    }
                    __syncthreads();
                    if (warp == 0) {                                               // warp 0: the occupied ones, in order
                        for (int cb = 0; cb < nchunks; cb += 32) {
                            const unsigned mine = cb + lane < nchunks ? bal[cb + lane] : 0u;
                            unsigned nzc = __ballot_sync(FULLK, mine != 0);
                            while (nzc) {
                                const int ci = __ffs(nzc) - 1;
                                nzc &= nzc - 1;
                                unsigned word = __shfl_sync(FULLK, mine, ci);
                                const int base = fin + (cb + ci) * 32;
                                while (word) {
                                    const int src = __ffs(word) - 1;
                                    word &= word - 1;
                                    const int sl = (base + src) & (ring - 1);
                                    const int cc = (int)cnt[sl], mn = (int)rmin[sl], mx = (int)rmax[sl];
                                    s_acc += cc - 1;
                                    const int l = 8 + (mx - mn);
                                    if (l > max_len && s_acc > (l - 8) / 2) {
                                        best_r = mn; best_g = mn + (base + src - L1); best_len = l;
                                        max_len = l; s_acc = 1;
                                    }
                                    __syncwarp();
                                    if (lane == 0) { cnt[sl] = 0; rmin[sl] = 0xFFFFFFFFu; rmax[sl] = 0; }
                                }
                            }
                        }
                    }
                    fin = fin_end;
                    __syncthreads();
                }
            }
        }
        if (tid == 0) { out[job].rpos = best_r; out[job].gpos = best_g; out[job].len = best_len; }
    }
}

static int pow2_at_least(int v) { int p = 32; while (p < v) p <<= 1; return p; }

void launch_kmer(const DevIndex &ix, const uint8_t *codes, const KmerJobDev *jobs, int n_jobs, int max_len1,
                 dartgpu_kmer_hit *out, cudaStream_t st)
{
    if (n_jobs <= 0) return;
    int tab_cap = pow2_at_least(max_len1 < 8 ? 8 : max_len1);
    int ring = pow2_at_least(max_len1 + KMER_TILE + 32);
    size_t smem = (size_t)(tab_cap + 3 * ring + ring / 32 + 8 + 2048) * 4;
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaFuncSetAttribute(k_kmer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = smem;
    }
    int grid = n_jobs < 148 * 8 ? n_jobs : 148 * 8;
    k_kmer<<<grid, KMER_THREADS, smem, st>>>(ix, codes, jobs, n_jobs, tab_cap, ring, out);
}

} // namespace dartgpu
