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
#include <cstdlib>

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
k_kmer(DevIndex ix, const uint8_t *__restrict__ codes, const KmerJobDev *__restrict__ jobs,
       const uint32_t *__restrict__ job_list, const uint32_t *__restrict__ n_listed,
       int tab_cap, int ring, const BatchCtl *__restrict__ ctl, dartgpu_kmer_hit *out)
{
    if (ctl->abort) return;
    extern __shared__ uint32_t smem[];
    uint32_t *tab = smem;                       // tab_cap entries: id << 16 | read position
    uint32_t *cnt = tab + tab_cap;              // ring of per-diagonal aggregates
    uint32_t *rmin = cnt + ring, *rmax = rmin + ring;
    uint32_t *bal = rmax + ring;                // occupancy bitmaps of the diagonals being retired
    uint32_t *bm = bal + ring / 32 + 8;          // 64 Kbit presence bitmap of the read gap's 8-mer ids
    __shared__ int s_nk, s_bad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    const int n_list = (int)*n_listed;
    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        const int job = (int)job_list[li];
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

// =====================================================================================================
// Fast path: windows scanned by the whole grid, matches kept as records
// =====================================================================================================
// Round-1 profile (config[2] at full size: 16 k jobs / 1 M reads, windows of 37 kb on average, up to 500 kb): the
// block-per-job kernel above spends its time in the three block barriers and the diagonal retirement of every 1024-position
// tile (~6 us per tile, 75x off the kernel's memory roofline), and a 500 kb window is walked by a single CTA.  But a
// window position matches one of the gap's ~100 8-mers with probability ~0.15 %, so the matches themselves are few:
//   k_kmer_prep   one thread per job: classify, count the chunks (up to 32 tiles of 4096 positions) and the record capacity of the job;
//   k_kmer_scan   the grid walks the flattened tile space of ALL jobs (a long window is spread over many CTAs); a thread
//                 forms 16 consecutive 8-mers from three coalesced words, tests them against a 64 Kbit bitmap of the gap's
//                 8-mers and resolves the rare hits through a small hash table; every match becomes one 32-bit record
//                 (diagonal << 10 | read position) in the job's slice of a record pool — no barrier per tile;
//   k_kmer_walk   one CTA per job sorts its records (a few hundred) and walks the diagonals in increasing order with the
//                 reference's carried counter `s` (KmerAnalysis.cpp:147-163) — the same arithmetic as the ring kernel.
// Jobs the fast path does not take — non-ACGT symbols in the gap (the N quirks), more matches than the job's capacity
// (low-complexity gaps), windows beyond 2^22 — go through the ring kernel above, which has no such limits.
constexpr int KS_THREADS = 256, KS_PPT = 16, KS_TILE = KS_THREADS * KS_PPT, KS_CHUNK = 32;
constexpr int KS_HASH = 2048;                    // >= 2 x the most 8-mers a gap can have (DARTGPU_MAX_RLEN - 7)
constexpr uint32_t KS_CAP_MAX = 4096, KS_EMPTY = 0xFFFFFFFFu;

__global__ void k_kmer_prep(const uint8_t *__restrict__ codes, const KmerJobDev *__restrict__ jobs, const int32_t *__restrict__ n_jobs_p, int cap_jobs,
                            int tab_cap, uint32_t cap_max, uint32_t *ntiles, uint32_t *cap, uint32_t *count, uint32_t *heavy_list,
                            uint32_t *heavy_count, BatchCtl *ctl, dartgpu_kmer_hit *out)
{
    if (ctl->abort) return;
    const int n_jobs = min(*n_jobs_p, cap_jobs);
    unsigned long long w_sum = 0, r_sum = 0;
    // the two prefix sums behind this kernel run over the queue's CAPACITY (its length is only known on the device):
    // slots beyond the length count zero
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= cap_jobs; j += gridDim.x * blockDim.x) {
        if (j >= n_jobs) { ntiles[j] = 0; cap[j] = 0; continue; }
        const KmerJobDev J = jobs[j];
        const int L1 = J.len1, L2 = J.len2;
        w_sum += (unsigned long long)L2; r_sum += (unsigned long long)L1;
        uint32_t nt = 0, cp = 0;
        if (L1 >= 8 && L2 >= 8 && L1 <= tab_cap) {
            bool plain = L1 <= DARTGPU_MAX_RLEN && (int64_t)L1 + L2 < (1 << 22);
            const uint8_t *s1 = codes + J.s1_off;
            for (int p = 0; p < L1 && plain; p++) plain = (s1[p] & 4) == 0;
            const uint64_t expect = ((uint64_t)(L2 - 7) * (uint64_t)(L1 - 7)) >> 16;     // chance matches
            uint64_t want = 2 * expect + 2 * (uint64_t)L1 + 64;
            if (cap_max < KS_CAP_MAX && want > cap_max) want = cap_max;      // test hook: force overflows into the ring kernel
            if (plain && want <= KS_CAP_MAX) { nt = (uint32_t)((L2 - 7 + KS_TILE * KS_CHUNK - 1) / (KS_TILE * KS_CHUNK)); cp = (uint32_t)want; }
            else heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)j;
        } else { out[j].rpos = 0; out[j].gpos = 0; out[j].len = 0; }
        ntiles[j] = nt; cap[j] = cp; count[j] = 0;
    }
    for (int d = 16; d > 0; d >>= 1) { w_sum += __shfl_xor_sync(FULLK, w_sum, d); r_sum += __shfl_xor_sync(FULLK, r_sum, d); }
    if ((threadIdx.x & 31) == 0 && (w_sum | r_sum)) { atomicAdd(&ctl->work[1], w_sum); atomicAdd(&ctl->work[2], r_sum); }
}

__device__ __forceinline__ uint32_t ks_hash(uint32_t id) { return ((id * 40503u) >> 4) & (KS_HASH - 1); }

__global__ void __launch_bounds__(KS_THREADS)
k_kmer_scan(DevIndex ix, const uint8_t *__restrict__ codes, const KmerJobDev *__restrict__ jobs, const int32_t *__restrict__ n_jobs_p, int cap_jobs,
            const int64_t *__restrict__ chunk_off, const int64_t *__restrict__ rec_off, const uint32_t *__restrict__ cap,
            uint32_t *count, uint32_t *recs, const BatchCtl *__restrict__ ctl)
{
    if (ctl->abort) return;
    const int n_jobs = min(*n_jobs_p, cap_jobs);
    __shared__ uint32_t bm[2048];
    __shared__ uint32_t ht[KS_HASH];
    __shared__ int s_job;
    const int tid = threadIdx.x;
    const int64_t nchunks = chunk_off[n_jobs];
    // a chunk = up to KS_CHUNK consecutive tiles of ONE job, so the gap's tables are built once per chunk
    for (int64_t ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
        __syncthreads();
        if (tid == 0) {                      // the job that owns chunk ch: last j with chunk_off[j] <= ch
            int lo = 0, hi = n_jobs;
            while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (chunk_off[mid] <= ch) lo = mid; else hi = mid - 1; }
            s_job = lo;
        }
        for (int i = tid; i < 2048; i += KS_THREADS) bm[i] = 0;
        for (int i = tid; i < KS_HASH; i += KS_THREADS) ht[i] = KS_EMPTY;
        __syncthreads();
        const int job = s_job;
        const KmerJobDev J = jobs[job];
        const int L1 = J.len1, ngk = J.len2 - 7;
        // ---- the gap's 8-mers: presence bitmap + hash table (id << 16 | read position) ----
        const uint8_t *s1 = codes + J.s1_off;
        for (int p = tid; p < L1 - 7; p += KS_THREADS) {
            uint32_t id = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) id = (id << 2) | (s1[p + i] & 3u);
            atomicOr(&bm[id >> 5], 1u << (id & 31));
            uint32_t slot = ks_hash(id);
            while (atomicCAS(&ht[slot], KS_EMPTY, id << 16 | (uint32_t)p) != KS_EMPTY) slot = (slot + 1) & (KS_HASH - 1);
        }
        __syncthreads();
        const int t0 = (int)(ch - chunk_off[job]) * KS_CHUNK;
        const int t1 = min(t0 + KS_CHUNK, (ngk + KS_TILE - 1) / KS_TILE);
        const uint32_t my_cap = cap[job];
        uint32_t *my_recs = recs + rec_off[job];
        // the three words of the next tile are in flight while the current one is matched
        auto fetch = [&](int tt, uint32_t &a0, uint32_t &a1, uint32_t &a2) {
            const int gf = tt * KS_TILE + tid * KS_PPT;
            if (tt < t1 && gf < ngk) {
                const uint32_t *w = ix.ref2 + ((J.gpos + gf) >> 4);
                a0 = __ldg(w); a1 = __ldg(w + 1); a2 = __ldg(w + 2);
            }
        };
        uint32_t w0 = 0, w1 = 0, w2 = 0, n0 = 0, n1 = 0, n2 = 0;
        fetch(t0, w0, w1, w2);
        for (int t = t0; t < t1; t++) {
            fetch(t + 1, n0, n1, n2);
            const int g_first = t * KS_TILE + tid * KS_PPT;
            if (g_first < ngk) {
                // align the 96-bit window to the thread's first position once; the 16 8-mers then sit at fixed offsets
                const int sh = 2 * (int)((J.gpos + g_first) & 15);
                const uint32_t X0 = __funnelshift_l(w1, w0, sh), X1 = __funnelshift_l(w2, w1, sh);
                const int left = ngk - g_first;                      // positions of this thread that are inside the window
                // branch-free presence test of the 16 8-mers (ncu: with a branch per position the test was 56 % of the
                // kernel's instructions); the rare hits are resolved afterwards
                uint32_t hits = 0;
#pragma unroll
                for (int i = 0; i < KS_PPT; i++) {
                    const uint32_t id = (i <= 8 ? X0 >> (16 - 2 * i) : __funnelshift_l(X1, X0, 2 * i) >> 16) & 0xFFFFu;
                    hits |= ((bm[id >> 5] >> (id & 31)) & 1u) << i;
                }
                if (left < KS_PPT) hits &= (1u << left) - 1u;
                while (hits) {
                    const int i = __ffs(hits) - 1;
                    hits &= hits - 1;
                    const uint32_t id = (__funnelshift_l(X1, X0, 2 * i) >> 16) & 0xFFFFu;
                    uint32_t slot = ks_hash(id), e;
                    while ((e = ht[slot]) != KS_EMPTY) {
                        if ((e >> 16) == id) {
                            const uint32_t r = e & 0xFFFFu;
                            const uint32_t dd = (uint32_t)(g_first + i - (int)r + L1);
                            const uint32_t k = atomicAdd(&count[job], 1u);
                            if (k < my_cap) my_recs[k] = dd << 10 | r;
                        }
                        slot = (slot + 1) & (KS_HASH - 1);
                    }
                }
            }
            w0 = n0; w1 = n1; w2 = n2;
        }
    }
}

__global__ void __launch_bounds__(128)
k_kmer_walk(const KmerJobDev *__restrict__ jobs, const int32_t *__restrict__ n_jobs_p, int cap_jobs, const int64_t *__restrict__ rec_off,
            const uint32_t *__restrict__ cap, const uint32_t *__restrict__ count, const uint32_t *__restrict__ recs, uint32_t *heavy_list,
            uint32_t *heavy_count, const BatchCtl *__restrict__ ctl, dartgpu_kmer_hit *out)
{
    if (ctl->abort) return;
    const int n_jobs = min(*n_jobs_p, cap_jobs);
    __shared__ uint32_t sm[KS_CAP_MAX];
    const int tid = threadIdx.x;
    for (int job = blockIdx.x; job < n_jobs; job += gridDim.x) {
        const uint32_t cp = cap[job];
        if (cp == 0) continue;                                   // not a fast-path job
        const uint32_t n = count[job];
        if (n > cp) { if (tid == 0) heavy_list[atomicAdd(heavy_count, 1u)] = (uint32_t)job; continue; }
        int N = 32;
        while (N < (int)n) N <<= 1;
        __syncthreads();
        const uint32_t *src = recs + rec_off[job];
        for (int i = tid; i < N; i += 128) sm[i] = i < (int)n ? src[i] : 0xFFFFFFFFu;
        __syncthreads();
        for (int k = 2; k <= N; k <<= 1)
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < N; i += 128) {
                    const int p = i ^ j;
                    if (p > i) {
                        const uint32_t x = sm[i], y = sm[p];
                        const bool up = (i & k) == 0;
                        if ((x > y) == up) { sm[i] = y; sm[p] = x; }
                    }
                }
                __syncthreads();
            }
        if (tid == 0) {
            const int L1 = jobs[job].len1;
            int best_r = 0, best_g = 0, best_len = 0, s_acc = 1, max_len = 0;
            uint32_t i = 0;
            while (i < n) {                                      // one run of equal PosDiff: records sorted by read position
                const uint32_t dd = sm[i] >> 10;
                const int mn = (int)(sm[i] & 1023u);
                uint32_t e = i + 1;
                while (e < n && (sm[e] >> 10) == dd) e++;
                const int mx = (int)(sm[e - 1] & 1023u);
                s_acc += (int)(e - i) - 1;
                const int l = 8 + (mx - mn);
                if (l > max_len && s_acc > (l - 8) / 2) {
                    best_r = mn; best_g = mn + ((int)dd - L1); best_len = l;
                    max_len = l; s_acc = 1;
                }
                i = e;
            }
            out[job].rpos = best_r; out[job].gpos = best_g; out[job].len = best_len;
        }
    }
}


void launch_kmer(const DevIndex &ix, const uint8_t *codes, const KmerJobDev *jobs, const int32_t *n_jobs, int cap_jobs, int max_len1,
                 dartgpu_kmer_hit *out, KmerScratch &S, BatchCtl *ctl, int64_t cap_recs, cudaStream_t st)
{
    if (cap_jobs <= 0) return;
    const int sms = sm_count();
    const int tab_cap = pow2_at_least(max_len1 < 8 ? 8 : max_len1);
    S.ntiles.reserve(cap_jobs + 2); S.cap.reserve(cap_jobs + 2); S.count.reserve(cap_jobs + 2);
    S.tile_off.reserve(cap_jobs + 2); S.rec_off.reserve(cap_jobs + 2);
    S.heavy_list.reserve(cap_jobs + 2);
    S.recs.reserve((size_t)cap_recs + 1);
    const size_t tmp = scan_tmp_bytes(cap_jobs);
    S.scan_tmp.reserve(tmp + 256);
    uint32_t *heavy_count = &ctl->heavy_count;                  // zeroed with the control block
    int g1 = (cap_jobs + 1 + 255) / 256; if (g1 > sms * 8) g1 = sms * 8;
    const uint32_t cap_max = getenv("DARTGPU_KMER_CAP") ? (uint32_t)atoi(getenv("DARTGPU_KMER_CAP")) : KS_CAP_MAX;
    k_kmer_prep<<<g1, 256, 0, st>>>(codes, jobs, n_jobs, cap_jobs, tab_cap, cap_max, S.ntiles.p, S.cap.p, S.count.p, S.heavy_list.p,
                                    heavy_count, ctl, out);
    launch_scan_u32_to_i64(S.ntiles.p, S.tile_off.p, cap_jobs, S.scan_tmp.p, tmp, st);
    launch_scan_u32_to_i64(S.cap.p, S.rec_off.p, cap_jobs, S.scan_tmp.p, tmp, st);
    launch_ctl_check(ctl, &ctl->kmer_recs, S.rec_off.p + cap_jobs, cap_recs, CAP_KRECS, st);
    k_kmer_scan<<<sms * 8, KS_THREADS, 0, st>>>(ix, codes, jobs, n_jobs, cap_jobs, S.tile_off.p, S.rec_off.p, S.cap.p, S.count.p, S.recs.p, ctl);
    k_kmer_walk<<<cap_jobs < sms * 16 ? cap_jobs : sms * 16, 128, 0, st>>>(jobs, n_jobs, cap_jobs, S.rec_off.p, S.cap.p, S.count.p, S.recs.p,
                                                                          S.heavy_list.p, heavy_count, ctl, out);
    // whatever the fast path declined: the ring kernel, reading the list's length on the device
    const int ring = pow2_at_least(max_len1 + KMER_TILE + 32);
    const size_t smem = (size_t)(tab_cap + 3 * ring + ring / 32 + 8 + 2048) * 4;
    if (smem > 48 * 1024) {          // the opt-in is per device and cheap: set it on the device this launch goes to
        DG_CUDA(cudaFuncSetAttribute(k_kmer, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int grid = cap_jobs < sms * 4 ? cap_jobs : sms * 4;
    k_kmer<<<grid, KMER_THREADS, smem, st>>>(ix, codes, jobs, S.heavy_list.p, heavy_count, tab_cap, ring, ctl, out);
}

} // namespace dartgpu
