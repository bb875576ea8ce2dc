// Seeding on the GPU: FM-index exact-match segment search, SA locate, per-read seed sort and
// simple-pair clustering into alignment candidates.
//
// Replaces (bit-exactly) /root/reference/src/bwt_search.cpp:43-182 (bwt_occ4 / bwt_2occ4 / bwt_occ /
// bwt_invPsi / bwt_sa / BWT_Search) and /root/reference/src/AlignmentCandidates.cpp:181-288
// (IdentifySeedPairs, GenerateAlignmentCandidate).
//
// Execution model
//   * ONE THREAD owns one FM chain.  A rank query is one 256-bit load of a single 32-byte sector (Occ32, rank.cuh: the
//     four counts and the two bit-planes of 64 symbols), one masked popcount and no cross-lane traffic.  Round-1 history: the first kernels ranked a 64-byte block with 4 cooperating lanes
//     (4 x 128-bit loads, popcounts combined with shuffles); ncu showed them bound by the integer pipes (~100
//     instructions per LANE per step, i.e. ~400 thread-instructions per step) with the 4.6 MB table L2-resident, and by
//     too few independent chains in flight once the table lives in HBM.  Thread-per-chain needs ~6x fewer
//     thread-instructions per step, puts 4x more chains in flight per SM and touches half the bytes per query.
//   * the nested loops of the reference (reads > search starts > extension steps) are flattened into a single loop whose
//     every iteration issues one 256-bit load per lane — a rank step or a search-start table lookup — so the 32 lanes of
//     a warp stay converged although their reads, search starts and match lengths differ (states in k_search).
//   * reads arrive 2-bit packed (+1 bit "not ACGT"), 16 bases per 8-byte entry, written by k_encode_reads.
//   * search only records SA intervals (of the reverse-complemented pattern, see k_search); after a prefix sum over
//     the hit counts a second kernel resolves every hit with its own thread (LF walk to the next kept SA entry — none at
//     all when the full suffix array is resident — then the mirror p -> 2G - p - len), so repetitive reads do not
//     serialise.
//   * seeds are 64-bit keys (gPos | rPos | len): sorting the keys is the reference's CompByGenomePos order.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>
#include <atomic>
#include <cstdlib>

#include "dartgpu_internal.h"
#include "rank.cuh"

namespace dartgpu {

constexpr int SEARCH_THREADS = 128;
constexpr unsigned FULL = 0xffffffffu;

// One Occ32 block = one 256-bit load (sm_100: LDG.E.256), not cached in L1 (the table is touched at random).  The first
// version issued a 128-bit load for the planes and a 32-bit load for the count: ncu showed the kernel bound by L1TEX tag
// throughput (4 fully divergent load instructions per step), so the loads per step were cut to one per distinct block.
struct OccBlock { uint32_t cnt[4]; uint64_t lo, hi; };
__device__ __forceinline__ OccBlock load_block(const char *occ, uint64_t blk)
{
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7) : "l"(occ + blk * 32));
    OccBlock b;
    b.cnt[0] = r0; b.cnt[1] = r1; b.cnt[2] = r2; b.cnt[3] = r3;
    b.lo = (uint64_t)r5 << 32 | r4; b.hi = (uint64_t)r7 << 32 | r6;
    return b;
}
// Occ(c, kk): occurrences of c in B[0..kk] (kk already adjusted for the primary row), bwt_occ (bwt_search.cpp:43-65)
__device__ __forceinline__ uint32_t block_rank(const OccBlock &b, uint32_t t, int c)
{
    const uint32_t cnt = (c & 2) ? ((c & 1) ? b.cnt[3] : b.cnt[2]) : ((c & 1) ? b.cnt[1] : b.cnt[0]);
    return cnt + occ32_eq_upto(b.lo, b.hi, c, t);
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// ---------------------------------------------------------------------------------------------------
// kernel 1: maximal exact-match segment search (IdentifySeedPairs' loop around BWT_Search, without locate)
//
// One forward-extension step per loop iteration (bwt_2occ4 + interval update, bwt_search.cpp:152-170).  The reference
// carries the bi-interval (x0 = SA interval of the pattern P, x1 = SA interval of revcomp(P), x2 = size) and locates from
// x0.  Because the text is its own reverse complement (forward strand followed by the reverse-complement strand), the
// occurrences of revcomp(P) are the occurrences of P mirrored: p -> 2G - p - |P|.  The seeds are sorted by (gPos,rPos)
// right after (IdentifySeedPairs), so locating from x1 and mirroring yields the identical seed list — and extending P
// forward by base b is then a plain backward step of revcomp(P) with c = 3 - b:
//     n2   = Occ(c,l) - Occ(c,k)          size of the new interval          (k = x1-1, l = x1-1+x2)
//     x1'  = L2[c] + 1 + Occ(c,k)         its start
// i.e. one "equal to c" popcount per query and no forward-interval bookkeeping at all.  Interval widths are < 2^32
// (checked at index load); IdxT = uint32_t when the whole text fits 32 bits (every tested genome and BASELINE configs
// 1, 2, 5), uint64_t otherwise (3.1 Gbp: 2G = 6.2e9) — same code, narrower interval arithmetic.
//
// Reads are handed out dynamically: a CTA first drains its own contiguous range through a shared-memory counter, then
// steals single reads from a global counter that covers the last part of the batch, so that no lane idles while another
// still has a queue (the grid is exactly one wave).  Results are indexed by read, so the order does not matter.
// The work counters keep the reference's definition (a step touches one 64-byte BWA block or two, bwt_search.cpp:88-93).
// ---------------------------------------------------------------------------------------------------
// COUNT: keep the reference's work counters (steps, 64-byte blocks, sectors requested).  They cost ~5 of the ~45 instructions
// of a step on a kernel that is bound by instruction issue, so the whole-path calls run without them; the stage entry points
// (and DARTGPU_STATS=1) run with them — that is where the tests compare them with the oracle's and where bench.py reads the
// algorithmic bytes of the roofline.
template <typename IdxT, bool COUNT>
__global__ void __launch_bounds__(SEARCH_THREADS)
k_search(DevIndex ix, SeedLaunch a, int per_cta)
{
    __shared__ uint64_t s_L2[5];
    __shared__ int s_next;
    if (threadIdx.x == 0) { s_L2[0] = ix.L2[0]; s_L2[1] = ix.L2[1]; s_L2[2] = ix.L2[2]; s_L2[3] = ix.L2[3]; s_L2[4] = ix.L2[4]; }
    if (threadIdx.x == 0) s_next = blockIdx.x * per_cta;
    if (threadIdx.x == 0 && blockIdx.x == 0) a.nhits[a.n_reads] = 0;     // terminator of the prefix sum over the hit counts
    __syncthreads();
    const int r_end = (blockIdx.x + 1) * per_cta;            // own range; [steal_base, n_reads) is shared by everyone
    const IdxT primary = (IdxT)ix.primary;
    const char *occ = reinterpret_cast<const char *>(ix.occ32);
    const int K = ix.ktab ? ix.ktab_k : 0;
    const uint32_t kmask2 = K >= 16 ? ~0u : (1u << (2 * K)) - 1u, kmask1 = (1u << K) - 1u;

    // A lane is in one of four states.  STEP: one rank step per iteration.  JUMP: the search has a start whose first K bases
    // can be looked up in the search-start table; the lookup is ONE 256-bit load like a rank step's and rides the same load
    // instruction, so its latency is hidden exactly like a step's.  NEED: the read is exhausted; fetching the next one
    // (two atomics' worth of hand-out and a chain of dependent loads) is the only long turn-around left, and lanes wait
    // until at least TURN_BATCH of them need it (or nobody is stepping) so that the warp pays for it together.
    // Everything else a finished segment needs — record it, advance the start, skip ambiguous bases, cut the next K-mer
    // out of the packed read — is register arithmetic and runs at once.
    // History (ncu, round 1): with the whole turn-around inline 14 of 32 lanes were active; batching all of it (start-table
    // gather included) cut the time by a quarter but left lanes waiting 60 % of the iterations, because after every
    // mismatch a read goes through several searches of only 2-4 steps.
    enum { ST_STEP = 0, ST_JUMP = 1, ST_NEED = 2, ST_DONE = 3, ST_END = 4 };
    const char *ktab = reinterpret_cast<const char *>(ix.ktab);
    const int TURN_BATCH = a.turn_batch, END_BATCH = a.end_batch;
    bool own = true, have_read = false;
    int st = ST_NEED;
    int r = 0, rl = 0, start = 0, p = 0, cw = -2;
    int64_t wbase = 0;
    uint32_t nr = 0, nh = 0, x2 = 0, wcode = 0, wamb = 0, ncode = 0, namb = 0, kidx = 0;
    IdxT x1 = 0;
    uint32_t st_steps = 0, st_splits = 0, st_loads = 0;

    // The packed view of the read under the cursor: word cw (16 bases) and, prefetched, word cw + 1.  Round-2 finding: the
    // step path loaded the next word when it crossed a 16-base boundary, and the ambiguity bit of that word gates the rank
    // load — two dependent L2 round trips in that iteration.  Now the word behind the cursor is requested 16 steps before it is
    // needed.  Measured on config[1]: 1.72 -> 1.71 ms per 2 M reads — no difference: on the L2-resident index the kernel is
    // bound by instruction issue (68-78 % issue-active at 19.5 of 32 lanes), not by latency.  Kept: it is the cheaper code.
    auto fetch = [&](int w) {
        if (w == cw) return;
        if (w == cw + 1) { wcode = ncode; wamb = namb; }
        else { const uint2 x = __ldg(a.packed + wbase + w); wcode = x.x; wamb = x.y; }
        const uint2 y = __ldg(a.packed + wbase + w + 1);          // one entry past the read is padding or the next read: never used
        ncode = y.x; namb = y.y;
        cw = w;
    };
    // next search start of the current read: skip ambiguous bases, then either a table lookup (JUMP) or the reference's
    // single-base start (STEP); NEED when the read has no start left (IdentifySeedPairs' `pos < rlen - 13`)
    auto advance = [&]() {
        while (start < rl - 13) {              // a search cannot start on an ambiguous base
            fetch(start >> 4);
            if (!((wamb >> (start & 15)) & 1u)) break;
            start++;
        }
        if (start >= rl - 13) { st = ST_NEED; return; }
        if (K > 0 && start + K <= rl) {
            const int sh = start & 15;
            uint32_t bits = wcode >> (2 * sh), ambs = wamb >> sh;
            if (sh + K > 16) { bits |= sh ? ncode << (32 - 2 * sh) : 0u; ambs |= namb << (16 - sh); }
            if ((ambs & kmask1) == 0) { kidx = bits & kmask2; st = ST_JUMP; return; }
        }
        const int c0 = (wcode >> ((start & 15) * 2)) & 3;
        x1 = (IdxT)s_L2[3 - c0] + 1; x2 = (uint32_t)(s_L2[c0 + 1] - s_L2[c0]);
        p = start + 1;
        st = ST_STEP;
    };

    for (;;) {
        const unsigned need_m = __ballot_sync(FULL, st == ST_NEED);
        const unsigned end_m = __ballot_sync(FULL, st == ST_END);
        const unsigned act_m = __ballot_sync(FULL, st == ST_STEP || st == ST_JUMP);
        if ((need_m | act_m | end_m) == 0) break;
        // a finished segment: record it, advance (register arithmetic, but ~130 instructions: a few lanes share it)
        if (st == ST_END && (__popc(end_m) >= END_BATCH || act_m == 0)) {       // bwt_search.cpp:173 and IdentifySeedPairs' advance
            const int len = p - start;
            if (x2 <= a.max_dup && len >= 16) {
                if ((int)nr < a.cap_rec) {
                    SearchRec rec; rec.sa_begin = x1; rec.freq = x2; rec.start = (uint16_t)start; rec.len = (uint16_t)len;
                    a.recs[(int64_t)r * a.cap_rec + nr] = rec;
                }
                nr++; nh += x2;
                start += len;
            } else start++;
            advance();
        }
        if (st == ST_NEED && (__popc(need_m) >= TURN_BATCH || (act_m | end_m) == 0)) {
            if (have_read) { a.nrec[r] = nr; a.nhits[r] = nh; have_read = false; }
            if (own) { r = atomicAdd(&s_next, 1); if (r >= r_end) own = false; }
            if (!own) r = a.steal_base + (int)atomicAdd(a.steal, 1u);
            if (r >= a.n_reads) st = ST_DONE;
            else {
                rl = a.rlen[r];
                wbase = a.dev_off[r] >> 4;
                start = 0; nr = 0; nh = 0; have_read = true; cw = -2;
                advance();
            }
        }
        if (st == ST_STEP || st == ST_JUMP) {
            const bool stepping = st == ST_STEP;
            bool end = false, two = false;
            const char *addr0 = ktab + (size_t)(kidx >> 1) * 32, *addr1 = addr0;
            IdxT kk = 0, ll = 0;
            int c = 0;
            if (stepping) {
                end = p >= rl;
                if (!end) {
                    fetch(p >> 4);
                    end = (wamb >> (p & 15)) & 1u;
                }
                if (!end) {
                    const IdxT k = x1 - 1, l = k + x2;
                    kk = k - (k >= primary); ll = l - (l >= primary);
                    c = 3 - (int)((wcode >> ((p & 15) * 2)) & 3u);       // complement: backward step of revcomp(P)
                    const uint64_t bk = (uint64_t)kk >> 6, bl = (uint64_t)ll >> 6;
                    addr0 = occ + bk * 32; addr1 = occ + bl * 32;
                    two = bl != bk;                                       // narrow intervals sit in one block: one load
                }
            }
            if (!end) {
                const OccBlock B0 = load_block(addr0, 0);
                OccBlock B1 = B0;
                if (two) B1 = load_block(addr1, 0);
                if (COUNT) st_loads += two ? 2u : 1u;                     // 32-byte sectors really requested (L2 roofline)
                if (stepping) {
                    const uint32_t ok = block_rank(B0, (uint32_t)kk & 63u, c), ol = block_rank(B1, (uint32_t)ll & 63u, c);
                    if (COUNT) { st_steps++; st_splits += (uint32_t)(((kk ^ ll) >> 7) != 0); }
                    const uint32_t n2 = ol - ok;
                    if (n2 == 0) end = true;
                    else { x1 = (IdxT)s_L2[c] + 1 + ok; x2 = n2; p++; }
                } else {
                    // the 32-byte load holds two 16-byte KmerStart entries: {u64 x1, u32 x2, u32 splits}
                    const bool hi = kidx & 1u;
                    const uint64_t e_x1 = hi ? B0.lo : ((uint64_t)B0.cnt[1] << 32 | B0.cnt[0]);
                    const uint32_t e_x2 = hi ? (uint32_t)B0.hi : B0.cnt[2];
                    const uint32_t e_sp = hi ? (uint32_t)(B0.hi >> 32) : B0.cnt[3];
                    if (e_x2 != 0) {
                        x1 = (IdxT)e_x1; x2 = e_x2; p = start + K;
                        if (COUNT) { st_steps += K - 1; st_splits += e_sp; }
                    } else {                        // the K-mer does not occur: start from the single base as the reference does
                        const int c0 = (int)(kidx & 3u);
                        x1 = (IdxT)s_L2[3 - c0] + 1; x2 = (uint32_t)(s_L2[c0 + 1] - s_L2[c0]);
                        p = start + 1;
                    }
                    st = ST_STEP;
                }
            }
            if (end) st = ST_END;
        }
    }
    if (have_read) { a.nrec[r] = nr; a.nhits[r] = nh; }
    if (!COUNT) return;
    __syncwarp();
    const unsigned long long ws = warp_sum(st_steps), wp = warp_sum(st_splits), wl = warp_sum(st_loads);
    if ((threadIdx.x & 31) == 0 && ws) {
        atomicAdd(&a.stats->ext_steps, ws);
        atomicAdd(&a.stats->ext_blocks, ws + wp);
        atomicAdd(&a.stats->sector_loads, wl);
    }
}

static bool fits32(const DevIndex &ix) { return !ix.sa_wide; }

// occupancy of the two instantiations, per device (a process drives several GPUs from several host threads)
struct SearchCfg { std::atomic<int> occ32{0}, occ64{0}; };
static SearchCfg g_search_cfg[64];

void launch_search(const DevIndex &ix, SeedLaunch a, cudaStream_t st)
{
    if (a.n_reads <= 0) return;
    // exactly one wave: as many CTAs as are resident at once (a partial second wave would idle most SMs at the end)
    int dev = 0;
    cudaGetDevice(&dev);
    SearchCfg &cfg = g_search_cfg[dev >= 0 && dev < 64 ? dev : 0];
    int occ32 = cfg.occ32.load(std::memory_order_acquire), occ64 = cfg.occ64.load(std::memory_order_acquire);
    if (!occ32 || !occ64) {
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, k_search<uint32_t, true>, SEARCH_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, k_search<uint64_t, true>, SEARCH_THREADS, 0);
        occ32 = std::max(1, occ32); occ64 = std::max(1, occ64);
        cfg.occ32.store(occ32, std::memory_order_release); cfg.occ64.store(occ64, std::memory_order_release);
    }
    const int sms = sm_count();
    const bool narrow = fits32(ix);
    int want = (a.n_reads + SEARCH_THREADS - 1) / SEARCH_THREADS;
    int grid = sms * (narrow ? occ32 : occ64);
    if (want < grid) grid = want;
    // 3/4 of the batch is split statically between the CTAs, the last quarter is stolen read by read
    const int per_cta = (int)((int64_t)a.n_reads * 3 / 4 / grid);
    a.steal_base = per_cta * grid;
    static const int turn_batch = getenv("DARTGPU_TURN_BATCH") ? atoi(getenv("DARTGPU_TURN_BATCH")) : 6;
    a.turn_batch = turn_batch;
    static const int end_batch = getenv("DARTGPU_END_BATCH") ? atoi(getenv("DARTGPU_END_BATCH")) : 4;
    a.end_batch = end_batch;
    if (a.count_work) {
        if (narrow) k_search<uint32_t, true><<<grid, SEARCH_THREADS, 0, st>>>(ix, a, per_cta);
        else k_search<uint64_t, true><<<grid, SEARCH_THREADS, 0, st>>>(ix, a, per_cta);
    } else {
        if (narrow) k_search<uint32_t, false><<<grid, SEARCH_THREADS, 0, st>>>(ix, a, per_cta);
        else k_search<uint64_t, false><<<grid, SEARCH_THREADS, 0, st>>>(ix, a, per_cta);
    }
}

// ---------------------------------------------------------------------------------------------------
// prefix sums (plumbing): seed_off = exclusive scan of nhits over n_reads+1 entries (nhits[n_reads] == 0)
// ---------------------------------------------------------------------------------------------------
struct U32ToI64 { __host__ __device__ int64_t operator()(const uint32_t &v) const { return (int64_t)v; } };

size_t scan_tmp_bytes(int n)
{
    // the size query walks the dispatch layer of cub (device attributes, kernel attributes): ~10 us of host time, paid for
    // every prefix sum of every batch on the one host thread that drives a GPU -- remember the answers
    static thread_local int last_n[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    static thread_local size_t last_bytes[8];
    static thread_local int next = 0;
    for (int i = 0; i < 8; i++) if (last_n[i] == n) return last_bytes[i];
    size_t bytes = 0;
    cub::TransformInputIterator<int64_t, U32ToI64, const uint32_t *> it((const uint32_t *)nullptr, U32ToI64());
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (int64_t *)nullptr, n + 1);
    last_n[next] = n; last_bytes[next] = bytes; next = (next + 1) & 7;
    return bytes;
}

void launch_scan_u32_to_i64(const uint32_t *in, int64_t *out, int n, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    cub::TransformInputIterator<int64_t, U32ToI64, const uint32_t *> it(in, U32ToI64());
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, out, n + 1, st);
}

void launch_scan_hits(const SeedLaunch &a, void *tmp, size_t tmp_bytes, cudaStream_t st)
{
    launch_scan_u32_to_i64(a.nhits, a.seed_off, a.n_reads, tmp, tmp_bytes, st);
}

// ---------------------------------------------------------------------------------------------------
// kernel 2a: expand the recorded SA intervals into one slot per hit (slot = SA index still to resolve)
// ---------------------------------------------------------------------------------------------------
__global__ void k_expand(SeedLaunch a)
{
    if (a.ctl->abort) return;
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    for (; r < a.n_reads; r += gridDim.x * blockDim.x) {
        int64_t o = a.seed_off[r];
        int nr = min((int)a.nrec[r], a.cap_rec);
        for (int i = 0; i < nr; i++) {
            SearchRec rec = a.recs[(int64_t)r * a.cap_rec + i];
            uint32_t m = (uint32_t)rec.start << 16 | rec.len;
            for (uint32_t j = 0; j < rec.freq; j++, o++) { a.keys[o] = rec.sa_begin + j; a.meta[o] = m; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// kernel 2b: SA locate — bwt_sa / bwt_invPsi (bwt_search.cpp:119-137), one thread per hit
// ---------------------------------------------------------------------------------------------------
// With the full suffix array resident (sa_shift = 0) this is a single gather per hit.  With a sparser array a hit walks
// LF steps until it reaches a kept entry (geometrically distributed, mean 2^sa_shift - 1): one loop, one LF step per
// iteration, a lane picks up its next hit the moment it finishes one so the warp stays busy.
template <typename IdxT, typename SaT>
__global__ void __launch_bounds__(256)
k_locate(DevIndex ix, SeedLaunch a)
{
    if (a.ctl->abort) return;
    const int64_t total = a.ctl->total_seeds;
    __shared__ uint64_t s_L2[5];
    if (threadIdx.x == 0) { s_L2[0] = ix.L2[0]; s_L2[1] = ix.L2[1]; s_L2[2] = ix.L2[2]; s_L2[3] = ix.L2[3]; s_L2[4] = ix.L2[4]; }
    __syncthreads();
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
    const IdxT primary = (IdxT)ix.primary, sa_mask = (IdxT)ix.sa_mask;
    const char *occ = reinterpret_cast<const char *>(ix.occ32);
    const SaT *sa = reinterpret_cast<const SaT *>(ix.sa);
    unsigned long long st_lf = 0, st_hits = 0;
    int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool walking = false;
    IdxT k = 0;
    uint32_t steps = 0;
    for (;;) {
        if (!walking) {
            if (s >= total) break;
            k = (IdxT)a.keys[s];
            steps = 0;
            walking = true;
        }
        if (k & sa_mask) {                  // one LF step: the symbol at k and its rank come from the same sector
            steps++;
            if (k == primary) k = 0;
            else {
                const IdxT kk = k - (k > primary);
                const OccBlock B = load_block(occ, (uint64_t)kk >> 6);
                const uint32_t t = (uint32_t)kk & 63u;
                const int c = occ32_symbol(B.lo, B.hi, t);
                k = (IdxT)s_L2[c] + block_rank(B, t, c);
            }
        } else {
            // position of revcomp(P) on the text (entry 0 stands for -1, bwt_index.cpp:31); P itself starts at the
            // mirrored coordinate 2G - p' - |P|
            const uint64_t prc = (uint64_t)steps + (k == 0 ? ~0ull : (uint64_t)__ldg(sa + ((uint64_t)k >> ix.sa_shift)));
            const uint32_t m = a.meta[s];
            a.keys[s] = seed_key(2 * (uint64_t)ix.G - prc - (m & 0xFFFF), m >> 16, m & 0xFFFF);
            st_lf += steps; st_hits++;
            s += nthreads;
            walking = false;
        }
    }
    __syncwarp();
    st_lf = warp_sum(st_lf); st_hits = warp_sum(st_hits);
    if ((threadIdx.x & 31) == 0 && st_hits) {
        atomicAdd(&a.stats->lf_steps, st_lf);
        atomicAdd(&a.stats->hits, st_hits);
        atomicAdd(&a.stats->seeds, st_hits);
    }
}

void launch_expand_locate(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st)
{
    if (a.n_reads <= 0) return;
    const int cap = sm_count() * 8;
    int grid = (a.n_reads + 255) / 256;
    if (grid > cap) grid = cap;
    k_expand<<<grid, 256, 0, st>>>(a);
    // the hit count lives on the device (a.ctl->total_seeds): a full grid, threads beyond the count leave at once
    if (ix.sa_wide) k_locate<uint64_t, uint64_t><<<cap, 256, 0, st>>>(ix, a);   // sa_wide <=> 64-bit intervals
    else k_locate<uint32_t, uint32_t><<<cap, 256, 0, st>>>(ix, a);
}

// ---------------------------------------------------------------------------------------------------
// kernel 3: per-read sort by (gPos,rPos) + clustering (GenerateAlignmentCandidate)
// ---------------------------------------------------------------------------------------------------
// ChrLocMap.lower_bound(g)->first : last coordinate of the sequence that contains g
__device__ __forceinline__ int64_t chr_end_of(const DevIndex &ix, int64_t g)
{
    int lo = 0, hi = ix.n_ends;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (ix.chr_ends[mid] < g) lo = mid + 1; else hi = mid;
    }
    return ix.chr_ends[min(lo, ix.n_ends - 1)];
}

// the chaining test of AlignmentCandidates.cpp:264-265 between consecutive sorted seeds j -> k
__device__ __forceinline__ bool chains(const DevIndex &ix, int64_t gj, int rj, int64_t gk, int rk, int max_gaps, int max_intron)
{
    int64_t d = (gk - rk) - (gj - rj);
    if (d < 0) d = -d;
    if (d < max_gaps) return true;
    return d < max_intron && gk < chr_end_of(ix, gj) && rk > rj;
}

__device__ __forceinline__ int cand_threshold(int rlen) { return (int)((double)rlen * 0.3); } // AlignmentCandidates.cpp:251

// Sort + cluster one read with a group of W lanes (one seed per lane, n <= W): register bitonic sort of the keys, then the
// chaining test between neighbours, a segmented sum of the seed lengths and a compaction of the kept clusters — ballots
// and shuffles only.
template <int W>
__device__ __forceinline__ void sort_cluster_group(const DevIndex &ix, const SeedLaunch &a, int r, int64_t off, int n, int gl,
                                                   unsigned gmask, int gshift)
{
    constexpr unsigned WM = W == 32 ? 0xffffffffu : ((1u << (W & 31)) - 1u);
    uint64_t key = gl < n ? a.keys[off + gl] : ~0ull;
#pragma unroll
    for (int k = 2; k <= W; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            uint64_t other = __shfl_xor_sync(gmask, key, j);
            bool keep_min = ((gl & j) == 0) == ((gl & k) == 0);
            key = keep_min ? (key < other ? key : other) : (key > other ? key : other);
        }
    }
    const bool valid = gl < n;
    if (valid) a.keys[off + gl] = key;
    const int64_t g = key_gpos(key);
    const int rp = key_rpos(key), len = key_len(key);
    const int64_t pd = g - rp;
    const unsigned nonneg = (__ballot_sync(gmask, valid && pd >= 0) >> gshift) & WM;
    const int first = nonneg ? __ffs(nonneg) - 1 : n;
    const int64_t gprev = __shfl_up_sync(gmask, g, 1, W);
    const int rprev = __shfl_up_sync(gmask, rp, 1, W);
    const bool in = valid && gl >= first;
    const bool chain = in && gl > first && chains(ix, gprev, rprev, g, rp, a.max_gaps, a.max_intron);
    const bool head = in && !chain;
    int v = in ? len : 0;
#pragma unroll
    for (int d = 1; d < W; d <<= 1) { int t = __shfl_up_sync(gmask, v, d, W); if (gl >= d) v += t; }
    const unsigned heads = (__ballot_sync(gmask, head) >> gshift) & WM;
    const unsigned higher = gl == W - 1 ? 0u : (heads & ~((2u << gl) - 1u));
    const int nh = higher ? __ffs(higher) - 1 : n;
    const int last = min(max(nh - 1, 0), W - 1);
    const int pend = __shfl_sync(gmask, v, last, W);
    int pprev = __shfl_up_sync(gmask, v, 1, W);
    if (gl == 0) pprev = 0;
    const int score = pend - pprev;
    const bool keep = head && score > cand_threshold(a.rlen[r]);
    const unsigned km = (__ballot_sync(gmask, keep) >> gshift) & WM;
    if (keep) {
        const int ci = __popc(km & ((1u << gl) - 1u));
        a.cand_begin[off + ci] = gl;
        a.cand_count[off + ci] = nh - gl;
        a.cand_score[off + ci] = score;
    }
    if (gl == 0) a.ncand[r] = __popc(km);
}

// first pass: 8 lanes per read (most reads have 1-6 seeds); reads with more are queued for wider groups
__global__ void __launch_bounds__(256)
k_sort_cluster_small(DevIndex ix, SeedLaunch a)
{
    if (a.ctl->abort) return;
    const int lane = threadIdx.x & 31, gl = lane & 7, gshift = lane & ~7;
    const unsigned gmask = 0xFFu << gshift;
    const int ngroups = (gridDim.x * blockDim.x) >> 3;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; r < a.n_reads; r += ngroups) {
        const int64_t off = a.seed_off[r];
        const int64_t n64 = a.seed_off[r + 1] - off;
        if (n64 == 0) { if (gl == 0) a.ncand[r] = 0; continue; }
        if (n64 > 8) {
            if (gl == 0) {
                if (n64 > 32) { uint32_t i = atomicAdd(a.big_count, 1u); a.big_list[i] = (uint32_t)r; }
                else { uint32_t i = atomicAdd(a.mid_count, 1u); a.mid_list[i] = (uint32_t)r; }
            }
            continue;
        }
        sort_cluster_group<8>(ix, a, r, off, (int)n64, gl, gmask, gshift);
    }
}

// second pass: one warp per queued read with 9..32 seeds
__global__ void __launch_bounds__(256)
k_sort_cluster_warp(DevIndex ix, SeedLaunch a)
{
    if (a.ctl->abort) return;
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const uint32_t nmid = *a.mid_count;
    for (uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < nmid; w += nwarps) {
        const int r = (int)a.mid_list[w];
        const int64_t off = a.seed_off[r];
        sort_cluster_group<32>(ix, a, r, off, (int)(a.seed_off[r + 1] - off), lane, FULL, 0);
    }
}

constexpr int BIG_SM_CAP = 4096; // keys sorted in shared memory up to this many (padded to a power of two)

__global__ void __launch_bounds__(256)
k_sort_cluster_big(DevIndex ix, SeedLaunch a)
{
    if (a.ctl->abort) return;
    __shared__ uint64_t sm[BIG_SM_CAP];
    const int tid = threadIdx.x;
    const uint32_t nbig = *a.big_count;
    for (uint32_t b = blockIdx.x; b < nbig; b += gridDim.x) {
        int r = (int)a.big_list[b];
        int64_t off = a.seed_off[r];
        int64_t n = a.seed_off[r + 1] - off;
        int64_t N = 64;
        while (N < n) N <<= 1;
        uint64_t *buf = sm;
        if (N > BIG_SM_CAP) {
            if ((size_t)N > a.big_scratch_per_cta) { if (tid == 0) { a.ncand[r] = 0; atomicOr(&a.ctl->err, ERR_SORT_SCRATCH); } continue; } // cannot happen: bound = cap_rec*max_dup
            buf = a.big_scratch + (size_t)blockIdx.x * a.big_scratch_per_cta;
        }
        for (int64_t i = tid; i < N; i += blockDim.x) buf[i] = i < n ? a.keys[off + i] : ~0ull;
        __syncthreads();
        for (int64_t k = 2; k <= N; k <<= 1)
            for (int64_t j = k >> 1; j > 0; j >>= 1) {
                for (int64_t i = tid; i < N; i += blockDim.x) {
                    int64_t p = i ^ j;
                    if (p > i) {
                        uint64_t x = buf[i], y = buf[p];
                        bool up = (i & k) == 0;
                        if ((x > y) == up) { buf[i] = y; buf[p] = x; }
                    }
                }
                __syncthreads();
            }
        for (int64_t i = tid; i < n; i += blockDim.x) a.keys[off + i] = buf[i];
        __syncthreads();
        if (tid == 0) { // the greedy scan itself, sequential: big reads are rare outside config 5
            int thr = cand_threshold(a.rlen[r]);
            int64_t i = 0;
            uint32_t nc = 0;
            while (i < n && key_gpos(buf[i]) - key_rpos(buf[i]) < 0) i++;
            while (i < n) {
                int score = key_len(buf[i]);
                int64_t k = i + 1;
                for (; k < n; k++) {
                    if (!chains(ix, key_gpos(buf[k - 1]), key_rpos(buf[k - 1]), key_gpos(buf[k]), key_rpos(buf[k]),
                                a.max_gaps, a.max_intron)) break;
                    score += key_len(buf[k]);
                }
                if (score > thr) {
                    a.cand_begin[off + nc] = (int32_t)i;
                    a.cand_count[off + nc] = (int32_t)(k - i);
                    a.cand_score[off + nc] = score;
                    nc++;
                }
                i = k;
            }
            a.ncand[r] = nc;
        }
        __syncthreads();
    }
}

void launch_sort_cluster(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st)
{
    if (a.n_reads <= 0) return;
    const int sms = sm_count();
    int64_t want = ((int64_t)a.n_reads * 8 + 255) / 256;
    int grid = (int)(want < sms * 8 ? want : sms * 8);
    k_sort_cluster_small<<<grid, 256, 0, st>>>(ix, a);
    k_sort_cluster_warp<<<sms * 4, 256, 0, st>>>(ix, a);
    k_sort_cluster_big<<<sms, 256, 0, st>>>(ix, a);
}

} // namespace dartgpu
