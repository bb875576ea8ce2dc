// Seeding on the GPU: FM-index exact-match segment search, SA locate, per-read seed sort and
// simple-pair clustering into alignment candidates.
//
// Replaces (bit-exactly) /root/reference/src/bwt_search.cpp:43-182 (bwt_occ4 / bwt_2occ4 / bwt_occ /
// bwt_invPsi / bwt_sa / BWT_Search) and /root/reference/src/AlignmentCandidates.cpp:181-288
// (IdentifySeedPairs, GenerateAlignmentCandidate).
//
// Execution model
//   * a GROUP of 4 lanes owns one read at a time; 8 groups per warp run 8 independent FM chains.
//     One rank query = one 128-bit load per lane = one 64-byte Occ block per group, each lane
//     popcounting its own 32 symbols (2 POPC: "equal to c" and "greater than c"), combined with xor-shuffles.
//   * the nested loops of the reference (reads > search starts > extension steps) are flattened into a
//     single loop whose every iteration performs exactly one rank step, so the 8 groups of a warp stay
//     converged although their reads, search starts and match lengths differ.
//   * reads are staged in shared memory 2-bit packed (+ a 1-bit "not ACGT" plane).
//   * search only records SA intervals (of the reverse-complemented pattern, see k_search); after a prefix sum over
//     the hit counts a second kernel resolves every hit with its own group (LF walk to the next SA sample, then the
//     mirror p -> 2G - p - len), so repetitive reads do not serialise.
//   * seeds are 64-bit keys (gPos | rPos | len): sorting the keys is the reference's CompByGenomePos order.
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include <algorithm>

#include "dartgpu_internal.h"
#include "rank.cuh"

namespace dartgpu {

constexpr int SEARCH_THREADS = 128;
constexpr int GROUPS_PER_CTA = SEARCH_THREADS / 4;
constexpr int RWORDS = DARTGPU_MAX_RLEN / 16;
constexpr unsigned FULL = 0xffffffffu;

// ---------------------------------------------------------------------------------------------------
// rank primitives (rank.cuh holds the per-quarter arithmetic, unit-tested on the host)
// ---------------------------------------------------------------------------------------------------
// stage one read in the group's shared-memory slot: 2 bits per base + 1 ambiguity bit per base
__device__ __forceinline__ void stage_read(const uint8_t *codes, int64_t off, int rl, uint32_t *pk, uint16_t *am,
                                           int q, unsigned gmask)
{
    int nw = (rl + 15) >> 4;
    __syncwarp(gmask);
    for (int w = q; w < nw; w += 4) {
        uint4 v = *reinterpret_cast<const uint4 *>(codes + off + 16 * (int64_t)w);
        uint32_t x[4] = {v.x, v.y, v.z, v.w};
        uint32_t p2 = 0, a = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t t = x[i];
            uint32_t two = (t & 3u) | ((t >> 6) & 0xCu) | ((t >> 12) & 0x30u) | ((t >> 18) & 0xC0u);
            uint32_t ab = ((t >> 2) & 1u) | ((t >> 9) & 2u) | ((t >> 16) & 4u) | ((t >> 23) & 8u);
            p2 |= two << (8 * i);
            a |= ab << (4 * i);
        }
        pk[w] = p2;
        am[w] = (uint16_t)a;
    }
    __syncwarp(gmask);
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
// i.e. one "equal to c" popcount per block and no forward-interval bookkeeping at all (round-1 ncu: the kernel is bound
// by the integer pipes, not by memory; dropping x0 removes ~30 % of the step's instructions).
// Occ(c,.) = block header count (lane c holds it) + in-block count; the in-block parts of the 4 lanes are summed as two
// 8-bit fields of one word (two xor-shuffles), the header difference travels as a 32-bit value (interval widths are
// < 2^32, checked at index load).  IdxT = uint32_t when the whole text fits 32 bits (every tested genome and BASELINE
// configs 1, 2, 5), uint64_t otherwise (3.1 Gbp: 2G = 6.2e9) — same code, narrower interval arithmetic.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stage_tables(const DevIndex &ix, uint64_t *s_L2, uint32_t *s_mask)
{
    if (threadIdx.x == 0) { s_L2[0] = ix.L2[0]; s_L2[1] = ix.L2[1]; s_L2[2] = ix.L2[2]; s_L2[3] = ix.L2[3]; s_L2[4] = ix.L2[4]; }
    if (threadIdx.x < 33) s_mask[threadIdx.x] = prefix_mask32((int)threadIdx.x);
    __syncthreads();
}

template <typename IdxT>
__global__ void __launch_bounds__(SEARCH_THREADS, 16)
k_search(DevIndex ix, SeedLaunch a)
{
    __shared__ uint32_t s_pk[GROUPS_PER_CTA][RWORDS];
    __shared__ uint16_t s_am[GROUPS_PER_CTA][RWORDS];
    __shared__ uint64_t s_L2[5];
    __shared__ uint32_t s_mask[33];

    const int lane = threadIdx.x & 31, q = lane & 3;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int grp = threadIdx.x >> 2;
    stage_tables(ix, s_L2, s_mask);
    uint32_t *pk = s_pk[grp];
    uint16_t *am = s_am[grp];
    const IdxT primary = (IdxT)ix.primary;
    const char *occq = reinterpret_cast<const char *>(ix.occ + q);   // this lane's quarter of block 0
    const int q32 = 32 * q - 1;
    // Reads are handed out dynamically: the CTA owns a contiguous range and its 32 groups draw the next read from a
    // shared counter when they finish one (round-1 ncu: with a static stride a quarter of the lanes idled in the step
    // waiting for the warp's slowest group).  Results are indexed by read, so the order does not matter.
    __shared__ int s_next;
    const int per_cta = (a.n_reads + gridDim.x - 1) / gridDim.x;
    const int r_end = min(a.n_reads, (int)(blockIdx.x + 1) * per_cta);
    if (threadIdx.x == 0) s_next = blockIdx.x * per_cta;
    __syncthreads();
    int r = 0;

    bool have_read = false, searching = false;
    int rl = 0, start = 0, p = 0, cw = -1;
    uint32_t nr = 0, nh = 0, x2 = 0, wcode = 0, wamb = 0;
    IdxT x1 = 0;
    uint32_t st_steps = 0, st_splits = 0;

    for (;;) {
        if (!searching) {
            bool done = false;
            for (;;) {
                if (!have_read) {
                    if (q == 0) r = atomicAdd(&s_next, 1);
                    r = __shfl_sync(gmask, r, 0, 4);
                    if (r >= r_end) { done = true; break; }
                    rl = a.rlen[r];
                    stage_read(a.codes, a.dev_off[r], rl, pk, am, q, gmask);
                    start = 0; nr = 0; nh = 0; have_read = true; cw = -1;
                }
                while (start < rl - 13 && ((am[start >> 4] >> (start & 15)) & 1)) start++;
                if (start < rl - 13) break;
                if (q == 0) { a.nrec[r] = nr; a.nhits[r] = nh; }
                have_read = false;
            }
            if (done) break;
            int c0 = (pk[start >> 4] >> ((start & 15) * 2)) & 3;
            x1 = (IdxT)s_L2[3 - c0] + 1; x2 = (uint32_t)(s_L2[c0 + 1] - s_L2[c0]);
            p = start + 1;
            searching = true;
        }
        bool end = p >= rl;
        if (!end) {
            if ((p >> 4) != cw) { cw = p >> 4; wcode = pk[cw]; wamb = am[cw]; }
            end = (wamb >> (p & 15)) & 1;
        }
        if (!end) {
            const IdxT k = x1 - 1, l = k + x2;
            const IdxT kk = k - (k >= primary), ll = l - (l >= primary);
            // quarter q of block (kk >> 7): byte offset (kk >> 7) * 64 = (kk & ~127) >> 1
            const ulonglong2 vk = __ldg(reinterpret_cast<const ulonglong2 *>(occq + ((uint64_t)(kk & ~(IdxT)127) >> 1)));
            const ulonglong2 vl = __ldg(reinterpret_cast<const ulonglong2 *>(occq + ((uint64_t)(ll & ~(IdxT)127) >> 1)));
            const uint32_t code = (wcode >> ((p & 15) * 2)) & 3u;      // c = 3 - code: complement
            const int c = 3 - (int)code;
            const uint32_t CH = (code & 2u) ? 0u : ~0u, CL = (code & 1u) ? 0u : ~0u;
            st_steps++; st_splits += (uint32_t)((kk ^ ll) >> 7 != 0);
            uint32_t lo, hi;
            planes32(vk.y, lo, hi);
            const int eqk = count_eq32(lo, hi, s_mask[max(0, min(32, (int)((uint32_t)kk & 127u) - q32))], CH, CL);
            planes32(vl.y, lo, hi);
            const int eql = count_eq32(lo, hi, s_mask[max(0, min(32, (int)((uint32_t)ll & 127u) - q32))], CH, CL);
            uint32_t pc = (uint32_t)eqk | (uint32_t)eql << 16;
            pc += __shfl_xor_sync(gmask, pc, 1);
            pc += __shfl_xor_sync(gmask, pc, 2);
            const uint32_t Dc = __shfl_sync(gmask, (uint32_t)vl.x - (uint32_t)vk.x, c, 4);
            const IdxT cntk = __shfl_sync(gmask, (IdxT)vk.x, c, 4);
            const uint32_t EK = pc & 0xffffu, EL = pc >> 16;
            const uint32_t n2 = Dc + EL - EK;
            if (n2 == 0) end = true;
            else {
                x1 = (IdxT)s_L2[c] + 1 + cntk + EK;
                x2 = n2;
                p++;
            }
        }
        if (end) {
            int len = p - start;
            if (x2 <= a.max_dup && len >= 16) { // bwt_search.cpp:173
                if (q == 0 && (int)nr < a.cap_rec) {
                    SearchRec rec; rec.sa_begin = x1; rec.freq = x2; rec.start = (uint16_t)start; rec.len = (uint16_t)len;
                    a.recs[(int64_t)r * a.cap_rec + nr] = rec;
                }
                nr++; nh += x2;
                start += len;
            } else start++;
            searching = false;
        }
    }
    if (q == 0 && st_steps) {
        atomicAdd(&a.stats->ext_steps, (unsigned long long)st_steps);
        atomicAdd(&a.stats->ext_blocks, (unsigned long long)st_steps + st_splits);
    }
}

static bool fits32(const DevIndex &ix) { return !ix.force64 && ix.seq_len + 2 < (1ull << 32); }

void launch_search(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st)
{
    if (a.n_reads <= 0) return;
    // exactly one wave: as many CTAs as are resident at once (a partial second wave would idle most SMs at the end)
    static int occ32 = 0, occ64 = 0, sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ32, k_search<uint32_t>, SEARCH_THREADS, 0);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ64, k_search<uint64_t>, SEARCH_THREADS, 0);
    }
    const bool narrow = fits32(ix);
    int want = (a.n_reads + GROUPS_PER_CTA - 1) / GROUPS_PER_CTA;
    int grid = sms * std::max(1, narrow ? occ32 : occ64);
    if (want < grid) grid = want;
    if (narrow) k_search<uint32_t><<<grid, SEARCH_THREADS, 0, st>>>(ix, a);
    else k_search<uint64_t><<<grid, SEARCH_THREADS, 0, st>>>(ix, a);
}

// ---------------------------------------------------------------------------------------------------
// prefix sums (plumbing): seed_off = exclusive scan of nhits over n_reads+1 entries (nhits[n_reads] == 0)
// ---------------------------------------------------------------------------------------------------
struct U32ToI64 { __host__ __device__ int64_t operator()(const uint32_t &v) const { return (int64_t)v; } };

size_t scan_tmp_bytes(int n)
{
    size_t bytes = 0;
    cub::TransformInputIterator<int64_t, U32ToI64, const uint32_t *> it((const uint32_t *)nullptr, U32ToI64());
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, (int64_t *)nullptr, n + 1);
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
// kernel 2b: SA locate — bwt_sa / bwt_invPsi (bwt_search.cpp:119-137), one group per hit
// ---------------------------------------------------------------------------------------------------
// Hits need 0..sa_intv-1 LF steps each (uniformly spread), so a group that walked one hit per pass would idle half the
// time waiting for the slowest of the warp's 8 groups (round-1 ncu: 12 of 32 lanes active).  Same cure as in k_search:
// one loop, one LF step per iteration, a group picks up its next hit the moment it finishes one.
template <typename IdxT>
__global__ void __launch_bounds__(SEARCH_THREADS)
k_locate(DevIndex ix, SeedLaunch a, int64_t total)
{
    __shared__ uint64_t s_L2[5];
    __shared__ uint32_t s_mask[33];
    stage_tables(ix, s_L2, s_mask);
    const int lane = threadIdx.x & 31, q = lane & 3;
    const unsigned gmask = 0xFu << (lane & ~3);
    const int64_t ngroups = (int64_t)gridDim.x * GROUPS_PER_CTA;
    const IdxT primary = (IdxT)ix.primary, sa_mask = (IdxT)ix.sa_mask;
    const char *occq = reinterpret_cast<const char *>(ix.occ + q);
    const int q32 = 32 * q - 1;
    unsigned long long st_lf = 0, st_hits = 0;
    int64_t s = (int64_t)blockIdx.x * GROUPS_PER_CTA + (threadIdx.x >> 2);
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
        if (k & sa_mask) {                  // one LF step = one block: the symbol at k and its rank come from the same 64 bytes
            steps++;
            if (k == primary) k = 0;
            else {
                const IdxT kk = k - (k > primary);
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2 *>(occq + ((uint64_t)(kk & ~(IdxT)127) >> 1)));
                const int o = (int)((uint32_t)kk & 127u);
                const int c = __shfl_sync(gmask, symbol_at(v.y, o & 31), o >> 5, 4);
                const uint32_t CH = (c & 2) ? ~0u : 0u, CL = (c & 1) ? ~0u : 0u;
                uint32_t lo, hi;
                planes32(v.y, lo, hi);
                uint32_t eq = (uint32_t)count_eq32(lo, hi, s_mask[max(0, min(32, o - q32))], CH, CL);
                eq += __shfl_xor_sync(gmask, eq, 1);
                eq += __shfl_xor_sync(gmask, eq, 2);
                const IdxT cnt = __shfl_sync(gmask, (IdxT)v.x, c, 4);
                k = (IdxT)s_L2[c] + cnt + eq;
            }
        } else {
            // position of revcomp(P) on the text; P itself starts at the mirrored coordinate 2G - p' - |P|
            const uint64_t prc = (uint64_t)steps + __ldg(ix.sa + ((uint64_t)k >> ix.sa_shift));
            if (q == 0) {
                const uint32_t m = a.meta[s];
                a.keys[s] = seed_key(2 * (uint64_t)ix.G - prc - (m & 0xFFFF), m >> 16, m & 0xFFFF);
            }
            st_lf += steps; st_hits++;
            s += ngroups;
            walking = false;
        }
    }
    if (q == 0 && st_hits) {
        atomicAdd(&a.stats->lf_steps, st_lf);
        atomicAdd(&a.stats->hits, st_hits);
        atomicAdd(&a.stats->seeds, st_hits);
    }
}

void launch_expand_locate(const DevIndex &ix, const SeedLaunch &a, int64_t total, cudaStream_t st)
{
    if (a.n_reads <= 0 || total <= 0) return;
    int grid = (a.n_reads + 255) / 256;
    if (grid > 148 * 8) grid = 148 * 8;
    k_expand<<<grid, 256, 0, st>>>(a);
    int64_t want = (total + GROUPS_PER_CTA - 1) / GROUPS_PER_CTA;
    int g2 = (int)(want < 148 * 16 ? want : 148 * 16);
    if (fits32(ix)) k_locate<uint32_t><<<g2, SEARCH_THREADS, 0, st>>>(ix, a, total);
    else k_locate<uint64_t><<<g2, SEARCH_THREADS, 0, st>>>(ix, a, total);
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
            if ((size_t)N > a.big_scratch_per_cta) { if (tid == 0) a.ncand[r] = 0xFFFFFFFFu; continue; } // cannot happen: bound = cap_rec*max_dup
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
    cudaMemsetAsync(a.big_count, 0, sizeof(uint32_t), st);
    cudaMemsetAsync(a.mid_count, 0, sizeof(uint32_t), st);
    int64_t want = ((int64_t)a.n_reads * 8 + 255) / 256;
    int grid = (int)(want < 148 * 8 ? want : 148 * 8);
    k_sort_cluster_small<<<grid, 256, 0, st>>>(ix, a);
    k_sort_cluster_warp<<<148 * 4, 256, 0, st>>>(ix, a);
    k_sort_cluster_big<<<148, 256, 0, st>>>(ix, a);
}

} // namespace dartgpu
