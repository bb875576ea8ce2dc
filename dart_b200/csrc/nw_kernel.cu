// Batched Needleman-Wunsch gap fill — replaces nw_alignment() (/root/reference/src/nw_alignment.cpp:18-82)
// bit-exactly, including its quirks:
//   * the reference's 3-way max resolves to `double max(short,short,short)`, so every S cell is truncated
//     toward zero to a whole unit while R (gap in the read string) and T (gap in the genome string) keep
//     half units.  All values are multiples of 0.5, so the recurrence is restated in integer half-units:
//         R = max(R_left - 1, S_left - 3)   T = max(T_up - 1, S_up - 3)
//         h = max(S_diag +/- 3, R, T)       S = (h / 2) * 2   (C division, truncating)
//     borders S[i][0] = T[i][0] = -2 - i, R[i][0] = -131072 (and symmetrically for row 0), S[0][0] = 0;
//   * full matrix, no band (a band could change the traceback);
//   * traceback priority S==R (gap in read, consumes genome), then S==T, else diagonal.
//
// Two kernels, jobs binned by shape:
//   k_nw_thread  jobs up to 64 x 64 (the bulk of every config): ONE THREAD per alignment, see below;
//   k_nw         everything larger: one warp per job.  The matrix is swept in strips of 128 rows; lane l owns 4 consecutive
//                rows and at step t computes column t-l+1 of all four (anti-diagonal wavefront over the lanes).  Left
//                neighbours stay in registers, the row above a lane's first row arrives by shuffle from lane l-1, the genome
//                base is handed down the lanes systolically: 3 shuffles per 4 cells.  Two traceback bits per cell go to
//                global scratch, 16 cells per word, and are walked back by the whole warp (32 rows of flag words in
//                registers, handed round by shuffle).  Integer pipes only.
#include <cub/block/block_scan.cuh>

#include "dartgpu_internal.h"

namespace dartgpu {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int NW_THREADS = 128;
constexpr int NW_NEG = -131072;
constexpr int NW_ROWS = 4;                 // rows per lane in the warp-per-job kernel: strips of 128 rows

__device__ __forceinline__ int ref_base(const DevIndex &ix, int64_t p)
{
    if (p < 0 || p >= 2 * ix.G) return 0;
    return (int)((__ldg(ix.ref2 + (p >> 4)) >> (30 - 2 * (int)(p & 15))) & 3);
}

__device__ __forceinline__ bool nw_thread_class(int m, int n);

__global__ void __launch_bounds__(NW_THREADS)
k_nw(DevIndex ix, const uint8_t *__restrict__ codes, const NwJobDev *__restrict__ jobs, const uint32_t *__restrict__ order,
     const BatchCtl *__restrict__ ctl, int round, int cap_jobs,
     uint32_t *flags, int32_t *rowbuf, size_t rowbuf_per_warp, uint8_t *ops, int32_t *nops)
{
    if (ctl->abort) return;
    const int n_jobs = min(ctl->nw_jobs[round], cap_jobs);
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int32_t *rb = rowbuf + (size_t)warp * rowbuf_per_warp;
    // the jobs k_nw_thread does not take are the tail of the shape-sorted list (the last bin)
    const int first = ctl->nw_small[round];

    for (int k = first + warp; k < n_jobs; k += nwarps) {
        const int job = (int)order[k];
        const NwJobDev J = jobs[job];
        const int m = J.m, n = J.n;
        if (m <= 0 || n <= 0) {                       // one empty side: all gaps (nw_alignment.cpp:61-74 with i or j at 0)
            const int cnt = max(m, 0) + max(n, 0);
            for (int t = lane; t < cnt; t += 32) ops[J.op_off + t] = (uint8_t)(m <= 0 ? 1 : 2);
            if (lane == 0) nops[job] = cnt;
            continue;
        }
        if (m > 32 * NW_ROWS && (size_t)2 * (n + 1) > rowbuf_per_warp) {   // cannot happen: the host sizes the row buffer from the bound on n
            if (lane == 0) { nops[job] = 0; atomicOr(const_cast<int32_t *>(&ctl->err), ERR_NW_WIDTH); }
            continue;
        }
        const int wpr = (n + 15) >> 4;
        uint32_t *fl = flags + J.flag_off;

        // strips of NW_ROWS x 32 rows: lane l owns NW_ROWS consecutive rows and, at step t, column t - l + 1 of all of them
        for (int s0 = 0; s0 < m; s0 += 32 * NW_ROWS) {
            const int i0 = s0 + NW_ROWS * lane + 1;               // my first row (1-based)
            int a[NW_ROWS], Sl[NW_ROWS], Rl[NW_ROWS];
            uint32_t fw[NW_ROWS];
#pragma unroll
            for (int r = 0; r < NW_ROWS; r++) {
                const int i = i0 + r;
                int c = i <= m ? (int)codes[J.s1_off + i - 1] : 7;
                a[r] = (c & 4) ? 7 : (c & 3);              // 8..11 are lower-case ACGT: same base for the table compare
                Sl[r] = -2 - i; Rl[r] = NW_NEG; fw[r] = 0;  // S[i][0], R[i][0]
            }
            int Sd0 = (i0 == 1) ? 0 : -2 - (i0 - 1);       // S[i0-1][0]
            int So = 0, To = 0, bo = 0;                    // what I hand to the lane below: my last row
            const int rows = min(32 * NW_ROWS, m - s0);
            const int lanes_used = (rows + NW_ROWS - 1) / NW_ROWS;
            const int steps = n + lanes_used - 1;
            const bool last_strip = s0 + 32 * NW_ROWS >= m;
            int bch = 0, such = 0, tuch = 0;
            for (int t = 0; t < steps; t++) {
                if ((t & 31) == 0) {                  // fetch the next 32 columns of lane 0's inputs
                    int jc = t + lane + 1;
                    bch = jc <= n ? ref_base(ix, J.gpos + jc - 1) : 0;
                    if (s0 > 0 && jc <= n) { such = __ldcg(rb + jc); tuch = __ldcg(rb + (n + 1) + jc); }
                }
                int Su = __shfl_up_sync(FULLM, So, 1);
                int Tu = __shfl_up_sync(FULLM, To, 1);
                int b = __shfl_up_sync(FULLM, bo, 1);
                const int b0 = __shfl_sync(FULLM, bch, t & 31);
                const int j = t - lane + 1;
                if (s0 > 0) {
                    const int su0 = __shfl_sync(FULLM, such, t & 31), tu0 = __shfl_sync(FULLM, tuch, t & 31);
                    if (lane == 0) { Su = su0; Tu = tu0; }
                } else if (lane == 0) { Su = -2 - j; Tu = NW_NEG; }
                if (lane == 0) b = b0;
                if (i0 <= m && j >= 1 && j <= n) {
                    int upS = Su, upT = Tu, diag = Sd0;
                    const bool flush = ((j - 1) & 15) == 15 || j == n;
                    const int fsh = ((j - 1) & 15) * 2;
#pragma unroll
                    for (int r = 0; r < NW_ROWS; r++) {
                        if (i0 + r <= m) {
                            const int R = max(Rl[r] - 1, Sl[r] - 3);
                            const int T = max(upT - 1, upS - 3);
                            const int h = max(diag + (a[r] == b ? 3 : -3), max(R, T));
                            const int S = (h / 2) * 2;
                            fw[r] |= ((S == R ? 1u : 0u) | (S == T ? 2u : 0u)) << fsh;
                            if (flush) { __stcg(fl + (size_t)(i0 + r - 1) * wpr + ((j - 1) >> 4), fw[r]); fw[r] = 0; }
                            diag = Sl[r];                  // S[i][j-1] is the diagonal of the row below
                            Sl[r] = S; Rl[r] = R;
                            upS = S; upT = T;
                        }
                    }
                    Sd0 = Su;
                    So = upS; To = upT; bo = b;            // my last valid row (the lane below only has rows if all of mine are valid)
                    if (lane == 31 && !last_strip) { __stcg(rb + j, upS); __stcg(rb + (n + 1) + j, upT); }
                }
            }
            __syncwarp();
        }
        __syncwarp();
        {
            // Traceback, warp-uniform: the 32 lanes hold the flag words of 32 consecutive rows (ending at the current row) for
            // the current 16-column word and hand them round by shuffle; one reload every >= 16 steps instead of one
            // dependent L2 round trip per step (round-1 launch list of the 2x250 config: a 250 x 500 job spent ~3x longer
            // walking back through global memory on lane 0 than sweeping).
            int ti = m, tj = n, cnt = 0;
            int64_t pos = J.op_off + m + n;
            int base_i = -1, wcol = -1;
            uint32_t myw = 0;
            while (ti > 0 || tj > 0) {
                int op;
                if (ti == 0) op = 1;
                else if (tj == 0) op = 2;
                else {
                    const int wc = (tj - 1) >> 4;
                    if (wc != wcol || base_i - ti >= 32 || base_i < ti) {
                        base_i = ti; wcol = wc;
                        const int row = base_i - 1 - lane;
                        myw = row >= 0 ? __ldcg(fl + (size_t)row * wpr + wcol) : 0u;
                    }
                    const uint32_t w = __shfl_sync(FULLM, myw, base_i - ti);
                    const uint32_t f = (w >> (((tj - 1) & 15) * 2)) & 3u;
                    op = (f & 1u) ? 1 : ((f & 2u) ? 2 : 0);
                }
                --pos;
                if (lane == 0) ops[pos] = (uint8_t)op;
                cnt++;
                if (op == 1) tj--; else if (op == 2) ti--; else { ti--; tj--; }
            }
            if (lane == 0) nops[job] = cnt;
        }
        __syncwarp();
    }
}

// Jobs up to 64 x 64 cells — the bulk of every config: a mismatch or a short indel between two seeds, a gap of a few
// dozen bases against its flank — are aligned by ONE THREAD each, 32 independent alignments per warp: row-wise sweep, the
// previous row (S, T as int16: scores stay within +-2000 half-units, "minus infinity" is -30000) in shared memory laid out
// [column][thread] (conflict-free), the genome bases of the job in two 64-bit registers, the 2-bit traceback flags of
// a row written as ceil(n/16) words to the job's slice of the global flag scratch and walked back by the same thread.
// Round-1 ncu of the warp-per-job kernel: issue-bound (66 %), 21 of 32 lanes active, ~130 thread-instructions per cell
// (pipeline fill/drain of a 32-lane anti-diagonal sweep over a 10..1300-cell matrix, 4 shuffles per step, warp-uniform
// traceback); a thread needs ~25 per cell.  Jobs are sorted by (n, m) first so that the alignments of a warp have the
// same shape.
constexpr int NWT_THREADS = 128, NWT_MAX = 64;
constexpr int NWT_NEG = -30000;
__device__ __forceinline__ bool nw_thread_class(int m, int n) { return m >= 1 && n >= 1 && m <= NWT_MAX && n <= NWT_MAX; }

// ---- shape-class counting sort of the job queue (replaces round 1's cub radix sort + 3 scans + 4 host read-backs) ----
// bin = (n-1)*64 + (m-1) for the thread-per-alignment class, NW_BINS-1 for everything else.  k_nw_prepare also hands every
// job its slices of the column / traceback-flag / Rvec-Lvec pools (one atomic per CTA and pool: slice order is arbitrary and
// nothing depends on it) and accumulates the work counters.
__device__ __forceinline__ int nw_bin(int m, int n) { return nw_thread_class(m, n) ? (n - 1) * NWT_MAX + (m - 1) : NW_BINS - 1; }

__global__ void __launch_bounds__(256)
k_nw_prepare(NwJobDev *__restrict__ jobs, int cap_jobs, int round, int with_aux, uint32_t *__restrict__ hist, BatchCtl *ctl)
{
    if (ctl->abort) return;
    typedef cub::BlockScan<unsigned int, 256> Scan;
    __shared__ typename Scan::TempStorage tmp;
    __shared__ unsigned long long s_base[3];
    const int n_jobs = min(ctl->nw_jobs[round], cap_jobs);
    unsigned long long cells = 0;
    int mx = 0;
    for (int j0 = blockIdx.x * 256; j0 < n_jobs; j0 += gridDim.x * 256) {
        const int j = j0 + threadIdx.x;
        unsigned int so = 0, sf = 0, sa = 0;
        int m = 0, n = 0, bin = -1 - (int)(threadIdx.x & 31);
        if (j < n_jobs) {
            m = jobs[j].m; n = jobs[j].n;
            so = (unsigned)(max(m, 0) + max(n, 0));
            sf = (m > 0 && n > 0 && !nw_thread_class(m, n)) ? (unsigned)(m * ((n + 15) >> 4)) : 0u;   // the thread class keeps its flags on chip
            sa = (with_aux && !(j & 1)) ? (unsigned)(2 * (max(m, 0) + 1)) : 0u;
            bin = nw_bin(m, n);
            if (m > 0 && n > 0) cells += (unsigned long long)m * n;
            mx = max(mx, n);
        }
        {   // histogram: the lanes of a warp that hold the same shape share one atomic (a million single-address atomics otherwise)
            const unsigned peers = __match_any_sync(FULLM, bin);
            if (bin >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (unsigned)__popc(peers));
        }
        unsigned int po, pf, pa, to, tf, ta;
        Scan(tmp).ExclusiveSum(so, po, to); __syncthreads();
        Scan(tmp).ExclusiveSum(sf, pf, tf); __syncthreads();
        Scan(tmp).ExclusiveSum(sa, pa, ta); __syncthreads();
        if (threadIdx.x == 0) {
            s_base[0] = atomicAdd(&ctl->nw_ops[round], (unsigned long long)to);
            s_base[1] = atomicAdd(&ctl->nw_flags[round], (unsigned long long)tf);
            s_base[2] = ta ? atomicAdd(&ctl->nw_aux[round], (unsigned long long)ta) : 0ull;
        }
        __syncthreads();
        if (j < n_jobs) { jobs[j].op_off = (int64_t)(s_base[0] + po); jobs[j].flag_off = (int64_t)(s_base[1] + pf); jobs[j].aux_off = (int64_t)(s_base[2] + pa); }
        __syncthreads();
    }
    mx = __reduce_max_sync(FULLM, mx);
    for (int d = 16; d > 0; d >>= 1) cells += __shfl_xor_sync(FULLM, cells, d);
    if ((threadIdx.x & 31) == 0) {
        if (mx > 0) atomicMax(&ctl->nw_max_n[round], mx);
        if (cells) atomicAdd(&ctl->work[0], cells);
    }
}

// one CTA: exclusive scan of the histogram; the capacity check of the three pools
__global__ void __launch_bounds__(1024)
k_nw_bins(uint32_t *__restrict__ hist, uint32_t *__restrict__ bin_start, uint32_t *__restrict__ bin_cur, uint32_t *__restrict__ next_chunk, int round,
          long long cap_ops, long long cap_flags, long long cap_aux, BatchCtl *ctl)
{
    if (threadIdx.x == 0) *next_chunk = 0;              // k_nw_thread's hand-out counter
    // (no early return on abort: the histogram must be left zeroed for the next launch)
    typedef cub::BlockScan<unsigned int, 1024> Scan;
    __shared__ typename Scan::TempStorage tmp;
    constexpr int PER = (NW_BINS + 1023) / 1024;        // 5 consecutive bins per thread
    unsigned int v[PER], sum = 0;
#pragma unroll
    for (int i = 0; i < PER; i++) { const int b = threadIdx.x * PER + i; v[i] = b < NW_BINS ? hist[b] : 0u; sum += v[i]; if (b < NW_BINS) hist[b] = 0u; }
    unsigned int pre;
    Scan(tmp).ExclusiveSum(sum, pre);
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int b = threadIdx.x * PER + i;
        if (b < NW_BINS) { bin_start[b] = pre; bin_cur[b] = pre; if (b == NW_BINS - 1) ctl->nw_small[round] = (int32_t)pre; }
        pre += v[i];
    }
    if (threadIdx.x == 0 && ((long long)ctl->nw_ops[round] > cap_ops || (long long)ctl->nw_flags[round] > cap_flags || (long long)ctl->nw_aux[round] > cap_aux))
        atomicOr(&ctl->abort, round == 0 ? CAP_NW_B : CAP_NW_C);
}

// order[] = job ids grouped by bin; the lanes of a warp that hold jobs of the same bin share one atomic
__global__ void __launch_bounds__(256)
k_nw_scatter(const NwJobDev *__restrict__ jobs, int cap_jobs, int round, uint32_t *__restrict__ bin_cur, uint32_t *__restrict__ order,
             NwSorted *__restrict__ sorted, const BatchCtl *__restrict__ ctl)
{
    if (ctl->abort) return;
    const int n_jobs = min(ctl->nw_jobs[round], cap_jobs);
    const int lane = threadIdx.x & 31;
    for (int j0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31; j0 < n_jobs; j0 += gridDim.x * blockDim.x) {
        const int j = j0 + lane;
        const bool valid = j < n_jobs;
        const int bin = valid ? nw_bin(jobs[j].m, jobs[j].n) : -1 - lane;
        const unsigned peers = __match_any_sync(FULLM, bin);
        const int leader = __ffs(peers) - 1;
        unsigned base = 0;
        if (valid && lane == leader) base = atomicAdd(&bin_cur[bin], (unsigned)__popc(peers));
        base = __shfl_sync(FULLM, base, leader);
        if (valid) {
            const unsigned at = base + __popc(peers & ((1u << lane) - 1u));
            order[at] = (uint32_t)j;
            const NwJobDev J = jobs[j];
            NwSorted d; d.s1_off = J.s1_off; d.gpos = J.gpos; d.op_off = J.op_off; d.job = j; d.m = (int16_t)min(J.m, 32767); d.n = (int16_t)min(J.n, 32767);
            sorted[at] = d;           // the thread-per-alignment kernel reads its 32 descriptors as one coalesced kilobyte
        }
    }
}

// ---- the thread-per-alignment kernel, second version ----
// Round-1 ncu of the first version: warps active 21-26 % (33 KB of static shared rows per CTA), ~45 SASS instructions per
// cell, every traceback flag word stored to and re-loaded from the job's own slice of global memory (32 sectors per warp
// store), n calls of a bounds-checked single-base load for the genome window.  This version:
//   * a warp owns a 2048-word slice of shared memory and gives each of its alignments what its shape needs — one word per
//     column for the previous row (S in the low half, the NEXT row's T in the high half, both int16) plus m x ceil(n/16)
//     words of traceback flags — laid out [word][lane]; shapes that need more than 64 words per alignment run fewer lanes
//     per pass (a 64 x 64 alignment needs 321 words: 6 lanes), the small shapes that make up almost every batch run all 32;
//   * no traceback flag touches global memory; the walk back is a chain of shared-memory loads;
//   * the recurrence is DPX (max(a + b, c) and three-way max are one instruction each on sm_100), four columns per loop
//     iteration with immediate flag masks; the row's match mask comes from two bit-plane words of the genome window
//     (built once per alignment from three aligned loads), not from a per-cell compare of extracted bases;
//   * job descriptors arrive sorted and compact (k_nw_scatter), 32 bytes per lane, coalesced.
// Cells right of column n inside the last group of four are computed and ignored: nothing to their left depends on them.
constexpr int NWT_SLICE = 2112;           // shared-memory words per warp: 66 per lane = the rows of the widest alignment (4 * 16 + 1)
constexpr int NWT_FLAG_WORDS = 256;       // global flag scratch per lane for the shapes whose flags do not fit: 64 rows x 4 words

__device__ __forceinline__ uint32_t even_bits16(uint32_t y)
{   // bits 0,2,4,..,30 of y -> bits 0..15
    y &= 0x55555555u;
    y = (y | (y >> 1)) & 0x33333333u;
    y = (y | (y >> 2)) & 0x0F0F0F0Fu;
    y = (y | (y >> 4)) & 0x00FF00FFu;
    return (y | (y >> 8)) & 0x0000FFFFu;
}

__global__ void __launch_bounds__(NWT_THREADS)
k_nw_thread(DevIndex ix, const uint8_t *__restrict__ codes, const NwSorted *__restrict__ sorted, const BatchCtl *__restrict__ ctl, int round,
            uint32_t *next_chunk, uint32_t *__restrict__ gflags, uint8_t *ops, int32_t *nops)
{
    if (ctl->abort) return;
    __shared__ uint32_t s_all[NWT_THREADS / 32][NWT_SLICE];
    const int lane = threadIdx.x & 31;
    uint32_t *sm = s_all[threadIdx.x >> 5];
    // this warp's flag scratch in global memory, laid out [word][lane] like the shared slice: a warp's store of one flag word
    // is one 128-byte line (the first version's per-job slices cost 32 sectors per store), its loads stay in L1 / L2
    uint32_t *gw = gflags + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 * NWT_FLAG_WORDS);
    const int n_jobs = ctl->nw_small[round];              // the shape-sorted jobs of this class come first
    const int n_chunks = (n_jobs + 31) / 32;
    // chunks of 32 shape-sorted jobs are handed out dynamically to the WARPS, the largest shapes first, so that the grid
    // finishes together
    for (;;) {
        int chunk = 0;
        if (lane == 0) chunk = (int)atomicAdd(next_chunk, 1u);
        chunk = __shfl_sync(FULLM, chunk, 0);
        if (chunk >= n_chunks) break;
        const int k0 = (n_chunks - 1 - chunk) * 32;
        NwSorted J{};
        if (k0 + lane < n_jobs) J = sorted[k0 + lane];
        const int cnt = min(32, n_jobs - k0);
        // words per alignment of this chunk: rows + flags in the shared slice when 32 alignments fit (66 words each); otherwise
        // only the rows stay on chip and the flags go to the warp's global scratch.  Either way all 32 lanes run.
        const int my_need = k0 + lane < n_jobs ? (4 * ((J.n + 3) >> 2) + 1) + J.m * ((J.n + 15) >> 4) : 0;
        const int need = __reduce_max_sync(FULLM, my_need);
        const bool flags_on_chip = need * 32 <= NWT_SLICE;
        constexpr int lpp = 32;
        for (int p0 = 0; p0 < cnt; p0 += lpp) {
            __syncwarp();
            // the alignment of this lane in this pass
            const int src = p0 + lane;
            const bool on = lane < lpp && src < cnt;
            NwSorted A;
            A.s1_off = __shfl_sync(FULLM, J.s1_off, src & 31); A.gpos = __shfl_sync(FULLM, J.gpos, src & 31);
            A.op_off = __shfl_sync(FULLM, J.op_off, src & 31); A.job = __shfl_sync(FULLM, J.job, src & 31);
            const int mn = __shfl_sync(FULLM, (int)(uint16_t)J.m | (int)J.n << 16, src & 31);
            if (!on) continue;
            const int m = mn & 0xFFFF, n = mn >> 16;
            const int wpr = (n + 15) >> 4, n4 = (n + 3) >> 2;
            uint32_t *row = sm + lane;                               // row[j] at row[j * lpp]
            uint32_t *flg = flags_on_chip ? sm + (4 * n4 + 1) * lpp + lane : gw + lane;   // flag word w at flg[w * lpp]
            // ---- genome window -> bit planes (bit j = low / high bit of base j) ----
            uint64_t glo = 0, ghi = 0;
            if (A.gpos >= 0 && A.gpos + 80 <= 2 * ix.G) {
                const uint32_t *w = ix.ref2 + (A.gpos >> 4);
                const int sh = 2 * (int)(A.gpos & 15);
                uint32_t prev = __ldg(w);
                for (int q = 0; q < wpr; q++) {
                    const uint32_t nxt = __ldg(w + q + 1);
                    const uint32_t x = __funnelshift_l(nxt, prev, sh);   // 16 bases, the first in the top bits
                    const uint32_t y = __brev(x);                         // base k: high bit at 2k, low bit at 2k+1
                    ghi |= (uint64_t)even_bits16(y) << (16 * q);
                    glo |= (uint64_t)even_bits16(y >> 1) << (16 * q);
                    prev = nxt;
                }
            } else {
                for (int j = 0; j < n; j++) { const uint64_t b = (uint64_t)ref_base(ix, A.gpos + j); glo |= (b & 1) << j; ghi |= (b >> 1) << j; }
            }
            // ---- row 0: S[0][j] = -2 - j; the T of row 1 = max(-inf - 1, S[0][j] - 3) = -5 - j ----
            for (int j = 1; j <= 4 * n4; j++) row[j * lpp] = (uint32_t)((-2 - j) & 0xFFFF) | (uint32_t)(-5 - j) << 16;
            const uint8_t *s1 = codes + A.s1_off;
            for (int i = 1; i <= m; i++) {
                const int a = (int)s1[i - 1];
                // columns whose genome base equals the read base of this row (codes 8..11 are lower-case ACGT; bit 2 = not ACGT)
                const uint64_t alo = (a & 1) ? ~0ull : 0ull, ahi = (a & 2) ? ~0ull : 0ull;
                uint64_t eq = (a & 4) ? 0ull : ~((glo ^ alo) | (ghi ^ ahi));
                int Sl = -2 - i, Rn = NWT_NEG, Sd = i == 1 ? 0 : -1 - i;      // S[i][0], R[i][1] before its max, S[i-1][0]
                uint32_t fw = 0;
                uint32_t *rp = row + lpp, *fp = flg + (i - 1) * wpr * lpp;
                for (int g = 0; g < n4; g++) {
                    const uint32_t e4 = (uint32_t)eq & 15u;
                    eq >>= 4;
                    uint32_t f8 = 0;
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        const uint32_t wv = rp[c * lpp];
                        const int Su = (int)(int16_t)(wv & 0xFFFFu), T = (int)wv >> 16;
                        const int R = __viaddmax_s32(Rn, -1, Sl - 3);
                        const int d = Sd + ((e4 & (1u << c)) ? 3 : -3);
                        const int h = __vimax3_s32(d, R, T);
                        const int S = (h + (int)((uint32_t)h >> 31)) & ~1;          // (h / 2) * 2, truncating toward zero
                        if (S == R) f8 |= 1u << (2 * c);
                        if (S == T) f8 |= 2u << (2 * c);
                        const int Tn = __viaddmax_s32(T, -1, S - 3);
                        rp[c * lpp] = (uint32_t)(S & 0xFFFF) | (uint32_t)Tn << 16;
                        Sd = Su; Sl = S; Rn = R;
                    }
                    rp += 4 * lpp;
                    fw |= f8 << (8 * (g & 3));
                    if ((g & 3) == 3 || g == n4 - 1) { *fp = fw; fp += lpp; fw = 0; }
                }
            }
            // ---- traceback (nw_alignment.cpp:61-74): S == R first (gap in the read string), then S == T, else diagonal ----
            int ti = m, tj = n, cntc = 0;
            int64_t pos = A.op_off + m + n;
            while (ti > 0 || tj > 0) {
                int op;
                if (ti == 0) op = 1;
                else if (tj == 0) op = 2;
                else {
                    const uint32_t f = (flg[((ti - 1) * wpr + ((tj - 1) >> 4)) * lpp] >> (((tj - 1) & 15) * 2)) & 3u;
                    op = (f & 1u) ? 1 : ((f & 2u) ? 2 : 0);
                }
                ops[--pos] = (uint8_t)op;
                cntc++;
                if (op == 1) tj--; else if (op == 2) ti--; else { ti--; tj--; }
            }
            nops[A.job] = cntc;
        }
    }
}

// INT32 pipe microbenchmark for the NW roofline: 8 independent add/max chains per thread (the recurrence's operation
// mix), every SM full.  Returns integer add/max operations per second (SURVEY.md §8d asks for a measured peak, not the
// spec sheet).
__global__ void __launch_bounds__(256) k_int32_peak(int *out, int iters, int seed)
{
    int a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const int b = seed + blockIdx.x, c = seed - 3;
    for (int i = 0; i < iters; i++) {
        a0 = max(a0 + b, a1); a1 = max(a1 + c, a2); a2 = max(a2 + b, a3); a3 = max(a3 + c, a4);
        a4 = max(a4 + b, a5); a5 = max(a5 + c, a6); a6 = max(a6 + b, a7); a7 = max(a7 + c, a0);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}

double measure_int32_ops_per_second(cudaStream_t st)
{
    const int grid = sm_count() * 8, iters = 1 << 14;
    DevBuf<int> out;
    out.reserve((size_t)grid * 256);
    cudaEvent_t e0, e1;
    DG_CUDA(cudaEventCreate(&e0)); DG_CUDA(cudaEventCreate(&e1));
    k_int32_peak<<<grid, 256, 0, st>>>(out.p, iters, 1);                 // warm-up
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        DG_CUDA(cudaEventRecord(e0, st));
        k_int32_peak<<<grid, 256, 0, st>>>(out.p, iters, rep + 2);
        DG_CUDA(cudaEventRecord(e1, st));
        DG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        DG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return (double)grid * 256 * iters * 16.0 / (best * 1e-3);            // 8 adds + 8 max per iteration
}

// ---- L2 gather microbenchmark: the roof of k_search on an L2-resident index ("L2 and INT32 peaks must be
// measured the same way by a microbenchmark on the box").  Every thread issues independent 32-byte (one sector) loads at
// pseudo-random block addresses of a table that fits L2 -- the access pattern of a rank query (seed_kernels.cu load_block).
__global__ void __launch_bounds__(256) k_l2_gather(const uint4 *__restrict__ table, uint32_t n_blocks_mask, int iters, uint32_t *out)
{
    uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
    uint32_t acc = 0;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {            // 8 independent sector loads in flight per thread
            x = x * 1664525u + 1013904223u;
            const uint32_t blk = (x >> 7) & n_blocks_mask;
            uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                         : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                         : "l"(reinterpret_cast<const char *>(table) + (size_t)blk * 32));
            acc += r0 ^ r7;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

double measure_l2_gather_bytes_per_second(cudaStream_t st, size_t table_bytes)
{
    size_t blocks = 1;
    while (blocks * 2 * 32 <= table_bytes) blocks <<= 1;          // power of two, <= table_bytes
    const int grid = sm_count() * 8, iters = 512;
    DevBuf<uint4> table; DevBuf<uint32_t> out;
    table.reserve(blocks * 2); out.reserve((size_t)grid * 256);
    DG_CUDA(cudaMemsetAsync(table.p, 1, blocks * 32, st));
    cudaEvent_t e0, e1;
    DG_CUDA(cudaEventCreate(&e0)); DG_CUDA(cudaEventCreate(&e1));
    k_l2_gather<<<grid, 256, 0, st>>>(table.p, (uint32_t)(blocks - 1), iters, out.p);     // warm-up: pulls the table into L2
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        DG_CUDA(cudaEventRecord(e0, st));
        k_l2_gather<<<grid, 256, 0, st>>>(table.p, (uint32_t)(blocks - 1), iters, out.p);
        DG_CUDA(cudaEventRecord(e1, st));
        DG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        DG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return (double)grid * 256 * iters * 8 * 32.0 / (best * 1e-3);
}

int nw_grid_warps() { return sm_count() * 8 * (NW_THREADS / 32); }

void launch_nw(const DevIndex &ix, const uint8_t *codes, const NwRound &R, BatchCtl *ctl, NwScratch &S, cudaStream_t st)
{
    if (R.cap_jobs <= 0) return;
    const int sms = sm_count();
    if (!S.hist.p) {                                   // once per context; afterwards k_nw_bins leaves the histogram zeroed
        S.hist.reserve(NW_BINS + 1); S.counter.reserve(4);
        launch_zero(S.hist.p, (NW_BINS + 1) * sizeof(uint32_t), st);
    }
    S.bin_start.reserve(NW_BINS + 1); S.bin_cur.reserve(NW_BINS + 1); S.order.reserve(R.cap_jobs); S.sorted.reserve(R.cap_jobs);
    int g0 = (R.cap_jobs + 255) / 256; if (g0 > sms * 8) g0 = sms * 8;
    k_nw_prepare<<<g0, 256, 0, st>>>(R.jobs, R.cap_jobs, R.round, R.with_aux, S.hist.p, ctl);
    k_nw_bins<<<1, 1024, 0, st>>>(S.hist.p, S.bin_start.p, S.bin_cur.p, S.counter.p, R.round, R.cap_ops, R.cap_flags, R.cap_aux, ctl);
    k_nw_scatter<<<g0, 256, 0, st>>>(R.jobs, R.cap_jobs, R.round, S.bin_cur.p, S.order.p, S.sorted.p, ctl);
    int gt = (R.cap_jobs + NWT_THREADS - 1) / NWT_THREADS; if (gt > sms * 6) gt = sms * 6;      // 6 CTAs of 33 KB shared memory per SM
    S.gflags.reserve((size_t)gt * (NWT_THREADS / 32) * 32 * NWT_FLAG_WORDS);
    k_nw_thread<<<gt, NWT_THREADS, 0, st>>>(ix, codes, S.sorted.p, ctl, R.round, S.counter.p, S.gflags.p, R.ops, R.nops);
    // everything larger: a warp per job
    int want = (R.cap_jobs + (NW_THREADS / 32) - 1) / (NW_THREADS / 32);
    int grid = want < sms * 8 ? want : sms * 8;
    k_nw<<<grid, NW_THREADS, 0, st>>>(ix, codes, R.jobs, S.order.p, ctl, R.round, R.cap_jobs, R.flags, R.rowbuf, R.rowbuf_per_warp, R.ops, R.nops);
}

} // namespace dartgpu
