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
// One warp per job, many jobs in flight (jobs are small: tens to a few hundred cells on average).  The matrix
// is swept in strips of 32 rows; inside a strip the warp advances along anti-diagonals: lane l owns row l of
// the strip and at step t computes column t-l.  Left neighbours stay in registers, upper neighbours arrive
// by shuffle from lane l-1, the genome base is handed down the lanes systolically.  Integer pipes only:
// ~20 integer ops + 4 shuffles per cell-step.  Two traceback bits per cell: in registers for single-strip jobs
// (m <= 32, n <= 64 — walked back with shuffles), in global scratch for larger ones (lane 0 walks them back).
#include "dartgpu_internal.h"

namespace dartgpu {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int NW_THREADS = 128;
constexpr int NW_NEG = -131072;

__device__ __forceinline__ int ref_base(const DevIndex &ix, int64_t p)
{
    if (p < 0 || p >= 2 * ix.G) return 0;
    return (int)((__ldg(ix.ref2 + (p >> 4)) >> (30 - 2 * (int)(p & 15))) & 3);
}

__global__ void __launch_bounds__(NW_THREADS)
k_nw(DevIndex ix, const uint8_t *__restrict__ codes, const NwJobDev *__restrict__ jobs, int n_jobs,
     uint32_t *flags, int32_t *rowbuf, size_t rowbuf_per_warp, uint8_t *ops, int32_t *nops)
{
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    int32_t *rb = rowbuf + (size_t)warp * rowbuf_per_warp;

    for (int job = warp; job < n_jobs; job += nwarps) {
        const NwJobDev J = jobs[job];
        const int m = J.m, n = J.n;
        if (m <= 32 && n <= 64) {
            // ---- small job (the overwhelming majority: a mismatch or a short indel between two seeds): one strip, the
            // traceback bits stay in registers (2 x 64 bits per row) and are walked back with shuffles — no global
            // round trips between the sweep and the traceback.
            const int i = lane + 1;
            const bool rowok = i <= m;
            int a = rowok ? (int)codes[J.s1_off + lane] : 7;
            a = (a & 4) ? 7 : (a & 3);
            const int b0reg = lane < n ? ref_base(ix, J.gpos + lane) : 0;
            const int b1reg = lane + 32 < n ? ref_base(ix, J.gpos + lane + 32) : 0;
            int Sl = -2 - i, Rl = NW_NEG, Sd = (i == 1) ? 0 : -2 - (i - 1);
            int So = 0, To = 0, bo = 0;
            uint64_t flo = 0, fhi = 0;
            const int steps = n + m - 1;
            for (int t = 0; t < steps; t++) {
                int Su = __shfl_up_sync(FULLM, So, 1);
                int Tu = __shfl_up_sync(FULLM, To, 1);
                int b = __shfl_up_sync(FULLM, bo, 1);
                const int b0 = t < 32 ? __shfl_sync(FULLM, b0reg, t) : __shfl_sync(FULLM, b1reg, t - 32);
                const int j = t - lane + 1;
                if (lane == 0) { Su = -2 - j; Tu = NW_NEG; b = b0; }
                if (rowok && j >= 1 && j <= n) {
                    int R = max(Rl - 1, Sl - 3);
                    int T = max(Tu - 1, Su - 3);
                    int h = max(Sd + (a == b ? 3 : -3), max(R, T));
                    int S = (h / 2) * 2;
                    uint64_t f = (S == R ? 1ull : 0ull) | (S == T ? 2ull : 0ull);
                    if (j <= 32) flo |= f << (2 * (j - 1)); else fhi |= f << (2 * (j - 33));
                    Sd = Su; Sl = S; Rl = R;
                    So = S; To = T; bo = b;
                }
            }
            int ti = m, tj = n, k = 0;                       // identical in every lane: the walk is warp-uniform
            int64_t pos = J.op_off + m + n;
            while (ti > 0 || tj > 0) {
                int op;
                if (ti == 0) op = 1;
                else if (tj == 0) op = 2;
                else {
                    uint64_t w = __shfl_sync(FULLM, tj <= 32 ? flo : fhi, ti - 1);
                    uint32_t f = (uint32_t)(w >> (2 * ((tj - 1) & 31))) & 3u;
                    op = (f & 1u) ? 1 : ((f & 2u) ? 2 : 0);
                }
                --pos;
                if (lane == 0) ops[pos] = (uint8_t)op;
                k++;
                if (op == 1) tj--; else if (op == 2) ti--; else { ti--; tj--; }
            }
            if (lane == 0) nops[job] = k;
            continue;
        }
        const int wpr = (n + 15) >> 4;
        uint32_t *fl = flags + J.flag_off;

        for (int s0 = 0; s0 < m; s0 += 32) {
            const int i = s0 + lane + 1;              // my row (1-based)
            const bool rowok = i <= m;
            int a = rowok ? (int)codes[J.s1_off + i - 1] : 7;
            a = (a & 4) ? 7 : (a & 3);                // 8..11 are lower-case ACGT: same base for the table compare
            int Sl = -2 - i, Rl = NW_NEG;             // S[i][0], R[i][0]
            int Sd = (i == 1) ? 0 : -2 - (i - 1);     // S[i-1][0]
            int So = 0, To = 0, bo = 0;               // what I hand to the lane below
            uint32_t fw = 0;
            const int rows = min(32, m - s0);
            const int steps = n + rows - 1;
            const bool last_strip = s0 + 32 >= m;
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
                int b0 = __shfl_sync(FULLM, bch, t & 31);
                int j = t - lane + 1;
                if (s0 > 0) {
                    int su0 = __shfl_sync(FULLM, such, t & 31), tu0 = __shfl_sync(FULLM, tuch, t & 31);
                    if (lane == 0) { Su = su0; Tu = tu0; }
                } else if (lane == 0) { Su = -2 - j; Tu = NW_NEG; }
                if (lane == 0) b = b0;
                if (rowok && j >= 1 && j <= n) {
                    int R = max(Rl - 1, Sl - 3);
                    int T = max(Tu - 1, Su - 3);
                    int h = max(Sd + (a == b ? 3 : -3), max(R, T));
                    int S = (h / 2) * 2;
                    uint32_t f = (S == R ? 1u : 0u) | (S == T ? 2u : 0u);
                    fw |= f << (((j - 1) & 15) * 2);
                    if (((j - 1) & 15) == 15 || j == n) { __stcg(fl + (size_t)(i - 1) * wpr + ((j - 1) >> 4), fw); fw = 0; }
                    Sd = Su; Sl = S; Rl = R;
                    So = S; To = T; bo = b;
                    if (lane == 31 && !last_strip) { __stcg(rb + j, S); __stcg(rb + (n + 1) + j, T); }
                }
            }
            __syncwarp();
        }
        __syncwarp();
        if (lane == 0) {
            int i = m, j = n, k = 0;
            int64_t pos = J.op_off + m + n;
            while (i > 0 || j > 0) {
                int op;
                if (i == 0) op = 1;
                else if (j == 0) op = 2;
                else {
                    uint32_t f = (__ldcg(fl + (size_t)(i - 1) * wpr + ((j - 1) >> 4)) >> (((j - 1) & 15) * 2)) & 3u;
                    op = (f & 1u) ? 1 : ((f & 2u) ? 2 : 0);
                }
                ops[--pos] = (uint8_t)op;
                k++;
                if (op == 1) j--; else if (op == 2) i--; else { i--; j--; }
            }
            nops[job] = k;
        }
        __syncwarp();
    }
}

int nw_grid_warps() { return 148 * 8 * (NW_THREADS / 32); }

void launch_nw(const DevIndex &ix, const uint8_t *codes, const NwJobDev *jobs, int n_jobs,
               uint32_t *flags, int32_t *rowbuf, size_t rowbuf_per_warp, uint8_t *ops, int32_t *nops, cudaStream_t st)
{
    if (n_jobs <= 0) return;
    int want = (n_jobs + (NW_THREADS / 32) - 1) / (NW_THREADS / 32);
    int grid = want < 148 * 8 ? want : 148 * 8;
    k_nw<<<grid, NW_THREADS, 0, st>>>(ix, codes, jobs, n_jobs, flags, rowbuf, rowbuf_per_warp, ops, nops);
}

} // namespace dartgpu
