// Device-side orchestration of the per-read path: everything between the seeding kernels and the final records runs
// on the GPU, one thread per candidate (repair pipeline phases A-D of report_logic.cuh) or per read / pair (candidate
// pairing, best / second best, mate rescue, flags, MAPQ, junction records, CIGAR text).  The NW and 8-mer jobs the
// phases emit go into HBM queues consumed by k_kmer / k_nw; only the final records cross PCIe.
//
// Replaces GenMappingReport and its callers' per-read logic (/root/reference/src/AlignmentCandidates.cpp:1079-1207,
// /root/reference/src/Mapping.cpp:600-621) — see report_logic.cuh for the line-by-line map.
#include <algorithm>
#include <chrono>
#include <cstdio>

#include "context.h"
#include "report_logic.cuh"

namespace dartgpu {

namespace {

constexpr int TPB = 128;
inline int grid_for(int64_t n) { int64_t g = (n + TPB - 1) / TPB; return (int)std::max<int64_t>(1, std::min<int64_t>(g, 148 * 32)); }

// ---- candidate table ----
__global__ void k_cand_init(int n_reads, const int64_t *seed_off, const int64_t *cand_off, const uint32_t *ncand,
                            const int32_t *cbegin, const int32_t *ccount, const int32_t *cscore, const uint64_t *keys, CandState *cs)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
        const int64_t so = seed_off[r], co = cand_off[r];
        const int nc = (int)ncand[r];
        for (int k = 0; k < nc; k++) {
            CandState c;
            c.read = r; c.seed_begin = (int32_t)(so + cbegin[so + k]); c.seed_count = ccount[so + k]; c.Score = cscore[so + k];
            c.PairedIdx = -1; c.SJtype = -1;
            uint64_t key = keys[c.seed_begin];
            int64_t pd = key_gpos(key) - key_rpos(key);
            c.PosDiff = pd < 0 ? 0 : pd;
            c.pos = 0; c.sv_off = 0; c.cig_off = 0; c.text_off = 0; c.sv_n = 0; c.sv_cap = 0; c.cig_cap = 0; c.cig_n = 0; c.text_len = 0;
            c.AlnScore = 0; c.mis = 0; c.chr = 0; c.n_ext = 0; c.live = 0; c.skip = 0; c.dir = 0; c.pad = 0;
            cs[co + k] = c;
        }
    }
}

__global__ void k_pair_prune(int n_units, int paired, const int64_t *cand_off, CandState *cs)
{
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        if (paired) {
            int64_t a = cand_off[2 * u], b = cand_off[2 * u + 1], e = cand_off[2 * u + 2];
            pair_and_prune(cs + a, (int)(b - a), cs + b, (int)(e - b), true);
        } else {
            int64_t a = cand_off[u], e = cand_off[u + 1];
            pair_and_prune(cs + a, (int)(e - a), nullptr, 0, false);
        }
    }
}

__global__ void k_cand_live(int64_t ncand, CandState *cs, uint32_t *cap)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= ncand; c += (int64_t)gridDim.x * blockDim.x) {
        if (c == ncand) { cap[c] = 0; continue; }
        int live = cs[c].Score != 0;
        cs[c].live = (uint8_t)live;
        int k = live ? seed_capacity(cs[c].seed_count) : 0;
        cs[c].sv_cap = k;
        cap[c] = (uint32_t)k;
    }
}

__global__ void k_set_sv_off(int64_t ncand, CandState *cs, const int64_t *off)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncand; c += (int64_t)gridDim.x * blockDim.x) cs[c].sv_off = off[c];
}

// One thread per candidate.  The candidate's record is worked on in registers and its seeds, when few (the normal
// case: 1-4 seeds), in a shared-memory slot; larger candidates work in place on their HBM slice.
// A candidate that emits no job in a phase does not have to wait for that phase's batched k-mer / NW launch: it goes
// straight on to the next phase in the same thread, up to and including C (D needs the CIGAR offsets of a scan).  On
// BASELINE config[1] nearly every candidate is a single seed pair with nothing to repair, so A, B and C run in one pass
// over the candidate table and the later kernels skip it on a one-byte stage marker (round-1 launch list: the four
// phases were 1.6 ms of a 5.6 ms step, each re-reading and re-writing the 104-byte record and its seeds).
// Occupancy: the phases are latency-bound chains of dependent loads; C keeps the most state live (96 registers), and any
// kernel that may run it takes its configuration (5 CTAs per SM, 8 staged seeds); D alone runs 8 CTAs per SM.
template <int WHICH> struct PhaseCfg { static constexpr int STAGE = WHICH == 3 ? 5 : 8, MIN_CTAS = WHICH == 3 ? 8 : 5; };
template <int WHICH>
__global__ void __launch_bounds__(TPB, PhaseCfg<WHICH>::MIN_CTAS) k_phase(Env E, int64_t ncand)
{
    constexpr int STAGE_SEEDS = PhaseCfg<WHICH>::STAGE;
    __shared__ RSeed s_slot[TPB * STAGE_SEEDS];
    for (int64_t cid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; cid < ncand; cid += (int64_t)gridDim.x * blockDim.x) {
        if (WHICH != 0 && WHICH != 3 && E.stage[cid] != WHICH) continue;    // already past this phase
        CandState c = E.cs[cid];
        if (!c.live) {
            if (WHICH == 0) E.stage[cid] = 3;
            if (WHICH == 3) { E.cs[cid].AlnScore = 0; E.cs[cid].cig_n = 0; E.cs[cid].text_len = 0; }
            continue;
        }
        RSeed *g = E.pool + c.sv_off, *sv = g;
        const bool staged = phase_seed_bound(c, WHICH) <= STAGE_SEEDS;
        if (staged) {
            sv = s_slot + threadIdx.x * STAGE_SEEDS;
            if (WHICH != 0) for (int i = 0; i < c.sv_n; i++) sv[i] = g[i];
        }
        int next = WHICH + 1;
        if (WHICH == 0) {
            phase_a(E, c, sv);
            bool waits = false;
            for (int i = 1; i < c.sv_n; i++) waits |= sv[i].job >= 0;
            if (!waits && (!staged || phase_seed_bound(c, 1) <= STAGE_SEEDS)) { phase_b(E, c, sv); next = 2; }
        }
        if (WHICH == 1) phase_b(E, c, sv);
        if ((WHICH == 0 && next == 2) || WHICH == 1) {
            if (c.n_ext == 0 && (!staged || phase_seed_bound(c, 2) <= STAGE_SEEDS)) {
                Env E2 = E; E2.njobs = E.njobs_c; E2.njob_count = E.njob_count_c;
                phase_c(E2, c, sv); next = 3;
            }
        }
        if (WHICH == 2) phase_c(E, c, sv);
        if (WHICH == 3) phase_d(E, c, sv);
        if (staged) for (int i = 0; i < c.sv_n; i++) g[i] = sv[i];
        E.cs[cid] = c;
        if (WHICH != 3) E.stage[cid] = (uint8_t)next;
    }
}

// ---- NW job bookkeeping: sizes -> (scan) -> offsets ----
__global__ void k_nw_sizes(const NwJobDev *jobs, int n, int with_aux, uint32_t *s_ops, uint32_t *s_flags, uint32_t *s_aux)
{
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += gridDim.x * blockDim.x) {
        if (j == n) { s_ops[j] = s_flags[j] = s_aux[j] = 0; continue; }
        const int m = jobs[j].m, nn = jobs[j].n;
        s_ops[j] = (uint32_t)(m + nn);
        s_flags[j] = (uint32_t)(m * ((nn + 15) >> 4));
        s_aux[j] = (with_aux && !(j & 1)) ? (uint32_t)(2 * (m + 1)) : 0u;
    }
}
__global__ void k_nw_offsets(NwJobDev *jobs, int n, const int64_t *o_ops, const int64_t *o_flags, const int64_t *o_aux, int32_t *max_n, int32_t *any_multi,
                             unsigned long long *cells)
{
    unsigned long long mine = 0;
    int mx = 0, multi = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        jobs[j].op_off = o_ops[j]; jobs[j].flag_off = o_flags[j]; jobs[j].aux_off = o_aux[j];
        mx = max(mx, jobs[j].n);
        multi |= jobs[j].m > 32 || jobs[j].n > 64;          // leaves the register-resident path of k_nw
        mine += (unsigned long long)jobs[j].m * jobs[j].n;
    }
    // one atomic per warp, not per job (round-1 launch list: 427 us of contention on two addresses)
    mx = __reduce_max_sync(0xffffffffu, mx);
    multi = __any_sync(0xffffffffu, multi);
    for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, d);
    if ((threadIdx.x & 31) == 0) {
        if (mx > 0) atomicMax(max_n, mx);
        if (multi) atomicOr(any_multi, 1);
        if (mine) atomicAdd(cells, mine);
    }
}

__global__ void k_kmer_work(const KmerJobDev *jobs, int n, unsigned long long *acc)
{
    unsigned long long w = 0, r = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) { w += jobs[j].len2; r += jobs[j].len1; }
    if (w) atomicAdd(acc + 1, w);
    if (r) atomicAdd(acc + 2, r);
}

__global__ void k_cig_caps(int64_t ncand, const CandState *cs, uint32_t *cap)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= ncand; c += (int64_t)gridDim.x * blockDim.x)
        cap[c] = (c < ncand && cs[c].live && !cs[c].skip) ? (uint32_t)cs[c].cig_cap : 0u;
}
__global__ void k_set_cig_off(int64_t ncand, CandState *cs, const int64_t *off)
{
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c < ncand; c += (int64_t)gridDim.x * blockDim.x) cs[c].cig_off = off[c];
}

// ---- final pass ----
__global__ void k_report_counts(int n_reads, const uint32_t *ncand, uint32_t *nrep)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_reads; r += gridDim.x * blockDim.x)
        nrep[r] = r < n_reads ? (ncand[r] ? ncand[r] : 1u) : 0u;
}

__global__ void k_read_final(Env E, int n_units, int paired, const int64_t *cand_off, const int64_t *rep_off, dartgpu_read_result *rr,
                             dartgpu_report *rep, uint32_t *text_len, uint32_t *njunc, int64_t n_rep_total, int n_reads, int32_t *err)
{
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        const int nr = paired ? 2 : 1;
        ReadOut ro[2];
        for (int m = 0; m < nr; m++) {
            int r = paired ? 2 * u + m : u;
            read_best(E.cs + cand_off[r], (int)(cand_off[r + 1] - cand_off[r]), rep + rep_off[r], ro[m]);
        }
        if (!paired) finish_single(ro[0], rep + rep_off[u], E.cs + cand_off[u], (int)(cand_off[u + 1] - cand_off[u]));
        else {
            int a = 2 * u, b = a + 1;
            finish_pair(ro[0], rep + rep_off[a], E.cs + cand_off[a], (int)(cand_off[a + 1] - cand_off[a]),
                        ro[1], rep + rep_off[b], E.cs + cand_off[b], (int)(cand_off[b + 1] - cand_off[b]), E.P.multi_hit != 0);
        }
        for (int m = 0; m < nr; m++) {
            int r = paired ? 2 * u + m : u;
            dartgpu_read_result o;
            o.mapq = (uint8_t)ro[m].mapq; o.score = (int16_t)ro[m].score; o.sub_score = (int16_t)ro[m].sub_score; o.mis_num = (int16_t)ro[m].mis_num;
            o.n_reports = ro[m].n_reports; o.best = ro[m].best; o.report_off = rep_off[r]; o.reserved = 0;
            rr[r] = o;
            const int nc = (int)(cand_off[r + 1] - cand_off[r]);
            for (int k = 0; k < ro[m].n_reports; k++) {
                int tl = 0;
                if (k < nc) {
                    const CandState &c = E.cs[cand_off[r] + k];
                    if (c.cig_n < 0) atomicOr(err, 1);
                    if (c.live && !c.skip && c.AlnScore > 0) tl = c.text_len;
                }
                text_len[rep_off[r] + k] = (uint32_t)tl;
                rep[rep_off[r] + k].cigar_len = (int16_t)tl;
            }
            njunc[r] = (uint32_t)emit_junctions(E, ro[m], E.cs + cand_off[r], nc, r, nullptr);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) { text_len[n_rep_total] = 0; njunc[n_reads] = 0; }
}

__global__ void k_write_records(Env E, int n_reads, const int64_t *cand_off, const dartgpu_read_result *rr, dartgpu_report *rep,
                                const int64_t *text_off, char *text, const int64_t *junc_off, dartgpu_junction *junc)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_reads; r += gridDim.x * blockDim.x) {
        const dartgpu_read_result o = rr[r];
        const int nc = (int)(cand_off[r + 1] - cand_off[r]);
        for (int k = 0; k < o.n_reports; k++) {
            dartgpu_report &p = rep[o.report_off + k];
            p.cigar_off = (int32_t)text_off[o.report_off + k];
            if (p.cigar_len > 0 && k < nc) {
                const CandState &c = E.cs[cand_off[r] + k];
                write_cigar_text(E.cig + c.cig_off, c.cig_n, text + p.cigar_off);
            }
        }
        if (junc_off[r + 1] > junc_off[r]) {
            ReadOut ro; ro.mapq = o.mapq; ro.score = o.score; ro.sub_score = o.sub_score; ro.mis_num = o.mis_num; ro.best = o.best; ro.n_reports = o.n_reports;
            emit_junctions(E, ro, E.cs + cand_off[r], nc, r, junc + junc_off[r]);
        }
    }
}

} // namespace

// buffers owned by the device pipeline (kept across calls inside the context)
struct DevicePipe {
    DevBuf<int64_t> cand_off, sv_off, scan_a, scan_b, scan_c, rep_off, text_off, junc_off, cig_off;
    DevBuf<uint32_t> u32_a, u32_b, u32_c, text_len, njunc;
    DevBuf<CandState> cs;
    DevBuf<RSeed> pool;
    DevBuf<uint8_t> stage;
    DevBuf<KmerJobDev> kjobs;
    DevBuf<dartgpu_kmer_hit> khits;
    DevBuf<NwJobDev> jobsB, jobsC;
    DevBuf<uint8_t> opsB, opsC;
    DevBuf<int32_t> nopsB, nopsC, aux, cig, counters;
    DevBuf<unsigned long long> work;   // [0] NW cells, [1] 8-mer window bases, [2] 8-mer read bases
    PinBuf<unsigned long long> h_work;
    DevBuf<uint32_t> flags;
    DevBuf<int32_t> rowbuf;
    DevBuf<dartgpu_read_result> rr;
    DevBuf<dartgpu_report> rep;
    DevBuf<char> text;
    DevBuf<dartgpu_junction> junc;
    DevBuf<int64_t> d_chr_fwd; DevBuf<int32_t> d_end_chr;
    DevBuf<uint8_t> scan_tmp;
    PinBuf<int64_t> h_vals;
    PinBuf<int32_t> h_counters;
    PinBuf<dartgpu_read_result> h_rr;
    PinBuf<dartgpu_report> h_rep;
    PinBuf<char> h_text;
    PinBuf<dartgpu_junction> h_junc;
    bool tables = false;
};

static DevicePipe *pipe_of(dartgpu_ctx *c)
{
    if (!c->dpipe) c->dpipe = new DevicePipe;
    return static_cast<DevicePipe *>(c->dpipe);
}
void free_device_pipe(void *p) { delete static_cast<DevicePipe *>(p); }

static void scan_u32(dartgpu_ctx *c, DevicePipe *D, const uint32_t *in, int64_t *out, int64_t n)
{   // out[0..n] = exclusive scan of in[0..n], in[n] must be 0
    size_t tmp = scan_tmp_bytes((int)n);
    D->scan_tmp.reserve(tmp + 256);
    launch_scan_u32_to_i64(in, out, (int)n, D->scan_tmp.p, tmp, c->stream);
}

static int64_t fetch_i64(dartgpu_ctx *c, DevicePipe *D, const int64_t *dev)
{
    small_d2h(D->h_vals.p, dev, sizeof(int64_t), c->stream);
    DG_CUDA(dg_stream_sync(c->stream));
    return D->h_vals.p[0];
}

// runs one round of NW jobs that already sit in `jobs` on the device
static void nw_round(dartgpu_ctx *c, DevicePipe *D, NwJobDev *jobs, int nj, bool with_aux, DevBuf<uint8_t> &ops, DevBuf<int32_t> &nops)
{
    cudaStream_t st = c->stream;
    nops.reserve(nj + 1);
    if (nj == 0) { ops.reserve(1); if (with_aux) D->aux.reserve(1); return; }
    D->u32_a.reserve(nj + 1); D->u32_b.reserve(nj + 1); D->u32_c.reserve(nj + 1);
    D->scan_a.reserve(nj + 1); D->scan_b.reserve(nj + 1); D->scan_c.reserve(nj + 1);
    k_nw_sizes<<<grid_for(nj + 1), TPB, 0, st>>>(jobs, nj, with_aux ? 1 : 0, D->u32_a.p, D->u32_b.p, D->u32_c.p);
    scan_u32(c, D, D->u32_a.p, D->scan_a.p, nj);
    scan_u32(c, D, D->u32_b.p, D->scan_b.p, nj);
    scan_u32(c, D, D->u32_c.p, D->scan_c.p, nj);
    DG_CUDA(cudaMemsetAsync(D->counters.p + 4, 0, 2 * sizeof(int32_t), st));
    k_nw_offsets<<<grid_for(nj), TPB, 0, st>>>(jobs, nj, D->scan_a.p, D->scan_b.p, D->scan_c.p, D->counters.p + 4, D->counters.p + 5, D->work.p);
    small_d2h(D->h_vals.p + 0, D->scan_a.p + nj, 8, st);
    small_d2h(D->h_vals.p + 1, D->scan_b.p + nj, 8, st);
    small_d2h(D->h_vals.p + 2, D->scan_c.p + nj, 8, st);
    small_d2h(D->h_counters.p, D->counters.p + 4, 2 * sizeof(int32_t), st);
    DG_CUDA(dg_stream_sync(st));
    const int64_t ops_total = D->h_vals.p[0], flag_total = D->h_vals.p[1], aux_total = D->h_vals.p[2];
    const int max_n = D->h_counters.p[0];
    const bool multi = D->h_counters.p[1] != 0;
    ops.reserve(ops_total + 1); D->flags.reserve(flag_total + 1);
    if (with_aux) D->aux.reserve(aux_total + 1);
    size_t rb = multi ? (size_t)2 * (max_n + 1) : 0;
    D->rowbuf.reserve(rb * nw_grid_warps() + 1);
    DG_CUDA(cudaEventRecord(c->ev[10], st));
    launch_nw(c->ix, c->d_codes.p, jobs, nj, D->flags.p, D->rowbuf.p, rb, ops.p, nops.p, c->nwscratch, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[11], st));
    DG_CUDA(dg_stream_sync(st));
    add_ms(c, &c->stats.ms_nw, c->ev[10], c->ev[11]);
    c->stats.kernel_launches += 4 + NW_LAUNCHES;
    c->stats.nw_jobs += nj;
}

void run_pipeline_device(dartgpu_ctx *c, dartgpu_map_result *out)
{
    DevicePipe *D = pipe_of(c);
    cudaStream_t st = c->stream;
    const int n = c->n_reads;
    const dartgpu_params &P = c->prm;
    const int paired = P.pair_end != 0;
    const int units = paired ? n / 2 : n;
    D->h_vals.reserve(8); D->h_counters.reserve(8); D->counters.reserve(8); D->work.reserve(4); D->h_work.reserve(4);
    DG_CUDA(cudaMemsetAsync(D->work.p, 0, 4 * sizeof(unsigned long long), st));
    if (!D->tables) {
        D->d_chr_fwd.reserve(c->shared->chr_fwd.size()); D->d_end_chr.reserve(c->shared->end_chr.size());
        DG_CUDA(cudaMemcpyAsync(D->d_chr_fwd.p, c->shared->chr_fwd.data(), c->shared->chr_fwd.size() * 8, cudaMemcpyHostToDevice, st));
        std::vector<int32_t> ec(c->shared->end_chr.begin(), c->shared->end_chr.end());
        DG_CUDA(cudaMemcpyAsync(D->d_end_chr.p, ec.data(), ec.size() * 4, cudaMemcpyHostToDevice, st));
        DG_CUDA(dg_stream_sync(st));
        D->tables = true;
    }
    if (n == 0) { *out = dartgpu_map_result{nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0}; return; }
    DG_CUDA(cudaEventRecord(c->ev[12], st));

    // ---- candidate table ----
    D->cand_off.reserve(n + 2);
    DG_CUDA(cudaMemsetAsync(c->d_ncand.p + n, 0, sizeof(uint32_t), st));
    scan_u32(c, D, c->d_ncand.p, D->cand_off.p, n);
    const int64_t ncand = fetch_i64(c, D, D->cand_off.p + n);
    D->cs.reserve(ncand + 1);
    k_cand_init<<<grid_for(n), TPB, 0, st>>>(n, c->d_seed_off.p, D->cand_off.p, c->d_ncand.p, c->d_cand_begin.p, c->d_cand_count.p,
                                             c->d_cand_score.p, c->d_keys.p, D->cs.p);
    k_pair_prune<<<grid_for(units), TPB, 0, st>>>(units, paired, D->cand_off.p, D->cs.p);
    D->u32_a.reserve(ncand + 2); D->sv_off.reserve(ncand + 2);
    k_cand_live<<<grid_for(ncand + 1), TPB, 0, st>>>(ncand, D->cs.p, D->u32_a.p);
    scan_u32(c, D, D->u32_a.p, D->sv_off.p, ncand);
    const int64_t pool_total = fetch_i64(c, D, D->sv_off.p + ncand);
    k_set_sv_off<<<grid_for(ncand), TPB, 0, st>>>(ncand, D->cs.p, D->sv_off.p);
    D->pool.reserve(pool_total + 1);
    D->kjobs.reserve(pool_total / 12 + 2); D->khits.reserve(pool_total / 12 + 2);
    D->jobsB.reserve(pool_total / 3 + 4); D->jobsC.reserve(pool_total + 4);
    DG_CUDA(cudaMemsetAsync(D->counters.p, 0, 8 * sizeof(int32_t), st));
    c->stats.kernel_launches += 4 + NW_LAUNCHES;

    Env E{};
    E.P = PhaseParams{P.max_gaps, P.max_intron, P.min_intron, P.max_mismatch, P.multi_hit, P.pair_end, P.all_sj};
    E.ref = RefView{c->ix.ref2, nullptr, c->G}; E.G = c->G;
    E.ends = c->shared->d_ends.p; E.end_chr = D->d_end_chr.p; E.n_ends = (int)c->shared->ends.size(); E.chr_fwd = D->d_chr_fwd.p;
    E.codes = c->d_codes.p; E.code_off = c->d_dev_off.p; E.rlen = c->d_rlen.p; E.keys = c->d_keys.p;
    E.cs = D->cs.p; E.pool = D->pool.p;
    E.kjobs = D->kjobs.p; E.kjob_count = D->counters.p + 0; E.khits = D->khits.p;
    E.njobs = D->jobsB.p; E.njob_count = D->counters.p + 1;          // phase B's queue (also used by candidates fast-tracked out of A)
    E.njobs_c = D->jobsC.p; E.njob_count_c = D->counters.p + 2;      // phase C's queue
    D->stage.reserve(ncand + 1);
    DG_CUDA(cudaMemsetAsync(D->stage.p, 0, (size_t)ncand + 1, st));
    E.stage = D->stage.p;

    // ---- phase A -> 8-mer re-seeding ----
    k_phase<0><<<grid_for(ncand), TPB, 0, st>>>(E, ncand);
    DG_CUDA(cudaGetLastError());
    small_d2h(D->h_counters.p, D->counters.p, sizeof(int32_t), st);
    DG_CUDA(dg_stream_sync(st));
    const int nk = D->h_counters.p[0];
    if (nk > 0) {
        DG_CUDA(cudaEventRecord(c->ev[8], st));
        k_kmer_work<<<grid_for(nk), TPB, 0, st>>>(D->kjobs.p, nk, D->work.p);
        launch_kmer(c->ix, c->d_codes.p, D->kjobs.p, nk, std::max(c->max_rlen, 8), D->khits.p, c->kscratch, st);
        DG_CUDA(cudaGetLastError());
        DG_CUDA(cudaEventRecord(c->ev[9], st));
        DG_CUDA(dg_stream_sync(st));
        add_ms(c, &c->stats.ms_kmer, c->ev[8], c->ev[9]);
        c->stats.kernel_launches += KMER_LAUNCHES;
    }
    c->stats.kmer_jobs += nk;

    // ---- phase B -> NW of every gap against both flanks ----
    E.njobs = D->jobsB.p; E.njob_count = D->counters.p + 1;
    k_phase<1><<<grid_for(ncand), TPB, 0, st>>>(E, ncand);
    DG_CUDA(cudaGetLastError());
    small_d2h(D->h_counters.p, D->counters.p + 1, sizeof(int32_t), st);
    DG_CUDA(dg_stream_sync(st));
    const int nB = D->h_counters.p[0];
    nw_round(c, D, D->jobsB.p, nB, true, D->opsB, D->nopsB);

    // ---- phase C -> NW of every non-simple pair ----
    E.ops = D->opsB.p; E.nops = D->nopsB.p; E.done_jobs = D->jobsB.p; E.xscratch = D->aux.p;
    E.njobs = D->jobsC.p; E.njob_count = D->counters.p + 2;
    k_phase<2><<<grid_for(ncand), TPB, 0, st>>>(E, ncand);
    DG_CUDA(cudaGetLastError());
    small_d2h(D->h_counters.p, D->counters.p + 2, sizeof(int32_t), st);
    DG_CUDA(dg_stream_sync(st));
    const int nC = D->h_counters.p[0];
    nw_round(c, D, D->jobsC.p, nC, false, D->opsC, D->nopsC);

    // ---- phase D: CIGAR pairs, score, coordinates ----
    D->u32_a.reserve(ncand + 2); D->cig_off.reserve(ncand + 2);
    k_cig_caps<<<grid_for(ncand + 1), TPB, 0, st>>>(ncand, D->cs.p, D->u32_a.p);
    scan_u32(c, D, D->u32_a.p, D->cig_off.p, ncand);
    const int64_t cig_total = fetch_i64(c, D, D->cig_off.p + ncand);
    k_set_cig_off<<<grid_for(ncand), TPB, 0, st>>>(ncand, D->cs.p, D->cig_off.p);
    D->cig.reserve(cig_total + 1);
    E.ops = D->opsC.p; E.nops = D->nopsC.p; E.done_jobs = D->jobsC.p; E.cig = D->cig.p;
    k_phase<3><<<grid_for(ncand), TPB, 0, st>>>(E, ncand);
    DG_CUDA(cudaGetLastError());

    // ---- per read / pair: best, mate rescue, flags, MAPQ; record layout ----
    D->u32_b.reserve(n + 2); D->rep_off.reserve(n + 2);
    k_report_counts<<<grid_for(n + 1), TPB, 0, st>>>(n, c->d_ncand.p, D->u32_b.p);
    scan_u32(c, D, D->u32_b.p, D->rep_off.p, n);
    const int64_t nrep = fetch_i64(c, D, D->rep_off.p + n);
    D->rr.reserve(n + 1); D->rep.reserve(nrep + 1); D->text_len.reserve(nrep + 2); D->njunc.reserve(n + 2);
    D->text_off.reserve(nrep + 2); D->junc_off.reserve(n + 2);
    DG_CUDA(cudaMemsetAsync(D->counters.p + 6, 0, sizeof(int32_t), st));
    k_read_final<<<grid_for(units), TPB, 0, st>>>(E, units, paired, D->cand_off.p, D->rep_off.p, D->rr.p, D->rep.p, D->text_len.p, D->njunc.p,
                                                  nrep, n, D->counters.p + 6);
    DG_CUDA(cudaGetLastError());
    scan_u32(c, D, D->text_len.p, D->text_off.p, nrep);
    scan_u32(c, D, D->njunc.p, D->junc_off.p, n);
    small_d2h(D->h_vals.p + 0, D->text_off.p + nrep, 8, st);
    small_d2h(D->h_vals.p + 1, D->junc_off.p + n, 8, st);
    small_d2h(D->h_counters.p, D->counters.p + 6, sizeof(int32_t), st);
    DG_CUDA(dg_stream_sync(st));
    const int64_t text_total = D->h_vals.p[0], junc_total = D->h_vals.p[1];
    if (text_total >= (1ll << 31)) throw std::make_pair(DARTGPU_ERR_ARG, std::string("batch too large: CIGAR text exceeds 2 GB, split the batch"));
    if (D->h_counters.p[0]) throw std::make_pair(DARTGPU_ERR_CUDA, std::string("CIGAR pool capacity exceeded"));
    D->text.reserve(text_total + 1); D->junc.reserve(junc_total + 1);
    k_write_records<<<grid_for(n), TPB, 0, st>>>(E, n, D->cand_off.p, D->rr.p, D->rep.p, D->text_off.p, D->text.p, D->junc_off.p, D->junc.p);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[13], st));
    c->stats.kernel_launches += 16;

    // ---- only the final records cross PCIe ----
    D->h_rr.reserve(n + 1); D->h_rep.reserve(nrep + 1); D->h_text.reserve(text_total + 1); D->h_junc.reserve(junc_total + 1);
    static const bool trace = getenv("DARTGPU_TRACE") != nullptr;   // host timeline of the compute / result-copy phases (adds a sync)
    const double t_enq = trace ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0;
    if (trace) DG_CUDA(dg_stream_sync(st));
    const double t_cmp = trace ? std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count() : 0;
    DG_CUDA(cudaMemcpyAsync(D->h_rr.p, D->rr.p, (size_t)n * sizeof(dartgpu_read_result), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(D->h_rep.p, D->rep.p, (size_t)nrep * sizeof(dartgpu_report), cudaMemcpyDeviceToHost, st));
    if (text_total) DG_CUDA(cudaMemcpyAsync(D->h_text.p, D->text.p, text_total, cudaMemcpyDeviceToHost, st));
    if (junc_total) DG_CUDA(cudaMemcpyAsync(D->h_junc.p, D->junc.p, (size_t)junc_total * sizeof(dartgpu_junction), cudaMemcpyDeviceToHost, st));
    small_d2h(D->h_work.p, D->work.p, 4 * sizeof(unsigned long long), st);
    DG_CUDA(cudaEventRecord(c->ev[14], st));
    DG_CUDA(dg_stream_sync(st));
    if (trace) {
        const double t_done = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
        fprintf(stderr, "TRACE ctx %p enq %.3f compute_done %.3f d2h_done %.3f\n", (void *)c, t_enq, t_cmp, t_done);
    }
    c->stats.nw_cells += D->h_work.p[0]; c->stats.kmer_window_bases += D->h_work.p[1]; c->stats.kmer_read_bases += D->h_work.p[2];
    add_ms(c, &c->stats.ms_report, c->ev[12], c->ev[13]);
    add_ms(c, &c->stats.ms_d2h, c->ev[13], c->ev[14]);
    c->stats.ms_report -= c->stats.ms_kmer + c->stats.ms_nw;   // the phase kernels alone
    c->stats.d2h_bytes += (uint64_t)n * sizeof(dartgpu_read_result) + (uint64_t)nrep * sizeof(dartgpu_report) + text_total + junc_total * sizeof(dartgpu_junction);
    out->reads = D->h_rr.p; out->n_reads = n;
    out->reports = D->h_rep.p; out->n_reports = nrep;
    out->cigars = D->h_text.p; out->n_cigar_bytes = text_total;
    out->junctions = D->h_junc.p; out->n_junctions = junc_total;
}

} // namespace dartgpu
