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
inline int grid_for(int64_t n) { int64_t g = (n + TPB - 1) / TPB; return (int)std::max<int64_t>(1, std::min<int64_t>(g, (int64_t)sm_count() * 32)); }

// ---- candidate table ----
// reports per read = CanNum, at least one (an unmapped read still carries its FLAG): the prefix sums of both run over the reads
__global__ void k_report_counts(int n_reads, uint32_t *ncand, uint32_t *nrep)
{
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r <= n_reads; r += gridDim.x * blockDim.x) {
        if (r == n_reads) { ncand[r] = 0; nrep[r] = 0; continue; }
        const uint32_t nc = ncand[r];
        nrep[r] = nc ? nc : 1u;
    }
}

__global__ void k_ctl_cands(BatchCtl *ctl, const int64_t *cand_off, const int64_t *rep_off, int n_reads, long long cap_cands, long long cap_reps)
{
    const long long nc = cand_off[n_reads], nr = rep_off[n_reads];
    ctl->ncand = nc; ctl->nrep = nr;
    if (nc > cap_cands || nr > cap_reps) atomicOr(&ctl->abort, CAP_CANDS);
}

// candidate records of a read from the clustering output
__device__ __forceinline__ void cand_init_read(int r, const int64_t *seed_off, const int64_t *cand_off, const uint32_t *ncand, const int32_t *cbegin,
                                               const int32_t *ccount, const int32_t *cscore, const uint64_t *keys, CandState *cs)
{
    const int64_t so = seed_off[r], co = cand_off[r];
    const int nc = (int)ncand[r];
    for (int k = 0; k < nc; k++) {
        CandState c;
        c.read = r; c.seed_begin = cbegin[so + k]; c.seed_count = ccount[so + k]; c.Score = cscore[so + k];
        c.PairedIdx = -1; c.SJtype = -1;
        uint64_t key = keys[so + c.seed_begin];
        int64_t pd = key_gpos(key) - key_rpos(key);
        c.PosDiff = pd < 0 ? 0 : pd;
        c.pos = 0; c.sv_off = 0; c.cig_off = 0; c.text_off = 0; c.sv_n = 0; c.sv_cap = 0; c.cig_cap = 0; c.cig_n = 0; c.text_len = 0;
        c.AlnScore = 0; c.mis = 0; c.chr = 0; c.n_ext = 0; c.live = 0; c.skip = 0; c.dir = 0; c.pad = 0;
        cs[co + k] = c;
    }
}

// One thread per read (pair): the candidate records, candidate pairing / pruning, then the seed-pool demand of the survivors
// (zeros behind the candidate count: the prefix sum behind this kernel runs over the candidate table's capacity).
// Round 2: k_cand_init, k_pair_prune, k_cand_live and k_set_sv_off were four passes over the 104-byte records.
__global__ void k_pair_prune(int n_units, int paired, const int64_t *seed_off, const int64_t *cand_off, const uint32_t *ncand, const int32_t *cbegin,
                             const int32_t *ccount, const int32_t *cscore, const uint64_t *keys, CandState *cs, uint32_t *cap, int64_t cap_cands,
                             const BatchCtl *ctl)
{
    if (ctl->abort) return;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < n_units; u += gridDim.x * blockDim.x) {
        int64_t a, e;
        if (paired) {
            cand_init_read(2 * u, seed_off, cand_off, ncand, cbegin, ccount, cscore, keys, cs);
            cand_init_read(2 * u + 1, seed_off, cand_off, ncand, cbegin, ccount, cscore, keys, cs);
            a = cand_off[2 * u]; e = cand_off[2 * u + 2];
            const int64_t b = cand_off[2 * u + 1];
            pair_and_prune(cs + a, (int)(b - a), cs + b, (int)(e - b), true);
        } else {
            cand_init_read(u, seed_off, cand_off, ncand, cbegin, ccount, cscore, keys, cs);
            a = cand_off[u]; e = cand_off[u + 1];
            pair_and_prune(cs + a, (int)(e - a), nullptr, 0, false);
        }
        for (int64_t c = a; c < e; c++) {
            const int live = cs[c].Score != 0;
            cs[c].live = (uint8_t)live;
            const int k = live ? seed_capacity(cs[c].seed_count) : 0;
            cs[c].sv_cap = k;
            cap[c] = (uint32_t)k;
        }
    }
    // zeros behind the candidate count, up to the capacity the prefix sum runs over
    for (int64_t c = ctl->ncand + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= cap_cands; c += (int64_t)gridDim.x * blockDim.x) cap[c] = 0u;
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
// The candidate count lives on the device (E.ctl->ncand); `slice_off` = the prefix sum that places this phase's pool slices
// (A: seed pool, D: CIGAR pool).
// Phase D (CIGAR pairs, score, coordinates) needs a slice of the CIGAR pool sized by phase C.  Round 1 placed the slices with
// two helper kernels and a prefix sum over all candidates between C and D; now the slice is claimed from a counter in the
// control block at a point where the whole warp is converged (one atomic per warp).
// Measured and dropped (round 2, config[1]): running D in the same thread right behind C for the candidates that queued no
// alignment.  The merged kernel inherits C's 96 registers / 5 CTAs per SM, and these passes are chains of dependent loads
// whose only remedy is occupancy: phase A-C grew from 0.60 to 1.61 ms while the D pass only shrank from 0.52 to 0.06 ms.
// Also measured and dropped: two-ended queues that group light candidates (a single seed; nothing but simple pairs) and heavy
// ones into different warps, with the pool slices and queue places claimed inside k_pair_prune: phase A 0.59 -> 0.62 ms (the
// queue makes its accesses to the candidate table non-contiguous), phase D 0.60 -> 0.58 ms, k_pair_prune 0.30 -> 0.33 ms.
// Also measured and dropped: building small CIGARs in a shared-memory buffer instead of the candidate's pool slice (the pairs
// are written, re-read, reversed and merged in place): 12 KB more shared memory per CTA and generic-address accesses made the
// D pass 1.9x slower (0.59 -> 1.10 ms).
__device__ __forceinline__ long long warp_claim(long long *counter, int need)
{   // all 32 lanes call this together; returns the start of this lane's `need` slots
    const int lane = threadIdx.x & 31;
    int incl = need;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += t; }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    long long base = 0;
    if (lane == 31 && total > 0) base = (long long)atomicAdd(reinterpret_cast<unsigned long long *>(counter), (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    return base + incl - need;
}

// Candidates that have to wait for a batched launch (8-mer re-seeding, NW of the gap flanks, NW of the non-simple pairs) are
// appended to the queue of the phase that picks them up again, so that the later phase kernels run full warps over exactly
// the candidates they have work for (round-2 ncu of k_phase<3> scanning the whole table with a stage marker: 7.8 of 32 lanes
// active, 26 % of the warp slots occupied).
struct PhaseQueues { uint32_t *q[3]; };          // candidates waiting for phase 1 (B), 2 (C), 3 (D)

// Occupancy: C keeps the most state live (96 registers when unconstrained).  Round 1 ran every kernel that may run C at 5 CTAs
// per SM with 8 staged seeds (64 registers spill and halve the speed); 6 CTAs per SM = 80 registers with 6 staged seeds spills
// 12-32 bytes and is faster where it matters (config[1], 2 M reads: phase A 0.59 -> 0.46 ms, B 0.07 -> 0.10 ms); D alone runs 6 CTAs per SM (80 registers: with the in-place alignment of tiny blocks 8 CTAs spill; measured
// 0.68 ms at 8, 0.59 ms at 7 and 6 on config[1]).
#ifndef PHASE3_CTAS
#define PHASE3_CTAS 6
#endif
#ifndef PHASE012_CTAS
#define PHASE012_CTAS 6
#define PHASE012_STAGE 6
#endif
template <int WHICH> struct PhaseCfg { static constexpr int STAGE = WHICH == 3 ? 5 : PHASE012_STAGE, MIN_CTAS = WHICH == 3 ? PHASE3_CTAS : PHASE012_CTAS; };
template <int WHICH>
__global__ void __launch_bounds__(TPB, PhaseCfg<WHICH>::MIN_CTAS) k_phase(Env E, const int64_t *__restrict__ slice_off, long long cap_cig, PhaseQueues Q)
{
    if (E.ctl->abort) return;
    const int64_t total = WHICH == 0 ? (int64_t)E.ctl->ncand : (int64_t)E.ctl->phase_queue[WHICH - 1];
    constexpr int STAGE_SEEDS = PhaseCfg<WHICH>::STAGE;
    __shared__ RSeed s_slot[TPB * STAGE_SEEDS];
    const int lane = threadIdx.x & 31;
    // the loop condition is the same for the 32 lanes of a warp (the claims below need all of them)
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx - lane < total; idx += (int64_t)gridDim.x * blockDim.x) {
        bool run = idx < total;
        const int64_t cid = !run ? 0 : WHICH == 0 ? idx : (int64_t)Q.q[WHICH == 0 ? 0 : WHICH - 1][idx];
        CandState c{};
        if (run) {
            c = E.cs[cid];
            if (WHICH == 0) c.sv_off = slice_off[cid];
            if (!c.live) {                                                      // pruned candidates only exist in phase A's pass
                E.cs[cid].AlnScore = 0; E.cs[cid].cig_n = 0; E.cs[cid].text_len = 0;
                run = false;
            }
        }
        RSeed *g = E.pool + c.sv_off, *sv = g;
        bool staged = false, want_d = false;
        int next = WHICH + 1;
        if (run) {
            staged = phase_seed_bound(c, WHICH) <= STAGE_SEEDS;
            if (staged) {
                sv = s_slot + threadIdx.x * STAGE_SEEDS;
                if (WHICH != 0) for (int i = 0; i < c.sv_n; i++) sv[i] = g[i];
            }
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
            if (WHICH == 2) { phase_c(E, c, sv); next = 3; }
            if (WHICH == 3) want_d = true;
        }
        if (WHICH == 3) {
            // ---- the whole warp: claim the CIGAR slices ----
            __syncwarp();
            const int need = (want_d && !c.skip) ? c.cig_cap : 0;
            const long long off = warp_claim(&E.ctl->cig_total, need);
            if (want_d) {
                c.cig_off = off;
                if (off + need > cap_cig) { atomicOr(&E.ctl->abort, CAP_CIG); c.cig_cap = 0; }     // the batch is re-run with a bigger pool
                phase_d(E, c, sv);
                next = 4;
            }
        }
        if (run) {
            if (staged) for (int i = 0; i < c.sv_n; i++) g[i] = sv[i];
            E.cs[cid] = c;
        }
        // ---- the whole warp: queue the candidates that wait for a batched launch ----
        __syncwarp();
#pragma unroll
        for (int p = WHICH + 1; p <= 3; p++) {
            const unsigned m = __ballot_sync(0xffffffffu, run && next == p);
            if (m) {
                const int leader = __ffs(m) - 1;
                int base = 0;
                if (lane == leader) base = atomicAdd(&E.ctl->phase_queue[p - 1], __popc(m));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (run && next == p) Q.q[p - 1][base + __popc(m & ((1u << lane) - 1u))] = (uint32_t)cid;
            }
        }
    }
}

// ---- final pass ----
// One thread per read (pair): best / second best, mate rescue, flags, MAPQ, then — at a point where the warp is converged —
// the read's slice of the CIGAR-text pool and of the junction-record pool is claimed from the control block (one atomic per
// warp and pool) and filled at once.  Round 1 / early round 2 measured the text here, ran two prefix sums over the reads and
// wrote text and junction records in a second kernel that re-read every read, report and candidate record (0.28 ms of a
// 4.8 ms config[1] step).  The order of the slices inside the two pools is whatever order the warps arrive in; every report
// carries its own cigar_off and every junction record its read, and no consumer depends on the order.
__global__ void k_read_final(Env E, int n_units, int paired, const int64_t *cand_off, const int64_t *rep_off, dartgpu_read_result *rr,
                             dartgpu_report *rep, char *text, dartgpu_junction *junc, long long cap_text, long long cap_junc)
{
    if (E.ctl->abort) return;
    const int lane = threadIdx.x & 31;
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u - lane < n_units; u += gridDim.x * blockDim.x) {
        const bool run = u < n_units;
        const int nr = paired ? 2 : 1;
        ReadOut ro[2];
        int text_need = 0, junc_need = 0;
        if (run) {
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
                        if (c.cig_n < 0) atomicOr(&E.ctl->err, ERR_CIGAR_POOL);
                        if (c.live && !c.skip && c.AlnScore > 0) tl = c.text_len;
                    }
                    rep[rep_off[r] + k].cigar_len = (int16_t)tl;
                    text_need += tl;
                }
                junc_need += emit_junctions(E, ro[m], E.cs + cand_off[r], nc, r, nullptr);
            }
        }
        // ---- the whole warp: claim the slices ----
        __syncwarp();
        const long long t0 = warp_claim(&E.ctl->text_total, text_need);
        const long long j0 = warp_claim(&E.ctl->junc_total, junc_need);
        if (!run) continue;
        const bool text_ok = t0 + text_need <= cap_text, junc_ok = j0 + junc_need <= cap_junc;
        if (!text_ok) atomicOr(&E.ctl->abort, CAP_TEXT);        // the batch is re-run with bigger pools
        if (!junc_ok) atomicOr(&E.ctl->abort, CAP_JUNC);
        long long at = t0, jat = j0;
        for (int m = 0; m < nr; m++) {
            int r = paired ? 2 * u + m : u;
            const int nc = (int)(cand_off[r + 1] - cand_off[r]);
            for (int k = 0; k < ro[m].n_reports; k++) {
                dartgpu_report &p = rep[rep_off[r] + k];
                p.cigar_off = (int32_t)at;
                if (p.cigar_len > 0 && k < nc && text_ok) {
                    const CandState &c = E.cs[cand_off[r] + k];
                    write_cigar_text(E.cig + c.cig_off, c.cig_n, text + at);
                }
                at += p.cigar_len;
            }
            if (junc_ok && junc_need > 0) jat += emit_junctions(E, ro[m], E.cs + cand_off[r], nc, r, junc + jat);
        }
    }
}

} // namespace

// buffers owned by the device pipeline (kept across calls inside the context)
struct DevicePipe {
    DevBuf<int64_t> cand_off, sv_off, rep_off;
    DevBuf<uint32_t> u32_a, u32_b;
    DevBuf<CandState> cs;
    DevBuf<RSeed> pool;
    DevBuf<uint32_t> queue;                  // candidates waiting for phases B, C, D
    DevBuf<KmerJobDev> kjobs;
    DevBuf<dartgpu_kmer_hit> khits;
    DevBuf<NwJobDev> jobsB, jobsC;
    DevBuf<uint8_t> opsB, opsC;
    DevBuf<int32_t> nopsB, nopsC, aux, cig;
    DevBuf<uint32_t> flags;
    DevBuf<int32_t> rowbuf;
    DevBuf<dartgpu_read_result> rr;
    DevBuf<dartgpu_report> rep;
    DevBuf<char> text;
    DevBuf<dartgpu_junction> junc;
    DevBuf<int64_t> d_chr_fwd; DevBuf<int32_t> d_end_chr;
    DevBuf<uint8_t> scan_tmp;
    PinBuf<dartgpu_read_result> h_rr;
    PinBuf<dartgpu_report> h_rep;
    PinBuf<char> h_text;
    PinBuf<dartgpu_junction> h_junc;
    int64_t sent_rep = 0, sent_text = 0, sent_junc = 0;       // what the enqueued result copies cover
    int64_t last_rep = -1, last_text = -1, last_junc = -1;    // the previous batch's actual sizes: the prediction for the next
    int last_n = 0;
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

// one round of NW over a device-side job queue: everything the round needs is sized by capacities
static void nw_round(dartgpu_ctx *c, DevicePipe *D, int round, NwJobDev *jobs, int cap_jobs, DevBuf<uint8_t> &ops, DevBuf<int32_t> &nops)
{
    cudaStream_t st = c->stream;
    const Caps &K = c->caps;
    NwRound R{};
    R.jobs = jobs; R.n_jobs = &c->d_ctl.p->nw_jobs[round]; R.cap_jobs = cap_jobs; R.round = round; R.with_aux = round == 0;
    R.cap_ops = K.nw_ops[round]; R.cap_flags = K.nw_flags; R.cap_aux = round == 0 ? K.nw_aux : 0;
    nops.reserve(cap_jobs + 1); ops.reserve(R.cap_ops + 1); D->flags.reserve(R.cap_flags + 1);
    if (round == 0) D->aux.reserve(R.cap_aux + 1);
    // widest job a phase can emit: gLen <= max(30, 2 * rGaps) (AlignmentCandidates.cpp:1020), rGaps <= rlen
    const size_t rb = (size_t)2 * (std::max(64, 2 * c->max_rlen + 64) + 1);
    D->rowbuf.reserve(rb * nw_grid_warps() + 1);
    R.flags = D->flags.p; R.ops = ops.p; R.nops = nops.p; R.rowbuf = D->rowbuf.p; R.rowbuf_per_warp = rb;
    DG_CUDA(cudaEventRecord(c->ev[round == 0 ? 10 : 15], st));
    launch_nw(c->ix, c->d_codes.p, R, c->d_ctl.p, c->nwscratch, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[round == 0 ? 11 : 16], st));
    c->stats.kernel_launches += NW_LAUNCHES;
}

// Enqueues the whole device orchestration behind the seeding kernels and the copies of the predicted result sizes.
// Nothing here waits for the GPU.
void enqueue_pipeline(dartgpu_ctx *c)
{
    DevicePipe *D = pipe_of(c);
    cudaStream_t st = c->stream;
    const int n = c->n_reads;
    const dartgpu_params &P = c->prm;
    const Caps &K = c->caps;
    BatchCtl *ctl = c->d_ctl.p;
    const int paired = P.pair_end != 0;
    const int units = paired ? n / 2 : n;
    if (!D->tables) {
        D->d_chr_fwd.reserve(c->shared->chr_fwd.size()); D->d_end_chr.reserve(c->shared->end_chr.size());
        DG_CUDA(cudaMemcpyAsync(D->d_chr_fwd.p, c->shared->chr_fwd.data(), c->shared->chr_fwd.size() * 8, cudaMemcpyHostToDevice, st));
        std::vector<int32_t> ec(c->shared->end_chr.begin(), c->shared->end_chr.end());
        DG_CUDA(cudaMemcpyAsync(D->d_end_chr.p, ec.data(), ec.size() * 4, cudaMemcpyHostToDevice, st));
        DG_CUDA(dg_stream_sync(st));        // once per context: the staging vector above dies here
        D->tables = true;
    }
    DG_CUDA(cudaEventRecord(c->ev[12], st));

    // ---- candidate table: per-read offsets of candidates and reports (prefix sums over the reads) ----
    const int64_t cap_c = K.cands, cap_r = K.cands + n;
    D->cand_off.reserve(n + 2); D->rep_off.reserve(n + 2); D->u32_b.reserve(n + 2);
    k_report_counts<<<grid_for(n + 1), TPB, 0, st>>>(n, c->d_ncand.p, D->u32_b.p);
    scan_u32(c, D, c->d_ncand.p, D->cand_off.p, n);
    scan_u32(c, D, D->u32_b.p, D->rep_off.p, n);
    k_ctl_cands<<<1, 1, 0, st>>>(ctl, D->cand_off.p, D->rep_off.p, n, cap_c, cap_r);
    D->cs.reserve(cap_c + 1);
    // ---- pairing / pruning; seed-pool slices of the survivors (prefix sum over the table's capacity, zeros behind the count) ----
    D->u32_a.reserve(cap_c + 2); D->sv_off.reserve(cap_c + 2);
    k_pair_prune<<<grid_for(units), TPB, 0, st>>>(units, paired, c->d_seed_off.p, D->cand_off.p, c->d_ncand.p, c->d_cand_begin.p, c->d_cand_count.p,
                                                  c->d_cand_score.p, c->d_keys.p, D->cs.p, D->u32_a.p, cap_c, ctl);
    scan_u32(c, D, D->u32_a.p, D->sv_off.p, cap_c);
    launch_ctl_check(ctl, &ctl->pool_total, D->sv_off.p + cap_c, K.pool, CAP_POOL, st);
    D->pool.reserve(K.pool + 1);
    // job queues: bounded by the pool (a k-mer job needs a gap between two seeds of a >= 2-seed candidate: 28 pool slots; ...)
    const int cap_k = (int)std::min<int64_t>(K.pool / 12 + 2, INT32_MAX / 2), cap_B = (int)std::min<int64_t>(K.pool / 3 + 4, INT32_MAX / 2),
              cap_C = (int)std::min<int64_t>(K.pool + 4, INT32_MAX / 2);
    D->kjobs.reserve(cap_k); D->khits.reserve(cap_k); D->jobsB.reserve(cap_B); D->jobsC.reserve(cap_C);
    c->stats.kernel_launches += 9;

    Env E{};
    E.P = PhaseParams{P.max_gaps, P.max_intron, P.min_intron, P.max_mismatch, P.multi_hit, P.pair_end, P.all_sj};
    E.ref = RefView{c->ix.ref2, nullptr, c->G}; E.G = c->G;
    E.ends = c->shared->d_ends.p; E.end_chr = D->d_end_chr.p; E.n_ends = (int)c->shared->ends.size(); E.chr_fwd = D->d_chr_fwd.p;
    E.codes = c->d_codes.p; E.code_off = c->d_dev_off.p; E.rlen = c->d_rlen.p; E.keys = c->d_keys.p; E.seed_off = c->d_seed_off.p;
    E.ctl = ctl;
    E.cs = D->cs.p; E.pool = D->pool.p;
    E.kjobs = D->kjobs.p; E.kjob_count = &ctl->nk; E.khits = D->khits.p;
    E.njobs = D->jobsB.p; E.njob_count = &ctl->nw_jobs[0];          // phase B's queue (also used by candidates fast-tracked out of A)
    E.njobs_c = D->jobsC.p; E.njob_count_c = &ctl->nw_jobs[1];      // phase C's queue
    D->queue.reserve(3 * (size_t)(cap_c + 1));
    PhaseQueues Q{{D->queue.p, D->queue.p + (cap_c + 1), D->queue.p + 2 * (cap_c + 1)}};
    D->cig.reserve(K.cig + 1);
    E.cig = D->cig.p;

    // ---- phase A -> 8-mer re-seeding (candidates with nothing to repair run B, C and D in the same pass) ----
    k_phase<0><<<grid_for(cap_c), TPB, 0, st>>>(E, D->sv_off.p, K.cig, Q);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[8], st));
    launch_kmer(c->ix, c->d_codes.p, D->kjobs.p, &ctl->nk, cap_k, std::max(c->max_rlen, 8), D->khits.p, c->kscratch, ctl, K.krecs, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[9], st));
    c->stats.kernel_launches += 1 + KMER_LAUNCHES;

    // ---- phase B -> NW of every gap against both flanks ----
    k_phase<1><<<grid_for(cap_c), TPB, 0, st>>>(E, nullptr, K.cig, Q);
    DG_CUDA(cudaGetLastError());
    nw_round(c, D, 0, D->jobsB.p, cap_B, D->opsB, D->nopsB);

    // ---- phase C -> NW of every non-simple pair ----
    E.ops = D->opsB.p; E.nops = D->nopsB.p; E.done_jobs = D->jobsB.p; E.xscratch = D->aux.p;
    E.njobs = D->jobsC.p; E.njob_count = &ctl->nw_jobs[1];
    k_phase<2><<<grid_for(cap_c), TPB, 0, st>>>(E, nullptr, K.cig, Q);
    DG_CUDA(cudaGetLastError());
    nw_round(c, D, 1, D->jobsC.p, cap_C, D->opsC, D->nopsC);

    // ---- phase D for the candidates that waited for an alignment: CIGAR pairs, score, coordinates ----
    E.ops = D->opsC.p; E.nops = D->nopsC.p; E.done_jobs = D->jobsC.p;
    k_phase<3><<<grid_for(cap_c), TPB, 0, st>>>(E, nullptr, K.cig, Q);
    DG_CUDA(cudaGetLastError());

    // ---- per read / pair: best, mate rescue, flags, MAPQ; record layout (prefix sums over the reads) ----
    D->rr.reserve(n + 1); D->rep.reserve(cap_r + 1); D->text.reserve(K.text + 1); D->junc.reserve(K.junc + 1);
    k_read_final<<<grid_for(units), TPB, 0, st>>>(E, units, paired, D->cand_off.p, D->rep_off.p, D->rr.p, D->rep.p, D->text.p, D->junc.p,
                                                  std::min<int64_t>(K.text, (1ll << 31) - 1), K.junc);
    DG_CUDA(cudaGetLastError());
    c->stats.kernel_launches += 4;
    auto predict = [&](int64_t last, int64_t cap, int64_t first_guess) {
        int64_t want = last < 0 ? first_guess : (int64_t)((double)last * (double)n / std::max(1, D->last_n) * 1.02) + 4096;
        return std::max<int64_t>(0, std::min(want, cap));
    };
    if (c->emit_sam) {
        // ---- complete SAM lines on the device (sam_kernels.cu): only text and junction records cross PCIe ----
        enqueue_sam(c, D->rr.p, D->rep.p, D->text.p);
        DG_CUDA(cudaEventRecord(c->ev[13], st));
        compute_turn_end(c);
        FastqDev &F = c->fq;
        const int64_t guess = F.last_sam < 0 ? std::min<int64_t>(K.sam, (int64_t)n * (2 * std::max(c->max_rlen, 32) + 96)) : 0;
        F.sent_sam = predict(F.last_sam, K.sam, guess);
        D->sent_junc = predict(D->last_junc, K.junc, n / 8 + 4096);
        D->sent_rep = D->sent_text = 0;
        F.h_sam.reserve(F.sent_sam + 1); D->h_junc.reserve(D->sent_junc + 1);
        if (F.sent_sam) DG_CUDA(cudaMemcpyAsync(F.h_sam.p, F.sam.p, (size_t)F.sent_sam, cudaMemcpyDeviceToHost, st));
        if (D->sent_junc) DG_CUDA(cudaMemcpyAsync(D->h_junc.p, D->junc.p, (size_t)D->sent_junc * sizeof(dartgpu_junction), cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaEventRecord(c->ev[14], st));
        return;
    }
    DG_CUDA(cudaEventRecord(c->ev[13], st));
    compute_turn_end(c);                     // the next batch's kernels (any context of this device) may start: the copies below run under them

    if (c->results_on_device) { D->sent_rep = D->sent_text = D->sent_junc = 0; DG_CUDA(cudaEventRecord(c->ev[14], st)); return; }
    // ---- only the final records cross PCIe.  Their sizes are known on the device only: the copies cover what the previous
    // batch of this context needed plus a margin (a guess the first time); finish_pipeline tops up.
    D->sent_rep = predict(D->last_rep, cap_r, (int64_t)n + n / 2 + 1024);
    D->sent_text = predict(D->last_text, K.text, 6ll * n + 4096);
    D->sent_junc = predict(D->last_junc, K.junc, n / 8 + 4096);
    D->h_rr.reserve(n + 1); D->h_rep.reserve(D->sent_rep + 1); D->h_text.reserve(D->sent_text + 1); D->h_junc.reserve(D->sent_junc + 1);
    DG_CUDA(cudaMemcpyAsync(D->h_rr.p, D->rr.p, (size_t)n * sizeof(dartgpu_read_result), cudaMemcpyDeviceToHost, st));
    if (D->sent_rep) DG_CUDA(cudaMemcpyAsync(D->h_rep.p, D->rep.p, (size_t)D->sent_rep * sizeof(dartgpu_report), cudaMemcpyDeviceToHost, st));
    if (D->sent_text) DG_CUDA(cudaMemcpyAsync(D->h_text.p, D->text.p, (size_t)D->sent_text, cudaMemcpyDeviceToHost, st));
    if (D->sent_junc) DG_CUDA(cudaMemcpyAsync(D->h_junc.p, D->junc.p, (size_t)D->sent_junc * sizeof(dartgpu_junction), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaEventRecord(c->ev[14], st));
}

// After the batch's one synchronisation (c->h_ctl holds the control block, no abort): top up the result copies when the
// prediction fell short (first batch of a context, or a batch unlike the previous one), then publish.
void finish_pipeline(dartgpu_ctx *c, dartgpu_map_result *out)
{
    DevicePipe *D = pipe_of(c);
    cudaStream_t st = c->stream;
    const int n = c->n_reads;
    const BatchCtl &H = c->h_ctl.p[0];
    const int64_t nrep = H.nrep, text_total = H.text_total, junc_total = H.junc_total;
    if (c->results_on_device) {          // dartgpu_set_result_location(ctx, 1): the records stay in HBM, the pointers are device pointers
        out->reads = D->rr.p; out->n_reads = n;
        out->reports = D->rep.p; out->n_reports = nrep;
        out->cigars = D->text.p; out->n_cigar_bytes = text_total;
        out->junctions = D->junc.p; out->n_junctions = junc_total;
        return;
    }
    bool more = false;
    auto top_up = [&](auto &hbuf, const auto &dbuf, int64_t sent, int64_t total, size_t elem) {
        if (total <= sent) return;
        if ((size_t)total + 1 > hbuf.cap) {          // the pinned buffer itself is too small: a fresh one, everything again
            hbuf.reserve(total + 1);
            sent = 0;
        }
        DG_CUDA(cudaMemcpyAsync((char *)hbuf.p + sent * elem, (const char *)dbuf.p + sent * elem, (size_t)(total - sent) * elem, cudaMemcpyDeviceToHost, st));
        more = true;
    };
    top_up(D->h_rep, D->rep, D->sent_rep, nrep, sizeof(dartgpu_report));
    top_up(D->h_text, D->text, D->sent_text, text_total, 1);
    top_up(D->h_junc, D->junc, D->sent_junc, junc_total, sizeof(dartgpu_junction));
    if (more) DG_CUDA(dg_stream_sync(st));
    D->last_rep = nrep; D->last_text = text_total; D->last_junc = junc_total; D->last_n = n;
    c->stats.d2h_bytes += (uint64_t)n * sizeof(dartgpu_read_result) + (uint64_t)std::max(nrep, D->sent_rep) * sizeof(dartgpu_report) +
                          std::max(text_total, D->sent_text) + std::max(junc_total, D->sent_junc) * sizeof(dartgpu_junction);
    out->reads = D->h_rr.p; out->n_reads = n;
    out->reports = D->h_rep.p; out->n_reports = nrep;
    out->cigars = D->h_text.p; out->n_cigar_bytes = text_total;
    out->junctions = D->h_junc.p; out->n_junctions = junc_total;
}

// dartgpu_wait_sam: the same for a batch whose results are SAM text
void finish_sam(dartgpu_ctx *c, dartgpu_sam_result *out)
{
    DevicePipe *D = pipe_of(c);
    FastqDev &F = c->fq;
    cudaStream_t st = c->stream;
    const BatchCtl &H = c->h_ctl.p[0];
    const int64_t sam_total = H.sam_bytes, junc_total = H.junc_total;
    bool more = false;
    if (sam_total > F.sent_sam) {
        int64_t sent = F.sent_sam;
        if ((size_t)sam_total + 1 > F.h_sam.cap) { F.h_sam.reserve(sam_total + 1); sent = 0; }
        DG_CUDA(cudaMemcpyAsync(F.h_sam.p + sent, F.sam.p + sent, (size_t)(sam_total - sent), cudaMemcpyDeviceToHost, st));
        more = true;
    }
    if (junc_total > D->sent_junc) {
        int64_t sent = D->sent_junc;
        if ((size_t)junc_total + 1 > D->h_junc.cap) { D->h_junc.reserve(junc_total + 1); sent = 0; }
        DG_CUDA(cudaMemcpyAsync(D->h_junc.p + sent, D->junc.p + sent, (size_t)(junc_total - sent) * sizeof(dartgpu_junction), cudaMemcpyDeviceToHost, st));
        more = true;
    }
    if (more) DG_CUDA(dg_stream_sync(st));
    F.last_sam = sam_total; D->last_junc = junc_total; D->last_n = c->n_reads;
    c->stats.d2h_bytes += (uint64_t)std::max(sam_total, F.sent_sam) + (uint64_t)std::max(junc_total, D->sent_junc) * sizeof(dartgpu_junction);
    out->sam = F.h_sam.p; out->n_bytes = sam_total;
    out->n_reads = c->n_reads; out->n_unmapped = (int64_t)H.sam_counts[0]; out->n_unique = (int64_t)H.sam_counts[1]; out->n_paired = (int64_t)H.sam_counts[2];
    out->junctions = D->h_junc.p; out->n_junctions = junc_total;
}

} // namespace dartgpu
