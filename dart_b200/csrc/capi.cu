// C-ABI glue (include/dartgpu.h): context life cycle, index hand-over, stage runners.
// There is deliberately no CPU implementation behind any entry point: without a CUDA device
// dartgpu_create* fails with DARTGPU_ERR_NO_DEVICE and nothing else can be called.
#include <omp.h>
#include <sched.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <memory>
#include <mutex>
#include <new>

#include "context.h"
#include "rank.cuh"

using namespace dartgpu;

namespace dartgpu {

static uint8_t *make_code_table()
{
    static uint8_t t[256];
    for (int i = 0; i < 256; i++) t[i] = CODE_OTHER;
    t['A'] = 0; t['C'] = 1; t['G'] = 2; t['T'] = 3;
    t['a'] = 8; t['c'] = 9; t['g'] = 10; t['t'] = 11;   // same base codes, bit 3 = lower case (raw-character compares differ)
    t['N'] = CODE_N; // only the upper-case literal breaks an 8-mer (/root/reference/src/KmerAnalysis.cpp:44)
    return t;
}
static const uint8_t *g_code_table = make_code_table();
static inline uint8_t code_of(unsigned char ch) { return g_code_table[ch]; }

static thread_local std::string g_create_error;

static int fail(dartgpu_ctx *c, int code, const std::string &msg)
{
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

template <class F> static int guarded(dartgpu_ctx *c, F &&f)
{
    try {
        if (c) { cudaError_t e = cudaSetDevice(c->device); if (e != cudaSuccess) throw CudaError{e, "cudaSetDevice", __FILE__, __LINE__}; }
        f();
        return DARTGPU_OK;
    } catch (const CudaError &e) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e.e, cudaGetErrorString(e.e), e.file, e.line, e.what);
        cudaGetLastError();
        return fail(c, e.e == cudaErrorNoDevice || e.e == cudaErrorInsufficientDriver ? DARTGPU_ERR_NO_DEVICE : DARTGPU_ERR_CUDA, buf);
    } catch (const std::bad_alloc &) {
        return fail(c, DARTGPU_ERR_NOMEM, "out of host memory");
    } catch (const std::pair<int, std::string> &e) {
        return fail(c, e.first, e.second);
    } catch (const std::exception &e) {
        return fail(c, DARTGPU_ERR_ARG, e.what());
    }
}

void stats_begin(dartgpu_ctx *c)
{
    c->stats = dartgpu_stats{};
}

// a fresh control block for the batch (or for its next attempt)
static void ctl_begin(dartgpu_ctx *c)
{
    launch_zero(c->d_ctl.p, BATCHCTL_RESET_BYTES, c->stream);      // the ingest fields behind it belong to the upload, not to the attempt
}
static void ctl_fetch(dartgpu_ctx *c)      // enqueue the read-back of the control block (a one-warp kernel, not the copy engine)
{
    static_assert(sizeof(BatchCtl) % 4 == 0, "BatchCtl is copied word by word");
    small_d2h(c->h_ctl.p, c->d_ctl.p, sizeof(BatchCtl), c->stream);
}

// ---------------------------------------------------------------------------------------------------
// pool capacities (see BatchCtl): first guess from the batch's size, growth from an aborted attempt's counts
// ---------------------------------------------------------------------------------------------------
static void caps_for_batch(dartgpu_ctx *c)
{
    Caps &K = c->caps;
    const int64_t n = c->n_reads, L = std::max(c->max_rlen, 32);
    // test hook: DARTGPU_CAP_SHRINK=<d> divides every first guess by d, so that small batches overflow and take the retry path
    static const int64_t shrink = [] { const char *e = getenv("DARTGPU_CAP_SHRINK"); return e ? std::max<int64_t>(1, atoll(e)) : (int64_t)1; }();
    auto atleast = [](int64_t &v, int64_t want) { want = std::max<int64_t>(want / shrink, 16); if (v < want) v = want; };
    atleast(K.seeds, 3 * n + 4096);
    atleast(K.cands, 2 * n + 4096);
    atleast(K.pool, 18 * n + 4096);
    atleast(K.krecs, n / 2 + 65536);
    atleast(K.cig, 18 * n + 4096);
    atleast(K.text, 8 * n + 4096);
    atleast(K.junc, n / 4 + 4096);
    atleast(K.nw_ops[0], n * L / 16 + 65536);
    atleast(K.nw_ops[1], n * L / 8 + 65536);
    atleast(K.nw_flags, n * L / 8 + 65536);
    atleast(K.nw_aux, n * L / 32 + 65536);
    if (c->emit_sam) atleast(K.sam, n * (2 * L + 128));
}

// Grows whatever the aborted attempt is known to need (+25 %).  An attempt stops at the first pool that overflows, so the
// pools behind it are scaled by the same factor: a batch ten times the expected seed count needs ten times the rest too.
static void caps_grow(dartgpu_ctx *c, const BatchCtl &H)
{
    Caps &K = c->caps;
    double f = 1.0;
    auto need = [&](int64_t &cap, long long have) {
        if (have > cap) { f = std::max(f, (double)have / (double)std::max<int64_t>(cap, 1)); cap = have + have / 4 + 4096; return true; }
        return false;
    };
    const bool seeds = need(K.seeds, H.total_seeds);
    const bool cands = need(K.cands, std::max(H.ncand, H.nrep - c->n_reads));
    const bool pool = need(K.pool, H.pool_total);
    const bool krecs = need(K.krecs, H.kmer_recs);
    const bool nwb = need(K.nw_ops[0], (long long)H.nw_ops[0]);
    const bool nwc = need(K.nw_ops[1], (long long)H.nw_ops[1]);
    const bool nwf = need(K.nw_flags, (long long)std::max(H.nw_flags[0], H.nw_flags[1]));
    const bool nwa = need(K.nw_aux, (long long)H.nw_aux[0]);
    const bool cig = need(K.cig, H.cig_total);
    const bool text = need(K.text, H.text_total);
    const bool junc = need(K.junc, H.junc_total);
    need(K.sam, H.sam_bytes);
    if (H.abort & CAP_RLEN) {          // a read longer than the bound the scratch was sized for (FASTQ ingest): take the real maximum
        c->fq.rlen_seen = std::max(c->fq.rlen_seen, (int)H.ingest_max_rlen);
        c->max_rlen = c->fq.rlen_seen;
        c->cap_rec = std::max(1, (c->max_rlen + 15) / 16);
    }
    f = std::min(f, 64.0);
    auto scale = [&](int64_t &cap) { cap = (int64_t)((double)cap * f * 1.1) + 4096; };
    // everything downstream of the first overflow has not been counted yet
    if (seeds) { if (!cands) scale(K.cands); if (!pool) scale(K.pool); }
    if (seeds || cands || pool) {
        if (!krecs) scale(K.krecs); if (!nwb) scale(K.nw_ops[0]); if (!nwc) scale(K.nw_ops[1]); if (!nwf) scale(K.nw_flags);
        if (!nwa) scale(K.nw_aux); if (!cig) scale(K.cig); if (!text) scale(K.text); if (!junc) scale(K.junc);
    }
    (void)text; (void)junc;
}

static void check_ctl_errors(const BatchCtl &H)
{
    if (H.err & ERR_CIGAR_POOL) throw std::make_pair(DARTGPU_ERR_CUDA, std::string("CIGAR pool capacity exceeded"));
    if (H.err & ERR_SORT_SCRATCH) throw std::make_pair(DARTGPU_ERR_CUDA, std::string("seed sort scratch overflow"));
    if (H.err & ERR_NW_WIDTH) throw std::make_pair(DARTGPU_ERR_CUDA, std::string("NW job wider than the row buffer"));
    if (H.ingest_err & ERR_FASTQ_LINES) throw std::make_pair(DARTGPU_ERR_ARG, std::string("FASTQ block: line count does not match 4 x n_records"));
    if (H.ingest_err & ERR_READ_TOO_LONG) throw std::make_pair(DARTGPU_ERR_READ_TOO_LONG, std::string("a read is longer than DARTGPU_MAX_RLEN"));
}

void add_ms(dartgpu_ctx *c, double *slot, cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0;
    if (cudaEventElapsedTime(&ms, a, b) == cudaSuccess) *slot += ms;
    (void)c;
}

// ---------------------------------------------------------------------------------------------------
// read batch: encode + upload.  Every read starts on a 16-byte boundary so a lane stages 16 bases per load.
// ---------------------------------------------------------------------------------------------------
void upload_reads(dartgpu_ctx *c, const dartgpu_reads *reads)
{
    const int n = reads->n_reads;
    if (n < 0 || (n > 0 && (!reads->bases || !reads->offsets))) throw std::make_pair(DARTGPU_ERR_ARG, std::string("bad read batch"));
    c->from_fastq = false;
    launch_zero(&c->d_ctl.p->ingest_err, sizeof(BatchCtl) - BATCHCTL_RESET_BYTES, c->stream);
    const int threads = c->prm.host_threads > 0 ? c->prm.host_threads : omp_get_max_threads();
    int max_rlen = 0, bad = 0;
#pragma omp parallel for schedule(static) num_threads(threads) reduction(max : max_rlen) reduction(| : bad)
    for (int i = 0; i < n; i++) {
        int64_t rl = reads->offsets[i + 1] - reads->offsets[i];
        if (rl < 0) bad |= 1;
        else if (rl > DARTGPU_MAX_RLEN) bad |= 2;
        else if ((int)rl > max_rlen) max_rlen = (int)rl;
    }
    if (bad & 1) throw std::make_pair(DARTGPU_ERR_ARG, std::string("read offsets are not monotone"));
    if (bad & 2) throw std::make_pair(DARTGPU_ERR_READ_TOO_LONG, std::string("a read is longer than DARTGPU_MAX_RLEN"));
    const int64_t first = n ? reads->offsets[0] : 0, n_bases = n ? reads->offsets[n] - first : 0;
    c->h_raw.reserve(n_bases + 16); c->h_off.reserve(n + 2);
    c->n_reads = n; c->max_rlen = max_rlen;
    c->cap_rec = std::max(1, (max_rlen + 15) / 16);
    c->d_raw.reserve(n_bases + 32); c->d_off.reserve(n + 2); c->d_padded.reserve(n + 2);
    c->d_dev_off.reserve(n + 2); c->d_rlen.reserve(n + 1);
    const int64_t code_bytes = n_bases + 15ll * n + 16;      // upper bound of the padded layout
    c->d_codes.reserve(code_bytes);
    c->d_packed.reserve(code_bytes / 16 + 4);
    c->n_code_bytes = code_bytes;
    cudaStream_t st = c->stream;
    DG_CUDA(cudaEventRecord(c->ev[0], st));
    if (n) {
        // Caller buffers that are already page-locked (cudaHostAlloc / cudaHostRegister) are handed to the DMA engine as they
        // are.  Pageable ones are staged through pinned memory in slices, every core copying, each slice handed over as soon
        // as it is staged so the host copy of slice k+1 overlaps the transfer of k.
        auto pinned = [](const void *p) {
            cudaPointerAttributes at{};
            if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
            return at.type == cudaMemoryTypeHost;
        };
        if (pinned(reads->offsets))
            DG_CUDA(cudaMemcpyAsync(c->d_off.p, reads->offsets, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        else {
            memcpy(c->h_off.p, reads->offsets, (size_t)(n + 1) * sizeof(int64_t));
            DG_CUDA(cudaMemcpyAsync(c->d_off.p, c->h_off.p, (size_t)(n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        }
        if (pinned(reads->bases + first))
            DG_CUDA(cudaMemcpyAsync(c->d_raw.p, reads->bases + first, (size_t)n_bases, cudaMemcpyHostToDevice, st));
        else {
            const int64_t slice = 16 << 20, piece = 1 << 20;
            for (int64_t s0 = 0; s0 < n_bases; s0 += slice) {
                const int64_t len = std::min(slice, n_bases - s0), npieces = (len + piece - 1) / piece;
#pragma omp parallel for schedule(static) num_threads(threads)
                for (int64_t k = 0; k < npieces; k++)
                    memcpy(c->h_raw.p + s0 + k * piece, reads->bases + first + s0 + k * piece, (size_t)std::min(piece, len - k * piece));
                DG_CUDA(cudaMemcpyAsync(c->d_raw.p + s0, c->h_raw.p + s0, (size_t)len, cudaMemcpyHostToDevice, st));
            }
        }
        launch_read_layout(c->d_off.p, n, c->d_rlen.p, c->d_padded.p, st);
        size_t tmp = scan_tmp_bytes(n);
        c->d_scan_tmp.reserve(tmp + 256);
        launch_scan_u32_to_i64(c->d_padded.p, c->d_dev_off.p, n, c->d_scan_tmp.p, tmp, st);
        launch_encode_reads(c->d_raw.p, c->d_off.p, c->d_dev_off.p, n, c->d_codes.p, c->d_packed.p, st);
        DG_CUDA(cudaGetLastError());
        c->stats.kernel_launches += 4;
    }
    DG_CUDA(cudaEventRecord(c->ev[1], st));
    c->stats.h2d_bytes += n_bases + (uint64_t)(n + 1) * 8;
    c->stats.read_bases = (uint64_t)n_bases;
}

// ---------------------------------------------------------------------------------------------------
// stage 1
// ---------------------------------------------------------------------------------------------------
void enqueue_seeding(dartgpu_ctx *c)
{
    const int n = c->n_reads;
    cudaStream_t st = c->stream;
    const Caps &K = c->caps;
    if (K.seeds >= (1ll << 31)) throw std::make_pair(DARTGPU_ERR_ARG, std::string("batch too large: more than 2^31 seeds, split the batch"));
    c->d_recs.reserve((size_t)n * c->cap_rec);
    c->d_nrec.reserve(n + 1); c->d_nhits.reserve(n + 1); c->d_ncand.reserve(n + 2);
    c->d_seed_off.reserve(n + 2);
    c->d_big_list.reserve(n + 1); c->d_mid_list.reserve(n + 1);
    size_t tmp = scan_tmp_bytes(n);
    c->d_scan_tmp.reserve(tmp + 256);
    c->d_keys.reserve(K.seeds + 1); c->d_meta.reserve(K.seeds + 1);
    c->d_cand_begin.reserve(K.seeds + 1); c->d_cand_count.reserve(K.seeds + 1); c->d_cand_score.reserve(K.seeds + 1);
    // reads with more seeds than fit the shared-memory sort need global scratch: bound = cap_rec * max_dup
    size_t bound = (size_t)c->cap_rec * c->prm.max_dup, per_cta = 0;
    if (bound > 4096) { per_cta = 64; while (per_cta < bound) per_cta <<= 1; c->d_big_scratch.reserve(per_cta * sm_count()); }

    SeedLaunch a{};
    a.codes = c->d_codes.p; a.dev_off = c->d_dev_off.p; a.rlen = c->d_rlen.p; a.n_reads = n;
    a.cap_rec = c->cap_rec; a.max_dup = c->prm.max_dup; a.max_gaps = c->prm.max_gaps; a.max_intron = c->prm.max_intron;
    a.recs = c->d_recs.p; a.nrec = c->d_nrec.p; a.nhits = c->d_nhits.p; a.seed_off = c->d_seed_off.p;
    a.ncand = c->d_ncand.p; a.big_list = c->d_big_list.p; a.big_count = &c->d_ctl.p->big_count; a.mid_list = c->d_mid_list.p; a.mid_count = &c->d_ctl.p->mid_count;
    static const bool always_count = [] { const char *e = getenv("DARTGPU_STATS"); return e && atoi(e) != 0; }();
    a.count_work = (!c->whole_path_enqueue || always_count) ? 1 : 0;
    a.ctl = c->d_ctl.p; a.stats = &c->d_ctl.p->stats; a.packed = c->d_packed.p; a.steal = &c->d_ctl.p->steal;
    a.keys = c->d_keys.p; a.meta = c->d_meta.p;
    a.cand_begin = c->d_cand_begin.p; a.cand_count = c->d_cand_count.p; a.cand_score = c->d_cand_score.p;
    a.big_scratch = c->d_big_scratch.p; a.big_scratch_per_cta = per_cta;

    DG_CUDA(cudaEventRecord(c->ev[2], st));
    launch_search(c->ix, a, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[3], st));
    launch_scan_hits(a, c->d_scan_tmp.p, tmp, st);
    launch_ctl_check(c->d_ctl.p, &c->d_ctl.p->total_seeds, c->d_seed_off.p + n, K.seeds, CAP_SEEDS, st);
    DG_CUDA(cudaEventRecord(c->ev[4], st));
    launch_expand_locate(c->ix, a, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[5], st));
    launch_sort_cluster(c->ix, a, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[6], st));
    c->stats.kernel_launches += 9;
}

// Several contexts of one device keep several batches in flight so that the copies of one batch overlap the kernels of
// another.  Left alone, the hardware time-shares the SMs between the streams, the batches advance in lockstep, finish their
// kernels together and then queue their result copies on the one D2H engine with nothing left to overlap them (measured on
// config[1], 4 contexts: 7.2 ms per step = kernels + copies, exactly what ONE context needs).  So kernels take turns: the
// first kernel of the batch submitted as number i on a device waits for the event behind the last kernel of batch i - T
// (any context of the device).  T = 1 is strict FIFO: no two batches compute at once, copies run under the next batch's
// kernels.  T = 2 (default) lets two batches share the SMs, half a batch out of phase by construction, so that one fills the
// launch gaps and small-kernel tails of the other while the convoy still cannot form.  Uploads are not part of a turn.
// DARTGPU_TURNS=0 switches the mechanism off.
static const int g_turns = [] { const char *e = getenv("DARTGPU_TURNS"); int v = e ? atoi(e) : 2; return std::max(0, std::min(v, SharedIndex::TURN_RING - 1)); }();
void compute_turn_begin(dartgpu_ctx *c)
{
    if (!g_turns) return;
    SharedIndex &S = *c->shared;
    std::lock_guard<std::mutex> lock(S.turn_mutex);
    if (S.turn_next >= (uint64_t)g_turns) {
        const int slot = (int)((S.turn_next - g_turns) % SharedIndex::TURN_RING);
        if (S.turn_ev[slot] && S.turn_owner[slot] != c) DG_CUDA(cudaStreamWaitEvent(c->stream, S.turn_ev[slot], 0));
    }
}
void compute_turn_end(dartgpu_ctx *c)
{
    if (!g_turns) return;
    SharedIndex &S = *c->shared;
    std::lock_guard<std::mutex> lock(S.turn_mutex);
    DG_CUDA(cudaEventRecord(c->compute_done, c->stream));
    const int slot = (int)(S.turn_next % SharedIndex::TURN_RING);
    S.turn_ev[slot] = c->compute_done; S.turn_owner[slot] = c;
    S.turn_next++;
}

// stats of the attempt that went through, from its events and its control block
static void collect_stats(dartgpu_ctx *c, bool whole_path)
{
    const BatchCtl &H = c->h_ctl.p[0];
    if (c->timed_upload) add_ms(c, &c->stats.ms_h2d, c->ev[0], c->ev[1]);
    add_ms(c, &c->stats.ms_search, c->ev[2], c->ev[3]);
    add_ms(c, &c->stats.ms_locate, c->ev[4], c->ev[5]);
    add_ms(c, &c->stats.ms_sort_cluster, c->ev[5], c->ev[6]);
    c->stats.ext_steps = H.stats.ext_steps; c->stats.ext_blocks = H.stats.ext_blocks; c->stats.lf_steps = H.stats.lf_steps;
    c->stats.hits = H.stats.hits; c->stats.seeds = H.stats.seeds; c->stats.search_sector_loads = H.stats.sector_loads;
    c->total_seeds = H.total_seeds;
    if (whole_path) {
        add_ms(c, &c->stats.ms_kmer, c->ev[8], c->ev[9]);
        add_ms(c, &c->stats.ms_nw, c->ev[10], c->ev[11]);
        add_ms(c, &c->stats.ms_nw, c->ev[15], c->ev[16]);
        add_ms(c, &c->stats.ms_report, c->ev[12], c->ev[13]);
        add_ms(c, &c->stats.ms_d2h, c->ev[13], c->ev[14]);
        c->stats.ms_report -= c->stats.ms_kmer + c->stats.ms_nw;   // the phase kernels alone
        add_ms(c, &c->stats.ms_total_device, c->ev[c->timed_upload ? 0 : 2], c->ev[14]);
        c->stats.nw_jobs = (uint64_t)H.nw_jobs[0] + (uint64_t)H.nw_jobs[1]; c->stats.nw_cells = H.work[0];
        c->stats.kmer_jobs = (uint64_t)H.nk; c->stats.kmer_window_bases = H.work[1]; c->stats.kmer_read_bases = H.work[2];
    } else add_ms(c, &c->stats.ms_total_device, c->ev[2], c->ev[7]);
}

// stage 1 alone, synchronous (the stage entry points and the kernel-only timing of bench.py)
static void run_seeding_sync(dartgpu_ctx *c, bool fetch)
{
    const int n = c->n_reads;
    cudaStream_t st = c->stream;
    c->total_seeds = 0;
    if (n == 0) { c->o_cand_off.assign(1, 0); c->h_seed_off.reserve(1); c->h_seed_off.p[0] = 0; return; }
    caps_for_batch(c);
    c->whole_path_enqueue = false;               // the stage entry points keep the work counters
    for (int attempt = 0;; attempt++) {
        ctl_begin(c);
        enqueue_seeding(c);
        ctl_fetch(c);
        DG_CUDA(cudaEventRecord(c->ev[7], st));
        DG_CUDA(dg_stream_sync(st));
        const BatchCtl &H = c->h_ctl.p[0];
        if (!H.abort) break;
        if (attempt >= 8) throw std::make_pair(DARTGPU_ERR_NOMEM, std::string("seed pool keeps overflowing"));
        caps_grow(c, H);
    }
    check_ctl_errors(c->h_ctl.p[0]);
    c->whole_path = false;
    collect_stats(c, false);
    const int64_t total = c->total_seeds;
    if (fetch) {
        c->h_seed_off.reserve(n + 1); c->h_ncand.reserve(n + 1);
        c->h_keys.reserve(total + 1);
        c->h_cand_begin.reserve(total + 1); c->h_cand_count.reserve(total + 1); c->h_cand_score.reserve(total + 1);
        DG_CUDA(cudaMemcpyAsync(c->h_seed_off.p, c->d_seed_off.p, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaMemcpyAsync(c->h_ncand.p, c->d_ncand.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        if (total) {
            DG_CUDA(cudaMemcpyAsync(c->h_keys.p, c->d_keys.p, total * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
            DG_CUDA(cudaMemcpyAsync(c->h_cand_begin.p, c->d_cand_begin.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            DG_CUDA(cudaMemcpyAsync(c->h_cand_count.p, c->d_cand_count.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
            DG_CUDA(cudaMemcpyAsync(c->h_cand_score.p, c->d_cand_score.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        }
        c->stats.d2h_bytes += (n + 1) * 8 + n * 4 + total * 20;
        DG_CUDA(dg_stream_sync(st));
    }
}

static void unpack_seeds(dartgpu_ctx *c, dartgpu_seeds *out)
{
    const int n = c->n_reads;
    const int64_t total = c->total_seeds;
    c->o_seed_gpos.resize(total); c->o_seed_rpos.resize(total); c->o_seed_len.resize(total);
    c->o_cand_off.assign(n + 1, 0);
    for (int i = 0; i < n; i++) c->o_cand_off[i + 1] = c->o_cand_off[i] + c->h_ncand.p[i];
    const int64_t nc = c->o_cand_off[n];
    c->o_cand_begin.resize(nc); c->o_cand_count.resize(nc); c->o_cand_score.resize(nc);
    const int threads = c->prm.host_threads > 0 ? c->prm.host_threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t s = 0; s < total; s++) {
        uint64_t k = c->h_keys.p[s];
        c->o_seed_gpos[s] = key_gpos(k); c->o_seed_rpos[s] = key_rpos(k); c->o_seed_len[s] = key_len(k);
    }
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n; i++) {
        int64_t so = c->h_seed_off.p[i], co = c->o_cand_off[i];
        for (uint32_t k = 0; k < c->h_ncand.p[i]; k++) {
            c->o_cand_begin[co + k] = c->h_cand_begin.p[so + k];
            c->o_cand_count[co + k] = c->h_cand_count.p[so + k];
            c->o_cand_score[co + k] = c->h_cand_score.p[so + k];
        }
    }
    out->seed_off = c->h_seed_off.p;
    out->seed_gpos = c->o_seed_gpos.data(); out->seed_rpos = c->o_seed_rpos.data(); out->seed_len = c->o_seed_len.data();
    out->cand_off = c->o_cand_off.data();
    out->cand_begin = c->o_cand_begin.data(); out->cand_count = c->o_cand_count.data(); out->cand_score = c->o_cand_score.data();
}

// ---------------------------------------------------------------------------------------------------
// stage 2 / 3 runners
// ---------------------------------------------------------------------------------------------------
void run_kmer(dartgpu_ctx *c, const uint8_t *codes_dev, const KmerJobDev *jobs, int n_jobs, int max_len1)
{
    cudaStream_t st = c->stream;
    c->h_khits.reserve(n_jobs + 1);
    if (n_jobs == 0) return;
    c->d_kjobs.reserve(n_jobs); c->d_khits.reserve(n_jobs);
    int64_t cap_recs = 64;                       // the fast path's record demand (k_kmer_prep's formula): known here, the jobs are the caller's
    for (int i = 0; i < n_jobs; i++) {
        const int64_t L1 = jobs[i].len1, L2 = jobs[i].len2;
        if (L1 >= 8 && L2 >= 8) cap_recs += std::min<int64_t>(4096, 2 * (((L2 - 7) * (L1 - 7)) >> 16) + 2 * L1 + 64);
        c->stats.kmer_window_bases += jobs[i].len2; c->stats.kmer_read_bases += jobs[i].len1;
    }
    ctl_begin(c);
    launch_set_i32(&c->d_ctl.p->nk, n_jobs, st);
    DG_CUDA(cudaMemcpyAsync(c->d_kjobs.p, jobs, (size_t)n_jobs * sizeof(KmerJobDev), cudaMemcpyHostToDevice, st));
    DG_CUDA(cudaEventRecord(c->ev[8], st));
    launch_kmer(c->ix, codes_dev, c->d_kjobs.p, &c->d_ctl.p->nk, n_jobs, max_len1, c->d_khits.p, c->kscratch, c->d_ctl.p, cap_recs, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[9], st));
    DG_CUDA(cudaMemcpyAsync(c->h_khits.p, c->d_khits.p, (size_t)n_jobs * sizeof(dartgpu_kmer_hit), cudaMemcpyDeviceToHost, st));
    ctl_fetch(c);
    DG_CUDA(dg_stream_sync(st));
    if (c->h_ctl.p[0].abort) throw std::make_pair(DARTGPU_ERR_CUDA, std::string("8-mer record pool overflow"));
    add_ms(c, &c->stats.ms_kmer, c->ev[8], c->ev[9]);
    c->stats.kernel_launches += KMER_LAUNCHES;
    c->stats.kmer_jobs += n_jobs;
    c->stats.h2d_bytes += (uint64_t)n_jobs * sizeof(KmerJobDev);
    c->stats.d2h_bytes += (uint64_t)n_jobs * sizeof(dartgpu_kmer_hit);
}

void run_nw(dartgpu_ctx *c, const uint8_t *codes_dev, NwJobDev *jobs, int n_jobs)
{
    cudaStream_t st = c->stream;
    c->o_op_off.assign(n_jobs + 1, 0);
    c->o_ops.clear();
    if (n_jobs == 0) return;
    int64_t ops_total = 0, flag_total = 0;
    int max_n = 0;
    for (int i = 0; i < n_jobs; i++) {
        ops_total += std::max(jobs[i].m, 0) + std::max(jobs[i].n, 0);
        flag_total += (int64_t)jobs[i].m * ((jobs[i].n + 15) >> 4);
        max_n = std::max(max_n, jobs[i].n);
        c->stats.nw_cells += (uint64_t)jobs[i].m * jobs[i].n;
    }
    c->stats.nw_jobs += n_jobs;
    const size_t rb_per_warp = (size_t)2 * (max_n + 1);
    c->d_njobs.reserve(n_jobs); c->d_nw_flags.reserve(flag_total + 1); c->d_nw_ops.reserve(ops_total + 1);
    c->d_nw_nops.reserve(n_jobs);
    c->d_nw_rowbuf.reserve(rb_per_warp * nw_grid_warps() + 1);
    c->h_nw_ops.reserve(ops_total + 1); c->h_nw_nops.reserve(n_jobs);
    ctl_begin(c);
    launch_set_i32(&c->d_ctl.p->nw_jobs[1], n_jobs, st);
    DG_CUDA(cudaMemcpyAsync(c->d_njobs.p, jobs, (size_t)n_jobs * sizeof(NwJobDev), cudaMemcpyHostToDevice, st));
    NwRound R{};
    R.jobs = c->d_njobs.p; R.n_jobs = &c->d_ctl.p->nw_jobs[1]; R.cap_jobs = n_jobs; R.round = 1; R.with_aux = 0;
    R.cap_ops = ops_total; R.cap_flags = flag_total; R.cap_aux = 0;
    R.flags = c->d_nw_flags.p; R.ops = c->d_nw_ops.p; R.nops = c->d_nw_nops.p; R.rowbuf = c->d_nw_rowbuf.p; R.rowbuf_per_warp = rb_per_warp;
    DG_CUDA(cudaEventRecord(c->ev[10], st));
    launch_nw(c->ix, codes_dev, R, c->d_ctl.p, c->nwscratch, st);
    DG_CUDA(cudaGetLastError());
    DG_CUDA(cudaEventRecord(c->ev[11], st));
    DG_CUDA(cudaMemcpyAsync(c->h_nw_ops.p, c->d_nw_ops.p, ops_total, cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(c->h_nw_nops.p, c->d_nw_nops.p, (size_t)n_jobs * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    DG_CUDA(cudaMemcpyAsync(jobs, c->d_njobs.p, (size_t)n_jobs * sizeof(NwJobDev), cudaMemcpyDeviceToHost, st));   // the slices the device handed out
    ctl_fetch(c);
    DG_CUDA(dg_stream_sync(st));
    if (c->h_ctl.p[0].abort) throw std::make_pair(DARTGPU_ERR_CUDA, std::string("NW pool overflow"));
    check_ctl_errors(c->h_ctl.p[0]);
    add_ms(c, &c->stats.ms_nw, c->ev[10], c->ev[11]);
    c->stats.kernel_launches += NW_LAUNCHES;
    c->stats.h2d_bytes += (uint64_t)n_jobs * sizeof(NwJobDev);
    c->stats.d2h_bytes += ops_total + (uint64_t)n_jobs * 4;
    for (int i = 0; i < n_jobs; i++) c->o_op_off[i + 1] = c->o_op_off[i] + c->h_nw_nops.p[i];
    c->o_ops.resize(c->o_op_off[n_jobs]);
    const int threads = c->prm.host_threads > 0 ? c->prm.host_threads : omp_get_max_threads();
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n_jobs; i++) {
        int k = c->h_nw_nops.p[i];
        const uint8_t *src = c->h_nw_ops.p + jobs[i].op_off + std::max(jobs[i].m, 0) + std::max(jobs[i].n, 0) - k; // written right-aligned by the traceback
        if (k) memcpy(c->o_ops.data() + c->o_op_off[i], src, k);
    }
}

// fragment bases of the stage entry points -> device codes
static const uint8_t *upload_job_bases(dartgpu_ctx *c, const char *bases, int64_t n_bases)
{
    c->h_job_codes.reserve(n_bases + 16); c->d_job_codes.reserve(n_bases + 16);
    for (int64_t i = 0; i < n_bases; i++) c->h_job_codes.p[i] = code_of((unsigned char)bases[i]);
    if (n_bases) DG_CUDA(cudaMemcpyAsync(c->d_job_codes.p, c->h_job_codes.p, n_bases, cudaMemcpyHostToDevice, c->stream));
    c->stats.h2d_bytes += n_bases;
    return c->d_job_codes.p;
}

// ---------------------------------------------------------------------------------------------------
// index hand-over
// ---------------------------------------------------------------------------------------------------
// the DevBuf members must be freed on their device; the caller's current device is put back afterwards (here the members are
// still alive, so: free them by hand, then restore)
SharedIndex::~SharedIndex()
{
    int cur = -1;
    cudaGetDevice(&cur);
    cudaSetDevice(device);
    d_occ32.release(); d_sa.release(); d_ktab.release(); d_ref2.release(); d_ends.release(); d_chr_names.release(); d_chr_name_off.release();
    if (cur >= 0) cudaSetDevice(cur);
}

// identity of an index for sharing (device and SA density not included): header fields + a hash of samples of the tables
static std::string index_key(int device, const dartgpu_index_view *v, int sa_shift, bool force64)
{
    uint64_t h = 1469598103934665603ull;
    auto mix = [&](uint64_t x) { h = (h ^ x) * 1099511628211ull; };
    mix(v->primary); mix(v->seq_len); mix(v->bwt_size); mix(v->sa_intv); mix(v->n_sa); mix((uint64_t)v->l_pac); mix((uint64_t)v->n_seqs);
    for (int i = 0; i < 5; i++) mix(v->L2[i]);
    const uint64_t step_b = std::max<uint64_t>(1, v->bwt_size / 4096), step_s = std::max<uint64_t>(1, v->n_sa / 4096);
    for (uint64_t i = 0; i < v->bwt_size; i += step_b) mix(v->bwt[i]);
    for (uint64_t i = 0; i < v->n_sa; i += step_s) mix(v->sa[i]);
    for (int i = 0; i < v->n_seqs; i++) mix((uint64_t)v->seq_len_arr[i]);
    char buf[96];
    const char *kt = getenv("DARTGPU_KTAB");
    snprintf(buf, sizeof buf, "%016llx:%d:%s", (unsigned long long)h, (int)force64, kt ? kt : "auto");
    (void)device; (void)sa_shift;
    return buf;
}

static std::mutex g_index_mutex;
static std::map<std::string, std::weak_ptr<SharedIndex>> g_indexes;

static std::shared_ptr<SharedIndex> load_shared_index(int device, const dartgpu_index_view *v, cudaStream_t st)
{
    const bool force64 = getenv("DARTGPU_FORCE_IDX64") != nullptr;
    const bool wide = force64 || v->seq_len + 2 >= (1ull << 32);
    // how dense a suffix array to keep: DARTGPU_SA_SAMPLE=<power of two <= the file's interval> pins it (tests compare the
    // LF-step counter with the reference's at 32); otherwise the densest one that fits a third of the free HBM
    int file_shift = 0;
    while ((1ull << file_shift) < v->sa_intv) file_shift++;
    int sa_shift = -1;
    if (const char *e = getenv("DARTGPU_SA_SAMPLE")) {
        long want = atol(e);
        if (want < 1 || (want & (want - 1)) || (uint64_t)want > v->sa_intv)
            throw std::make_pair(DARTGPU_ERR_ARG, std::string("DARTGPU_SA_SAMPLE must be a power of two <= the index's SA interval"));
        sa_shift = 0;
        while ((1l << sa_shift) < want) sa_shift++;
    }
    const uint64_t n_blocks32 = (v->seq_len + 63) / 64;
    const size_t occ_bytes = (n_blocks32 + 2) * 32, ref_words = (size_t)((2 * v->l_pac + 15) / 16 + 2);
    const size_t sa_width = wide ? 8 : 4;
    if (sa_shift < 0) {
        size_t free_b = 0, total_b = 0;
        DG_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t fixed = occ_bytes + ref_words * 4 + v->n_sa * 8 + (64u << 20);
        const size_t budget = free_b > fixed ? (free_b - fixed) / 3 : 0;
        sa_shift = 0;
        while (sa_shift < file_shift && ((v->seq_len >> sa_shift) + 2) * sa_width > budget) sa_shift++;
    }

    std::lock_guard<std::mutex> lock(g_index_mutex);
    const std::string ident = index_key(device, v, sa_shift, force64);
    const bool shift_pinned = getenv("DARTGPU_SA_SAMPLE") != nullptr;
    std::shared_ptr<SharedIndex> peer;                       // the same index, resident on another GPU of this process
    for (auto it = g_indexes.begin(); it != g_indexes.end();) {
        auto sp = it->second.lock();
        if (!sp) { it = g_indexes.erase(it); continue; }     // its last context is gone
        ++it;
        if (sp->ident != ident || (shift_pinned && sp->ix.sa_shift != sa_shift)) continue;
        if (sp->device == device) return sp;                 // same device: share it
        if (!peer) peer = sp;
    }
    if (peer) {
        // the peer's layout (its SA density and start table) must fit THIS device's free memory, else load from the host with
        // a density chosen for this device
        size_t free_b = 0, total_b = 0;
        DG_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = occ_bytes + ref_words * 4 + ((v->seq_len >> peer->ix.sa_shift) + 2) * sa_width +
                            (peer->ix.ktab ? ((size_t)1 << (2 * peer->ix.ktab_k)) * sizeof(KmerStart) : 0) + (256u << 20);
        int can = 0;
        cudaDeviceCanAccessPeer(&can, device, peer->device);
        if (need > free_b || !can) { cudaGetLastError(); peer.reset(); }
    }
    if (peer) sa_shift = peer->ix.sa_shift;
    const std::string key = std::to_string(device) + ":" + ident + ":" + std::to_string(sa_shift);

    auto S = std::make_shared<SharedIndex>();
    S->device = device; S->key = key; S->ident = ident; S->G = v->l_pac;
    int64_t acc = 0;
    std::vector<std::pair<int64_t, int>> ends;
    for (int i = 0; i < v->n_seqs; i++) {
        S->names.push_back(v->seq_names && v->seq_names[i] ? v->seq_names[i] : ("seq" + std::to_string(i)));
        S->chr_len.push_back(v->seq_len_arr[i]);
        S->chr_fwd.push_back(acc);
        acc += v->seq_len_arr[i];
        ends.push_back({S->chr_fwd[i] + v->seq_len_arr[i] - 1, i});          // forward copy
        ends.push_back({2 * S->G - acc + v->seq_len_arr[i] - 1, i});         // reverse-complement copy
    }
    std::sort(ends.begin(), ends.end());
    for (auto &p : ends) { S->ends.push_back(p.first); S->end_chr.push_back(p.second); }
    {   // sequence names for the device-side SAM text (sam_kernels.cu)
        std::string all; std::vector<int32_t> off{0};
        for (auto &nm : S->names) { all += nm; off.push_back((int32_t)all.size()); }
        S->d_chr_names.reserve(all.size() + 1); S->d_chr_name_off.reserve(off.size());
        DG_CUDA(cudaMemcpyAsync(S->d_chr_names.p, all.data(), all.size(), cudaMemcpyHostToDevice, st));
        DG_CUDA(cudaMemcpyAsync(S->d_chr_name_off.p, off.data(), off.size() * 4, cudaMemcpyHostToDevice, st));
        DG_CUDA(dg_stream_sync(st));
    }

    DevIndex &ix = S->ix;
    ix.n_blocks32 = n_blocks32;
    ix.sa_shift = sa_shift; ix.sa_mask = (1ull << sa_shift) - 1; ix.sa_wide = wide;
    ix.primary = v->primary; ix.seq_len = v->seq_len;
    for (int i = 0; i < 5; i++) ix.L2[i] = v->L2[i];
    ix.G = S->G; ix.n_ends = (int)S->ends.size(); ix.force64 = force64;

    if (peer) {
        // Replicate GPU -> GPU (NVLink / NVSwitch: cudaMemcpyPeer) instead of a second host->device upload, re-layout and
        // LF pass: for the 3.1 Gbp index that is ~31 GB over the switch instead of 5.4 GB over PCIe + 0.5 s of kernels.
        { cudaError_t pe = cudaDeviceEnablePeerAccess(peer->device, 0); if (pe != cudaSuccess) cudaGetLastError(); }
        const size_t sa_bytes = ((v->seq_len >> sa_shift) + 2) * sa_width;
        const size_t kt_entries = peer->ix.ktab ? ((size_t)1 << (2 * peer->ix.ktab_k)) : 0;
        S->d_occ32.reserve(occ_bytes); S->d_sa.reserve(sa_bytes); S->d_ref2.reserve(ref_words); S->d_ends.reserve(S->ends.size());
        if (kt_entries) S->d_ktab.reserve(kt_entries);
        DG_CUDA(cudaMemcpyPeerAsync(S->d_occ32.p, device, peer->d_occ32.p, peer->device, occ_bytes, st));
        DG_CUDA(cudaMemcpyPeerAsync(S->d_sa.p, device, peer->d_sa.p, peer->device, sa_bytes, st));
        DG_CUDA(cudaMemcpyPeerAsync(S->d_ref2.p, device, peer->d_ref2.p, peer->device, ref_words * 4, st));
        DG_CUDA(cudaMemcpyPeerAsync(S->d_ends.p, device, peer->d_ends.p, peer->device, S->ends.size() * 8, st));
        if (kt_entries) DG_CUDA(cudaMemcpyPeerAsync(S->d_ktab.p, device, peer->d_ktab.p, peer->device, kt_entries * sizeof(KmerStart), st));
        DG_CUDA(dg_stream_sync(st));
        ix.occ32 = reinterpret_cast<const Occ32 *>(S->d_occ32.p); ix.sa = S->d_sa.p;
        ix.ktab = kt_entries ? S->d_ktab.p : nullptr; ix.ktab_k = peer->ix.ktab_k;
        ix.ref2 = S->d_ref2.p; ix.chr_ends = S->d_ends.p;
        g_indexes[key] = S;
        return S;
    }

    // Occ blocks: raw BWA words up, re-laid-out on the device
    const uint64_t n_blocks128 = (v->seq_len + 127) / 128;
    if (v->bwt_size < n_blocks128 * 16 - 8) throw std::make_pair(DARTGPU_ERR_INDEX, std::string(".bwt is shorter than its header implies"));
    {
        DevBuf<uint32_t> raw;
        const uint64_t n_words = std::min<uint64_t>(v->bwt_size, n_blocks128 * 16);
        raw.reserve(n_words + 16);
        DG_CUDA(cudaMemcpyAsync(raw.p, v->bwt, n_words * 4, cudaMemcpyHostToDevice, st));
        S->d_occ32.reserve(occ_bytes);
        launch_relayout_occ32(raw.p, n_words, reinterpret_cast<Occ32 *>(S->d_occ32.p), n_blocks32, st);
        DG_CUDA(cudaGetLastError());
        DG_CUDA(dg_stream_sync(st));
    }
    ix.occ32 = reinterpret_cast<const Occ32 *>(S->d_occ32.p);
    {
        DevBuf<uint64_t> sa_file;
        sa_file.reserve(v->n_sa);
        DG_CUDA(cudaMemcpyAsync(sa_file.p, v->sa, v->n_sa * 8, cudaMemcpyHostToDevice, st));
        S->d_sa.reserve(((v->seq_len >> sa_shift) + 2) * sa_width);
        ix.sa = S->d_sa.p;
        launch_sa_densify(ix, sa_file.p, v->sa_intv, v->n_sa, S->d_sa.p, st);
        DG_CUDA(cudaGetLastError());
        DG_CUDA(dg_stream_sync(st));
    }
    {   // search-start table: K = floor(log4(text length)) - 1 (most K-mers of a read occur), at most 13 (1 GB);
        // DARTGPU_KTAB=<K> overrides, 0 disables
        int K = 0;
        while ((4ull << (2 * K)) <= v->seq_len) K++;          // K = floor(log4(seq_len))
        K = std::max(4, std::min(13, K - 1));
        if (const char *e = getenv("DARTGPU_KTAB")) K = std::max(0, std::min(15, atoi(e)));
        ix.ktab = nullptr; ix.ktab_k = 0;
        if (K > 0) {
            DevBuf<KmerStart> tmp;
            const size_t n_ent = (size_t)1 << (2 * K);
            S->d_ktab.reserve(n_ent); tmp.reserve(n_ent);
            launch_build_ktab(ix, K, S->d_ktab.p, tmp.p, st);
            DG_CUDA(cudaGetLastError());
            DG_CUDA(dg_stream_sync(st));
            ix.ktab = S->d_ktab.p; ix.ktab_k = K;
        }
    }
    {
        DevBuf<uint8_t> dpac;
        const size_t pac_bytes = (size_t)v->l_pac / 4 + 1;
        dpac.reserve(pac_bytes);
        DG_CUDA(cudaMemcpyAsync(dpac.p, v->pac, pac_bytes, cudaMemcpyHostToDevice, st));
        S->d_ref2.reserve(ref_words);
        launch_build_ref2(dpac.p, S->d_ref2.p, S->G, st);
        DG_CUDA(cudaGetLastError());
        DG_CUDA(dg_stream_sync(st));
    }
    S->d_ends.reserve(S->ends.size());
    DG_CUDA(cudaMemcpyAsync(S->d_ends.p, S->ends.data(), S->ends.size() * 8, cudaMemcpyHostToDevice, st));
    DG_CUDA(dg_stream_sync(st));
    ix.ref2 = S->d_ref2.p; ix.chr_ends = S->d_ends.p;
    g_indexes[key] = S;
    return S;
}

static void build_context(dartgpu_ctx *c, const dartgpu_index_view *v)
{
    if (!v || !v->bwt || !v->sa || !v->pac || v->n_seqs <= 0 || !v->seq_len_arr)
        throw std::make_pair(DARTGPU_ERR_ARG, std::string("incomplete index view"));
    if (v->sa_intv == 0 || (v->sa_intv & (v->sa_intv - 1)))
        throw std::make_pair(DARTGPU_ERR_INDEX, std::string("SA sampling interval is not a power of two"));
    for (int i = 0; i < 4; i++)
        if (v->L2[i + 1] - v->L2[i] >= (1ull << 32))
            throw std::make_pair(DARTGPU_ERR_INDEX, std::string("a symbol occurs 2^32 times or more: interval widths would not fit 32 bits"));
    if (v->seq_len != 2 * (uint64_t)v->l_pac)
        throw std::make_pair(DARTGPU_ERR_INDEX, std::string("index is not a forward+reverse-complement (FMD) index"));
    // a seed is one 64-bit key, gPos << 31 | rPos << 15 | len (dartgpu_internal.h seed_key): 33 bits of text coordinate
    if (v->seq_len >= (1ull << 33))
        throw std::make_pair(DARTGPU_ERR_INDEX, std::string("text of 2^33 symbols or more (genome over 4.29 Gbp): coordinates do not fit the 33-bit field of the seed key"));
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); throw std::make_pair(DARTGPU_ERR_NO_DEVICE, std::string("no CUDA device: libdartgpu has no CPU fallback")); }
    if (c->device < 0 || c->device >= ndev) throw std::make_pair(DARTGPU_ERR_ARG, std::string("device ordinal out of range"));
    DG_CUDA(cudaSetDevice(c->device));
    DG_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    for (auto &ev : c->ev) DG_CUDA(cudaEventCreate(&ev));
    c->shared = load_shared_index(c->device, v, c->stream);
    c->ix = c->shared->ix;
    c->G = c->shared->G;
    c->d_ctl.reserve(1); c->h_ctl.reserve(1);
    launch_zero(c->d_ctl.p, sizeof(BatchCtl), c->stream);     // incl. the ingest fields, which only the uploads reset
    DG_CUDA(cudaEventCreateWithFlags(&c->done, cudaEventBlockingSync | cudaEventDisableTiming));
    DG_CUDA(cudaEventCreateWithFlags(&c->compute_done, cudaEventDisableTiming));
    DG_CUDA(dg_stream_sync(c->stream));
}

static bool slurp(const std::string &fn, std::vector<uint8_t> &buf)
{
    FILE *fp = fopen(fn.c_str(), "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    buf.resize(n);
    size_t got = n ? fread(buf.data(), 1, n, fp) : 0;
    fclose(fp);
    return got == (size_t)n;
}

} // namespace dartgpu

// =====================================================================================================
// extern "C"
// =====================================================================================================
extern "C" {

void dartgpu_default_params(dartgpu_params *p)
{
    if (!p) return;
    p->max_gaps = 5; p->max_intron = 500000; p->min_intron = 5; p->max_mismatch = 0; p->max_dup = 100;
    p->multi_hit = 0; p->pair_end = 0; p->all_sj = 0; p->unique = 0; p->host_threads = 0;
}

static void sanitize(dartgpu_params &p)
{   // the clamps of the reference's flag parser (/root/reference/src/main.cpp:173-178, :187)
    if (p.max_dup < 100) p.max_dup = 100; else if (p.max_dup >= 10000) p.max_dup = 10000;
    if (p.max_intron < 100000) p.max_intron = 100000;
}

int dartgpu_create(dartgpu_ctx **out, int device, const dartgpu_index_view *idx, const dartgpu_params *p)
{
    if (!out) return fail(nullptr, DARTGPU_ERR_ARG, "out is NULL");
    *out = nullptr;
    dartgpu_ctx *c = new (std::nothrow) dartgpu_ctx;
    if (!c) return fail(nullptr, DARTGPU_ERR_NOMEM, "out of host memory");
    c->device = device;
    if (p) c->prm = *p; else dartgpu_default_params(&c->prm);
    sanitize(c->prm);
    int rc = guarded(nullptr, [&] { build_context(c, idx); });
    if (rc != DARTGPU_OK) { dartgpu_destroy(c); return rc; }
    *out = c;
    return DARTGPU_OK;
}

int dartgpu_create_from_files(dartgpu_ctx **out, int device, const char *prefix, const dartgpu_params *p)
{
    if (!out || !prefix) return fail(nullptr, DARTGPU_ERR_ARG, "NULL argument");
    *out = nullptr;
    std::vector<uint8_t> bwt, sa, pac;
    std::string pre(prefix);
    if (!slurp(pre + ".bwt", bwt) || bwt.size() < 40) return fail(nullptr, DARTGPU_ERR_INDEX, "cannot read " + pre + ".bwt");
    if (!slurp(pre + ".sa", sa) || sa.size() < 56) return fail(nullptr, DARTGPU_ERR_INDEX, "cannot read " + pre + ".sa");
    if (!slurp(pre + ".pac", pac)) return fail(nullptr, DARTGPU_ERR_INDEX, "cannot read " + pre + ".pac");
    dartgpu_index_view v{};
    memcpy(&v.primary, bwt.data(), 8);
    memcpy(&v.L2[1], bwt.data() + 8, 32);
    v.L2[0] = 0;
    v.seq_len = v.L2[4];
    v.bwt_size = (bwt.size() - 40) / 4;
    v.bwt = reinterpret_cast<const uint32_t *>(bwt.data() + 40);
    memcpy(&v.sa_intv, sa.data() + 40, 8);
    if (v.sa_intv == 0) return fail(nullptr, DARTGPU_ERR_INDEX, pre + ".sa: bad sampling interval");
    v.n_sa = (v.seq_len + v.sa_intv) / v.sa_intv;
    if (sa.size() < 56 + (v.n_sa - 1) * 8) return fail(nullptr, DARTGPU_ERR_INDEX, pre + ".sa is truncated");
    std::vector<uint64_t> sav(v.n_sa);
    sav[0] = (uint64_t)-1;
    memcpy(sav.data() + 1, sa.data() + 56, (v.n_sa - 1) * 8);
    v.sa = sav.data();
    FILE *fp = fopen((pre + ".ann").c_str(), "r");
    if (!fp) return fail(nullptr, DARTGPU_ERR_INDEX, "cannot read " + pre + ".ann");
    long long lpac; int nseq; unsigned seed;
    std::vector<std::string> names; std::vector<int64_t> lens;
    if (fscanf(fp, "%lld%d%u", &lpac, &nseq, &seed) != 3) { fclose(fp); return fail(nullptr, DARTGPU_ERR_INDEX, pre + ".ann: bad header"); }
    for (int i = 0; i < nseq; i++) {
        unsigned gi; char name[1024]; long long off; int len, namb;
        if (fscanf(fp, "%u%1023s", &gi, name) != 2) break;
        int ch; while ((ch = fgetc(fp)) != '\n' && ch != EOF) {}
        if (fscanf(fp, "%lld%d%d", &off, &len, &namb) != 3) break;
        names.push_back(name); lens.push_back(len);
    }
    fclose(fp);
    if ((int)names.size() != nseq) return fail(nullptr, DARTGPU_ERR_INDEX, pre + ".ann: truncated");
    std::vector<const char *> namep;
    for (auto &s : names) namep.push_back(s.c_str());
    v.l_pac = lpac; v.pac = pac.data(); v.n_seqs = nseq; v.seq_len_arr = lens.data(); v.seq_names = namep.data();
    if ((int64_t)pac.size() < lpac / 4 + 1) return fail(nullptr, DARTGPU_ERR_INDEX, pre + ".pac is truncated");
    return dartgpu_create(out, device, &v, p);
}

int dartgpu_index_build(int device, const uint8_t *pac, int64_t l_pac, const char *prefix, uint64_t max_suffixes_per_pass)
{
    if (!pac || !prefix || l_pac <= 0) return fail(nullptr, DARTGPU_ERR_ARG, "bad argument");
    return guarded(nullptr, [&] {
        int ndev = 0;
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); throw std::make_pair(DARTGPU_ERR_NO_DEVICE, std::string("no CUDA device: libdartgpu has no CPU fallback")); }
        if (device < 0 || device >= ndev) throw std::make_pair(DARTGPU_ERR_ARG, std::string("device ordinal out of range"));
        build_index_files(device, pac, l_pac, prefix, max_suffixes_per_pass);
    });
}

void dartgpu_destroy(dartgpu_ctx *c)
{
    if (!c) return;
    int cur = -1;
    cudaGetDevice(&cur);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{cur};      // the caller's current device stays what it was
    cudaSetDevice(c->device);
    if (c->in_flight) cudaStreamSynchronize(c->stream);
    for (auto &ev : c->ev) if (ev) cudaEventDestroy(ev);
    if (c->done) cudaEventDestroy(c->done);
    if (c->compute_done) {
        if (c->shared) {
            std::lock_guard<std::mutex> lock(c->shared->turn_mutex);
            for (int i = 0; i < SharedIndex::TURN_RING; i++)
                if (c->shared->turn_owner[i] == c) { c->shared->turn_ev[i] = nullptr; c->shared->turn_owner[i] = nullptr; }
        }
        cudaStreamSynchronize(c->stream);
        cudaEventDestroy(c->compute_done);
    }
    if (c->dpipe) free_device_pipe(c->dpipe);
    if (c->own_stream) cudaStreamDestroy(c->own_stream);
    delete c;
}

int dartgpu_set_params(dartgpu_ctx *c, const dartgpu_params *p)
{
    if (!c || !p) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    c->prm = *p; sanitize(c->prm);
    return DARTGPU_OK;
}
const char *dartgpu_last_error(const dartgpu_ctx *c) { return c ? c->err.c_str() : g_create_error.c_str(); }
int64_t dartgpu_genome_size(const dartgpu_ctx *c) { return c ? c->G : 0; }
int dartgpu_num_sequences(const dartgpu_ctx *c) { return c ? (int)c->shared->names.size() : 0; }
const char *dartgpu_sequence_name(const dartgpu_ctx *c, int i) { return (c && i >= 0 && i < (int)c->shared->names.size()) ? c->shared->names[i].c_str() : ""; }
int64_t dartgpu_sequence_length(const dartgpu_ctx *c, int i) { return (c && i >= 0 && i < (int)c->shared->chr_len.size()) ? c->shared->chr_len[i] : 0; }
int dartgpu_set_result_location(dartgpu_ctx *c, int on_device)
{
    if (!c) return DARTGPU_ERR_ARG;
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    c->results_on_device = on_device != 0;
    return DARTGPU_OK;
}

// Pins the calling host thread to the CPUs next to `device` (its PCIe root: /sys/bus/pci/devices/<bus id>/local_cpulist), so
// that the page-locked buffers the thread allocates afterwards and its copies stay on the GPU's own NUMA node.
int dartgpu_bind_host_thread(int device)
{
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) { cudaGetLastError(); return DARTGPU_ERR_ARG; }
    for (char *p = bus; *p; p++) if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/local_cpulist";
    FILE *fp = fopen(path.c_str(), "r");
    if (!fp) return DARTGPU_ERR_ARG;
    char line[4096] = {0};
    const bool ok = fgets(line, sizeof line, fp) != nullptr;
    fclose(fp);
    if (!ok) return DARTGPU_ERR_ARG;
    cpu_set_t set; CPU_ZERO(&set);
    int n_set = 0;
    for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        if (sscanf(tok, "%d-%d", &a, &b) == 2) { for (int i = a; i <= b && i < CPU_SETSIZE; i++) { CPU_SET(i, &set); n_set++; } }
        else if (sscanf(tok, "%d", &a) == 1 && a < CPU_SETSIZE) { CPU_SET(a, &set); n_set++; }
    }
    if (!n_set) return DARTGPU_ERR_ARG;
    return sched_setaffinity(0, sizeof set, &set) == 0 ? DARTGPU_OK : DARTGPU_ERR_ARG;
}

int dartgpu_set_stream(dartgpu_ctx *c, void *s)
{
    if (!c) return DARTGPU_ERR_ARG;
    c->stream = s ? (cudaStream_t)s : c->own_stream;
    return DARTGPU_OK;
}

int dartgpu_seed_and_cluster(dartgpu_ctx *c, const dartgpu_reads *reads, dartgpu_seeds *out)
{
    if (!c || !reads || !out) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    return guarded(c, [&] {
        Timer t;
        stats_begin(c);
        upload_reads(c, reads);
        c->timed_upload = true;
        run_seeding_sync(c, true);
        unpack_seeds(c, out);
        c->stats.ms_host = t.ms();
    });
}

int dartgpu_upload_reads(dartgpu_ctx *c, const dartgpu_reads *reads)
{
    if (!c || !reads) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    return guarded(c, [&] { stats_begin(c); upload_reads(c, reads); DG_CUDA(dg_stream_sync(c->stream)); });
}

int dartgpu_seed_and_cluster_resident(dartgpu_ctx *c)
{
    if (!c) return DARTGPU_ERR_ARG;
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    return guarded(c, [&] {
        uint64_t rb = c->stats.read_bases;
        stats_begin(c);
        c->stats.read_bases = rb;
        c->timed_upload = false;
        run_seeding_sync(c, false);
    });
}

int dartgpu_synchronize(dartgpu_ctx *c)
{
    if (!c) return DARTGPU_ERR_ARG;
    return guarded(c, [&] { DG_CUDA(dg_stream_sync(c->stream)); });
}

int dartgpu_kmer_reseed(dartgpu_ctx *c, const char *bases, int64_t n_bases, const dartgpu_kmer_job *jobs, int32_t n_jobs,
                        const dartgpu_kmer_hit **out)
{
    if (!c || !out || n_jobs < 0 || (n_jobs && (!jobs || !bases))) return fail(c, DARTGPU_ERR_ARG, "bad argument");
    return guarded(c, [&] {
        stats_begin(c);
        const uint8_t *dcodes = upload_job_bases(c, bases, n_bases);
        c->h_kjobs.reserve(n_jobs + 1);
        int max1 = 8;
        for (int i = 0; i < n_jobs; i++) {
            const dartgpu_kmer_job &j = jobs[i];
            if (j.frag_off < 0 || j.frag_len < 0 || j.frag_off + j.frag_len > n_bases || j.glen < 0 || j.gpos < 0 ||
                j.gpos + j.glen > 2 * c->G)
                throw std::make_pair(DARTGPU_ERR_ARG, std::string("k-mer job out of range"));
            if (j.frag_len > DARTGPU_MAX_RLEN) throw std::make_pair(DARTGPU_ERR_READ_TOO_LONG, std::string("k-mer fragment too long"));
            c->h_kjobs.p[i] = KmerJobDev{j.frag_off, j.gpos, j.frag_len, j.glen};
            max1 = std::max(max1, j.frag_len);
        }
        run_kmer(c, dcodes, c->h_kjobs.p, n_jobs, max1);
        *out = c->h_khits.p;
    });
}

int dartgpu_nw_align(dartgpu_ctx *c, const char *bases, int64_t n_bases, const dartgpu_nw_job *jobs, int32_t n_jobs,
                     dartgpu_nw_result *out)
{
    if (!c || !out || n_jobs < 0 || (n_jobs && (!jobs || !bases))) return fail(c, DARTGPU_ERR_ARG, "bad argument");
    return guarded(c, [&] {
        stats_begin(c);
        const uint8_t *dcodes = upload_job_bases(c, bases, n_bases);
        c->h_njobs.reserve(n_jobs + 1);
        for (int i = 0; i < n_jobs; i++) {
            const dartgpu_nw_job &j = jobs[i];
            if (j.frag_off < 0 || j.m < 0 || j.frag_off + j.m > n_bases || j.n < 0 || j.gpos < 0 || j.gpos + j.n > 2 * c->G)
                throw std::make_pair(DARTGPU_ERR_ARG, std::string("NW job out of range"));
            c->h_njobs.p[i] = NwJobDev{j.frag_off, j.gpos, 0, 0, 0, j.m, j.n};
        }
        run_nw(c, dcodes, c->h_njobs.p, n_jobs);
        out->op_off = c->o_op_off.data();
        out->ops = c->o_ops.data();
    });
}

// ---- the whole per-read path: submit (enqueue everything, return) / wait (sleep until the batch is done) ----
namespace dartgpu {
// everything of one attempt behind the (already enqueued or resident) read upload; records the `done` event
__global__ void k_ctl_ingest(BatchCtl *ctl, int max_rlen_host)
{
    if (ctl->ingest_err) atomicOr(&ctl->abort, 1 << 30);   // bad text: nothing downstream may run
    else if (ctl->ingest_max_rlen > max_rlen_host) atomicOr(&ctl->abort, CAP_RLEN);
}

static void enqueue_whole_path(dartgpu_ctx *c)
{
    compute_turn_begin(c);
    ctl_begin(c);
    c->whole_path_enqueue = true;
    if (c->from_fastq) k_ctl_ingest<<<1, 1, 0, c->stream>>>(c->d_ctl.p, c->max_rlen);
    enqueue_seeding(c);
    c->whole_path_enqueue = false;
    enqueue_pipeline(c);
    ctl_fetch(c);
    DG_CUDA(cudaEventRecord(c->done, c->stream));
}

static void submit_batch(dartgpu_ctx *c, const dartgpu_reads *reads, const dartgpu_fastq_block *fq = nullptr)
{   // reads == nullptr and fq == nullptr: the batch uploaded with dartgpu_upload_reads
    Timer t;
    uint64_t rb = c->stats.read_bases;
    stats_begin(c);
    c->emit_sam = fq != nullptr;
    if (fq) upload_fastq(c, fq);
    else if (reads) upload_reads(c, reads);
    else c->stats.read_bases = rb;
    c->timed_upload = reads != nullptr || fq != nullptr;
    c->whole_path = true;
    c->attempts = 0;
    if (c->n_reads > 0) {
        caps_for_batch(c);
        enqueue_whole_path(c);
    }
    c->in_flight = true;
    c->t_submit_ms = t.ms();
    static const bool trace = getenv("DARTGPU_TRACE") != nullptr;
    if (trace) fprintf(stderr, "TRACE ctx %p submit %d reads: %.2f ms on the host\n", (void *)c, c->n_reads, c->t_submit_ms);
}

static void wait_batch(dartgpu_ctx *c, dartgpu_map_result *out, dartgpu_sam_result *sam_out)
{
    Timer t;
    c->in_flight = false;
    if (c->n_reads == 0) {
        if (out) *out = dartgpu_map_result{nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0};
        if (sam_out) *sam_out = dartgpu_sam_result{nullptr, 0, 0, 0, 0, 0, nullptr, 0};
        return;
    }
    for (;;) {
        DG_CUDA(cudaEventSynchronize(c->done));            // the batch's one wait: a sleeping thread, not a spinning one
        const BatchCtl &H = c->h_ctl.p[0];
        if (H.ingest_err) check_ctl_errors(H);
        if (!H.abort) break;
        // a pool was too small (first batch of a context, or a batch unlike the ones before): grow it and run the batch again
        if (++c->attempts > 12) throw std::make_pair(DARTGPU_ERR_NOMEM, std::string("device pools keep overflowing"));
        static const bool trace = getenv("DARTGPU_TRACE") != nullptr;
        if (trace) fprintf(stderr, "TRACE ctx %p attempt %d aborted: caps 0x%x (seeds %lld cands %lld pool %lld krecs %lld cig %lld text %lld junc %lld sam %lld nw %llu/%llu)\n",
                           (void *)c, c->attempts, (unsigned)H.abort, H.total_seeds, H.ncand, H.pool_total, H.kmer_recs, H.cig_total, H.text_total, H.junc_total,
                           H.sam_bytes, H.nw_ops[0], H.nw_ops[1]);
        caps_grow(c, H);
        const uint64_t keep_launches = c->stats.kernel_launches;
        enqueue_whole_path(c);
        c->stats.kernel_launches = keep_launches;        // only the attempt that goes through is counted
    }
    check_ctl_errors(c->h_ctl.p[0]);
    if (sam_out) finish_sam(c, sam_out); else finish_pipeline(c, out);
    if (c->from_fastq) c->stats.read_bases = c->h_ctl.p[0].ingest_bases;
    collect_stats(c, true);
    c->stats.ms_host = c->t_submit_ms + t.ms();
    c->stats.ms_submit = c->t_submit_ms;
}
} // namespace dartgpu

int dartgpu_submit(dartgpu_ctx *c, const dartgpu_reads *reads)
{
    if (!c || !reads) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    if (c->prm.pair_end && (reads->n_reads & 1)) return fail(c, DARTGPU_ERR_ARG, "paired-end batch with an odd number of reads");
    return guarded(c, [&] { submit_batch(c, reads); });
}

int dartgpu_submit_resident(dartgpu_ctx *c)
{
    if (!c) return DARTGPU_ERR_ARG;
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    if (c->prm.pair_end && (c->n_reads & 1)) return fail(c, DARTGPU_ERR_ARG, "paired-end batch with an odd number of reads");
    return guarded(c, [&] { submit_batch(c, nullptr); });
}

int dartgpu_submit_fastq(dartgpu_ctx *c, const dartgpu_fastq_block *block)
{
    if (!c || !block) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (c->in_flight) return fail(c, DARTGPU_ERR_ARG, "a submitted batch is still in flight on this context: dartgpu_wait first");
    return guarded(c, [&] { submit_batch(c, nullptr, block); });
}

int dartgpu_wait_sam(dartgpu_ctx *c, dartgpu_sam_result *out)
{
    if (!c || !out) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (!c->in_flight || !c->emit_sam) return fail(c, DARTGPU_ERR_ARG, "no FASTQ batch in flight on this context");
    int rc = guarded(c, [&] { wait_batch(c, nullptr, out); });
    if (rc != DARTGPU_OK) { c->in_flight = false; cudaStreamSynchronize(c->stream); cudaGetLastError(); }
    return rc;
}

void *dartgpu_alloc_pinned(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void dartgpu_free_pinned(void *p) { if (p) cudaFreeHost(p); }

// Cuts a buffer of FASTQ text at a record boundary: counts newlines (memchr) up to `max_records` records and returns the
// bytes they span; *n_records = complete records found.  A last line without '\n' is not counted (read more, or append one).
int64_t dartgpu_fastq_cut(const char *text, int64_t len, int32_t max_records, int32_t *n_records)
{
    // Newlines are counted in 64 KB pieces (a plain byte loop: the host compiler vectorises it, ~10 GB/s) and only the piece
    // that holds the cut is walked line by line: one memchr call per LINE was the reader's biggest cost (20 ms per 60 MB).
    const int64_t want = 4ll * (max_records > 0 ? max_records : INT32_MAX / 8);
    if (len < 0) len = 0;
    int64_t lines = 0, pos = 0;
    const int64_t PIECE = 1 << 16;
    while (pos < len) {
        const int64_t e = std::min(len, pos + PIECE);
        int64_t k = 0;
        const char *p = text + pos;
        for (int64_t i = 0, m = e - pos; i < m; i++) k += p[i] == '\n';
        if (lines + k >= want) break;
        lines += k; pos = e;
    }
    int64_t end_of_record = 0;
    if (pos < len) {                     // the wanted newline is in [pos, pos + PIECE)
        const char *p = text + pos, *e = text + len;
        while (lines < want) { p = (const char *)memchr(p, '\n', (size_t)(e - p)) + 1; lines++; }
        end_of_record = p - text;
    } else {                             // fewer newlines than wanted: cut behind the last complete record
        const int64_t target = lines & ~(int64_t)3;
        int64_t skip = lines - target;    // newlines behind the cut
        const char *q = text + len;
        while (target > 0) {
            q = (const char *)memrchr(text, '\n', (size_t)(q - text));
            if (skip-- == 0) { end_of_record = q + 1 - text; break; }
        }
        lines = target;
    }
    if (n_records) *n_records = (int32_t)(lines / 4);
    return end_of_record;
}

int dartgpu_wait(dartgpu_ctx *c, dartgpu_map_result *out)
{
    if (!c || !out) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (!c->in_flight || c->emit_sam) return fail(c, DARTGPU_ERR_ARG, "no batch in flight on this context");
    int rc = guarded(c, [&] { wait_batch(c, out, nullptr); });
    if (rc != DARTGPU_OK) { c->in_flight = false; cudaStreamSynchronize(c->stream); cudaGetLastError(); }
    return rc;
}

int dartgpu_map_reads(dartgpu_ctx *c, const dartgpu_reads *reads, dartgpu_map_result *out)
{
    if (!c || !reads || !out) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    int rc = dartgpu_submit(c, reads);
    return rc != DARTGPU_OK ? rc : dartgpu_wait(c, out);
}

int dartgpu_map_reads_resident(dartgpu_ctx *c, const dartgpu_reads *reads, dartgpu_map_result *out)
{
    if (!c || !reads || !out) return fail(c, DARTGPU_ERR_ARG, "NULL argument");
    if (reads->n_reads != c->n_reads) return fail(c, DARTGPU_ERR_ARG, "batch differs from the uploaded one");
    int rc = dartgpu_submit_resident(c);
    return rc != DARTGPU_OK ? rc : dartgpu_wait(c, out);
}

int dartgpu_measure_int32_peak(dartgpu_ctx *c, double *ops_per_second)
{
    if (!c || !ops_per_second) return DARTGPU_ERR_ARG;
    return guarded(c, [&] { *ops_per_second = measure_int32_ops_per_second(c->stream); });
}

int dartgpu_measure_l2_peak(dartgpu_ctx *c, uint64_t table_bytes, double *bytes_per_second)
{
    if (!c || !bytes_per_second || table_bytes < 64) return DARTGPU_ERR_ARG;
    return guarded(c, [&] { *bytes_per_second = measure_l2_gather_bytes_per_second(c->stream, (size_t)table_bytes); });
}

int dartgpu_get_stats(const dartgpu_ctx *c, dartgpu_stats *out)
{
    if (!c || !out) return DARTGPU_ERR_ARG;
    *out = c->stats;
    return DARTGPU_OK;
}

} // extern "C"
