// Internal declarations shared by the CUDA kernels, the host orchestration and the C-ABI glue.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/dartgpu.h"

namespace dartgpu {

// ---------------------------------------------------------------------------------------------------
// Device-resident index (replicated into every GPU's HBM once, at context creation)
// ---------------------------------------------------------------------------------------------------
// Occ table: BWA's 64-byte block per 128 BWT symbols (4 x u64 counts followed by 8 x u32 of symbols,
// /root/reference/src/BWT_Index/bwtindex.c:53-75) is re-laid-out on the GPU into Occ32 blocks (rank.cuh): 32 bytes = one
// sector per 64 symbols, {u32 count[4], u64 low bit-plane, u64 high bit-plane}, so that ONE thread resolves a rank query
// with one 128-bit load + one 32-bit load from a single sector.
// Suffix array: the file holds every 32nd entry (bwt_cal_sa, /root/reference/src/BWT_Index/bwt.c:101-123).  HBM has room
// for more, so at load the GPU walks the LF mapping once over the whole text and materialises every 2^sa_shift-th entry
// (shift 0 = the full suffix array when it fits the memory budget): locate becomes one gather instead of ~31 dependent
// LF steps per hit.  Entries are u32 when the text length fits 32 bits, u64 otherwise.
struct Occ32;
// Search-start table: the SA interval (of the reverse-complemented pattern, as k_search carries it) after the first K
// bases of a search, for every K-mer, plus the number of those K-1 steps whose two rank queries fall into different
// 128-symbol BWA blocks (the reference's work counter).  One 16-byte gather replaces K-1 dependent rank steps — on a
// human-sized index exactly the steps whose two queries land in two random HBM sectors each.
struct KmerStart { uint64_t x1; uint32_t x2; uint32_t splits; };
struct DevIndex {
    const KmerStart *ktab;   // 4^ktab_k entries, k-mer index = first base in the low bits; nullptr when disabled
    int ktab_k;
    const Occ32 *occ32;      // n_blocks32 blocks (+1 zero guard block)
    uint64_t n_blocks32;
    const void *sa;          // sa[i >> sa_shift] = SA[i] for i % 2^sa_shift == 0; entry 0 is never read (it stands for -1)
    uint64_t sa_mask;        // 2^sa_shift - 1
    int sa_shift;
    int sa_wide;             // 1: u64 entries, 0: u32
    uint64_t primary, seq_len;
    uint64_t L2[5];
    const uint32_t *ref2;    // reference over [0,2G), 2 bits/base, 16 bases per word, base i at bits 30-2(i&15)
    int64_t G;
    const int64_t *chr_ends; // sorted last coordinates of every sequence on both strands (ChrLocMap keys)
    int n_ends;
    int force64;             // tests only (DARTGPU_FORCE_IDX64=1): run the 64-bit interval kernels on a small index
};

struct SearchRec {           // one qualifying BWT_Search result: SA interval (of the reverse-complemented match) still to be located
    uint64_t sa_begin;
    uint32_t freq;
    uint16_t start, len;
};

// seed key: gPos << 31 | rPos << 15 | len  — sorting the keys sorts by (gPos, rPos) as CompByGenomePos does
__host__ __device__ inline uint64_t seed_key(uint64_t g, uint32_t r, uint32_t len) { return g << 31 | (uint64_t)r << 15 | len; }
__host__ __device__ inline int64_t key_gpos(uint64_t k) { return (int64_t)(k >> 31); }
__host__ __device__ inline int key_rpos(uint64_t k) { return (int)((k >> 15) & 0xFFFF); }
__host__ __device__ inline int key_len(uint64_t k) { return (int)(k & 0x7FFF); }

// device read encoding (one byte per base): 0..3 = A,C,G,T, 8..11 = a,c,g,t (bit 3 = lower case: the reference compares
// raw characters in places), 4 = any other symbol, 5 = a literal 'N' (the only symbol that breaks an 8-mer,
// /root/reference/src/KmerAnalysis.cpp:44).  Bit 2 set <=> not ACGT; (code & 3) is the base otherwise.
enum { CODE_OTHER = 4, CODE_N = 5 };

struct DevStats {            // device-side work counters (see dartgpu_stats)
    unsigned long long ext_steps, ext_blocks, lf_steps, hits, seeds, sector_loads;
};

// Control block of ONE batch, in device memory: every count the host used to read back between kernel launches
// (round 1: ~20 stream synchronisations per batch, which tied every context to a spinning host thread and capped 8-GPU
// weak scaling at 0.58) now stays here.  Kernels read their loop bounds from it, pools have host-chosen CAPACITIES, and a
// kernel that would exceed one sets `abort` instead: everything enqueued behind it returns at once, and the host -- which
// looks at this block exactly once per batch, together with the results -- grows the pool and runs the batch again.
// In steady state (batches of similar size) no batch is ever repeated and the host synchronises once per batch.
enum { CAP_SEEDS = 1, CAP_CANDS = 2, CAP_POOL = 4, CAP_KRECS = 8, CAP_NW_B = 16, CAP_NW_C = 32, CAP_CIG = 64, CAP_TEXT = 128,
       CAP_JUNC = 256, CAP_SAM = 512, CAP_RLEN = 1024 };
enum { ERR_CIGAR_POOL = 1, ERR_SORT_SCRATCH = 2, ERR_NW_WIDTH = 4, ERR_FASTQ_LINES = 8, ERR_READ_TOO_LONG = 16 };
struct BatchCtl {
    long long total_seeds, ncand, nrep, pool_total, cig_total, text_total, junc_total, kmer_recs;
    unsigned long long nw_ops[2], nw_flags[2], nw_aux[2];   // pool use of the two NW rounds (B: gap flanks, C: non-simple pairs)
    int32_t nk, nw_jobs[2], nw_small[2];                    // job-queue lengths; jobs of the thread-per-alignment class
    int32_t nw_max_n[2];
    int32_t phase_queue[3];                                 // candidates waiting for the repair phases B, C, D (report_kernels.cu)
    int32_t pad0;
    int32_t abort;                                          // CAP_* bits: a capacity was exceeded, the batch must be re-run
    int32_t err;                                            // ERR_* bits: internal invariants that must never fire
    uint32_t steal, big_count, mid_count, heavy_count;    // work-stealing / queue counters of the seeding and 8-mer kernels
    uint32_t pad;
    unsigned long long work[4];                             // [0] NW cells, [1] 8-mer window bases, [2] 8-mer read bases
    long long sam_bytes;                                    // SAM text of the batch (dartgpu_submit_fastq)
    unsigned long long sam_counts[4];                       // unmapped reads, unique reads, paired reads (the reference's summary lines)
    DevStats stats;
    // ---- written by the ingest kernels BEFORE the batch's first attempt and kept across attempts (not part of the reset) ----
    int32_t ingest_err;                                     // ERR_* bits of the FASTQ parse
    int32_t ingest_max_rlen;                                // longest read of the batch
    unsigned long long ingest_bases;
};
constexpr size_t BATCHCTL_RESET_BYTES = offsetof(BatchCtl, ingest_err);

// ---------------------------------------------------------------------------------------------------
// small RAII buffers
// ---------------------------------------------------------------------------------------------------
struct CudaError { cudaError_t e; const char *what; const char *file; int line; };
#define DG_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw ::dartgpu::CudaError{e_, #x, __FILE__, __LINE__}; } while (0)

template <class T> struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        DG_CUDA(cudaMalloc((void **)&p, want * sizeof(T)));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() { release(); }
};
template <class T> struct PinBuf {
    T *p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        DG_CUDA(cudaMallocHost((void **)&p, want * sizeof(T)));
        cap = want;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    ~PinBuf() { release(); }
};

// Host-side wait for a stream.  The waiting thread SLEEPS on a blocking event: with one wait per batch (BatchCtl above) the
// wake-up latency is paid once per batch instead of ~20 times, and one host thread per GPU is enough to keep several batches
// in flight without stealing the cores the other ranks' threads need.  DARTGPU_SYNC=spin restores the spinning wait.
inline cudaError_t dg_stream_sync(cudaStream_t st)
{
    static const bool spin = [] { const char *e = getenv("DARTGPU_SYNC"); return e && !strcmp(e, "spin"); }();
    if (spin) return cudaStreamSynchronize(st);
    thread_local cudaEvent_t ev = nullptr;
    thread_local int ev_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (!ev || ev_dev != dev) {
        if (ev) cudaEventDestroy(ev);
        cudaError_t e = cudaEventCreateWithFlags(&ev, cudaEventBlockingSync | cudaEventDisableTiming);
        if (e != cudaSuccess) { ev = nullptr; return e; }
        ev_dev = dev;
    }
    cudaError_t e = cudaEventRecord(ev, st);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev);
}

// ---------------------------------------------------------------------------------------------------
// kernel launchers (each file owns its kernels; all work is enqueued on `st`)
// ---------------------------------------------------------------------------------------------------
// index_device.cu
int sm_count();                                   // SMs of the current device (cached per device, thread-safe)
void launch_set_i32(int32_t *dst, int32_t v, cudaStream_t st);
// Zero-fill by a kernel.  Not cudaMemsetAsync: memsets of more than a few KB may be executed by a copy engine, where they
// queue behind the other contexts' multi-megabyte result copies and stall this context's kernel stream (measured: with
// two 4 MB memsets per batch no result copy overlapped any kernel, 7.1 ms per step instead of 5.3).
void launch_zero(void *p, size_t bytes, cudaStream_t st);   // p 4-byte aligned, bytes a multiple of 4
// dst (a field of *ctl) = *src; sets `bit` in ctl->abort when the value exceeds cap
void launch_ctl_check(BatchCtl *ctl, long long *dst, const int64_t *src, int64_t cap, int bit, cudaStream_t st);
void small_d2h(void *host_pinned, const void *dev, size_t bytes, cudaStream_t st);   // bytes: multiple of 4, pinned destination
void launch_relayout_occ32(const uint32_t *bwt_words, uint64_t n_words, Occ32 *occ, uint64_t n_blocks32, cudaStream_t st);
// sa_file: the reference's sampled SA (every sa_intv-th entry, entry 0 = -1) on the device; out: every 2^shift-th entry
void launch_build_ktab(const DevIndex &ix, int K, KmerStart *out, KmerStart *tmp, cudaStream_t st);   // out, tmp: 4^K entries each
void launch_sa_densify(const DevIndex &ix, const uint64_t *sa_file, uint64_t sa_intv, uint64_t n_sa_file, void *out, cudaStream_t st);
void launch_build_ref2(const uint8_t *pac, uint32_t *ref2, int64_t G, cudaStream_t st);
void launch_read_layout(const int64_t *off, int n, int32_t *rlen, uint32_t *padded, cudaStream_t st);
void launch_encode_reads(const uint8_t *raw, const int64_t *off, const int64_t *dev_off, int n, uint8_t *codes, uint2 *packed, cudaStream_t st);

// index_build.cu
void build_index_files(int device, const uint8_t *pac, int64_t l_pac, const char *prefix, uint64_t limit);

// seed_kernels.cu
struct SeedLaunch {
    const uint8_t *codes; const int64_t *dev_off; const int32_t *rlen; int n_reads;
    const uint2 *packed;                                        // 16 bases per entry: .x = 2-bit codes (base i at bits 2i), .y = "not ACGT" bits
    uint32_t *steal; int steal_base; int turn_batch, end_batch;                            // work-stealing counter for the tail of the search kernel (in the control block)
    int cap_rec; uint32_t max_dup; int max_gaps, max_intron;
    SearchRec *recs; uint32_t *nrec; uint32_t *nhits;          // search output
    int64_t *seed_off;                                          // n_reads+1, exclusive scan of nhits
    uint64_t *keys; uint32_t *meta;                             // per seed slot
    int32_t *cand_begin, *cand_count, *cand_score; uint32_t *ncand;
    uint32_t *mid_list; uint32_t *mid_count;                    // reads with 9..32 seeds (one warp each)
    uint32_t *big_list; uint32_t *big_count;                    // reads with more than 32 seeds (one block each)
    uint64_t *big_scratch; size_t big_scratch_per_cta;          // global sort scratch for reads that exceed smem
    DevStats *stats;
    BatchCtl *ctl;                                              // ctl->total_seeds bounds the per-hit kernels
    int count_work;                                             // k_search keeps the work counters (stage entry points, DARTGPU_STATS=1)
};
void launch_search(const DevIndex &ix, SeedLaunch a, cudaStream_t st);
void launch_scan_hits(const SeedLaunch &a, void *tmp, size_t tmp_bytes, cudaStream_t st);
size_t scan_tmp_bytes(int n);
void launch_expand_locate(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st);   // hit count: a.ctl->total_seeds
void launch_sort_cluster(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st);
void launch_scan_u32_to_i64(const uint32_t *in, int64_t *out, int n, void *tmp, size_t tmp_bytes, cudaStream_t st);

// nw_kernel.cu
struct NwJobDev { int64_t s1_off; int64_t gpos; int64_t op_off; int64_t flag_off; int64_t aux_off; int32_t m, n; };
constexpr int NW_BINS = 64 * 64 + 1;   // shape classes (n, m) of the thread-per-alignment kernel + one for everything larger
struct NwSorted { int64_t s1_off, gpos, op_off; int32_t job; int16_t m, n; };   // a job as the thread-per-alignment kernel reads it: 32 bytes
struct NwScratch {         // shape-class counting sort of the job queue, owned by the context
    DevBuf<uint32_t> hist, bin_start, bin_cur, order, counter;   // hist and counter are left zeroed by k_nw_bins / the kernels that use them
    DevBuf<NwSorted> sorted;
    DevBuf<uint32_t> gflags;                                     // per-warp traceback-flag scratch of the shapes whose flags do not fit on chip
};
struct NwRound {           // one batched NW launch over a job queue whose length lives on the device
    NwJobDev *jobs; const int32_t *n_jobs; int cap_jobs;
    int round;             // 0: phase B (gap flanks, with Rvec/Lvec scratch per pair), 1: phase C / the stage entry point
    int with_aux;
    int64_t cap_ops, cap_flags, cap_aux;
    uint32_t *flags; uint8_t *ops; int32_t *nops;
    int32_t *rowbuf; size_t rowbuf_per_warp;   // ints per warp: 2 * (widest job + 1)
};
constexpr int NW_LAUNCHES = 5;     // prepare, bins, scatter, k_nw_thread, k_nw
void launch_nw(const DevIndex &ix, const uint8_t *codes, const NwRound &R, BatchCtl *ctl, NwScratch &scratch, cudaStream_t st);
int nw_grid_warps();
double measure_int32_ops_per_second(cudaStream_t st);
double measure_l2_gather_bytes_per_second(cudaStream_t st, size_t table_bytes);

// kmer_kernel.cu
struct KmerJobDev { int64_t s1_off; int64_t gpos; int32_t len1, len2; };
struct KmerScratch {       // device scratch of the k-mer fast path, owned by the context
    DevBuf<uint32_t> ntiles, cap, count, recs, heavy_list;
    DevBuf<int64_t> tile_off, rec_off;
    DevBuf<uint8_t> scan_tmp;
};
constexpr int KMER_LAUNCHES = 7;   // prep, 2 scans (counted as one each), check, scan, walk, ring
// job-queue length on the device (*n_jobs <= cap_jobs); the record pool holds cap_recs records (CAP_KRECS on overflow)
void launch_kmer(const DevIndex &ix, const uint8_t *codes, const KmerJobDev *jobs, const int32_t *n_jobs, int cap_jobs, int max_len1,
                 dartgpu_kmer_hit *out, KmerScratch &scratch, BatchCtl *ctl, int64_t cap_recs, cudaStream_t st);

} // namespace dartgpu
