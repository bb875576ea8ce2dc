// The per-(host thread, GPU) context behind the C-ABI, and the internal stage runners the whole-path
// orchestration (pipeline.cu) shares with the stage entry points (capi.cu).
#pragma once
#include <chrono>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "dartgpu_internal.h"

namespace dartgpu {
// Everything that is read-only after load.  One copy per device: contexts created from the same index on the same device
// (one per host thread, the C-ABI's unit of concurrency) share it, so 180 GB of HBM hold one human-sized index plus its
// full suffix array instead of one per thread.
struct SharedIndex {
    int device = 0;
    std::string key, ident;
    DevIndex ix{};
    int64_t G = 0;
    std::vector<std::string> names;
    std::vector<int64_t> chr_len, chr_fwd;
    std::vector<int64_t> ends;      // sorted ChrLocMap keys (/root/reference/src/bwt_index.cpp:249-250)
    std::vector<int> end_chr;       // ChrLocMap values
    DevBuf<uint8_t> d_occ32;        // Occ32 blocks
    DevBuf<uint8_t> d_sa;           // u32 or u64 entries
    DevBuf<KmerStart> d_ktab;       // search-start table
    DevBuf<uint32_t> d_ref2;
    DevBuf<int64_t> d_ends;
    DevBuf<char> d_chr_names; DevBuf<int32_t> d_chr_name_off;   // sequence names for the device-side SAM text
    // Batches of the contexts of this device run their KERNELS in submission order (see compute_turn_begin in capi.cu):
    // the event behind the kernels of the most recently submitted batch, and the context that owns it.
    std::mutex turn_mutex;
    static constexpr int TURN_RING = 8;
    cudaEvent_t turn_ev[TURN_RING] = {};       // events behind the kernels of the last TURN_RING batches submitted on this device
    const void *turn_owner[TURN_RING] = {};
    uint64_t turn_next = 0;                     // batches submitted so far
    ~SharedIndex();
};
} // namespace dartgpu

namespace dartgpu {
// Capacities of the per-batch device pools whose fill is only known on the device (BatchCtl, dartgpu_internal.h).
// They only ever grow: from the batch's size the first time, from the control block of an aborted batch afterwards.
struct Caps {
    int64_t seeds = 0, cands = 0, pool = 0, krecs = 0, cig = 0, text = 0, junc = 0, sam = 0;
    int64_t nw_ops[2] = {0, 0}, nw_flags = 0, nw_aux = 0;
};
// device side of dartgpu_submit_fastq (sam_kernels.cu): the raw FASTQ text of the batch and where every read's name,
// bases and qualities sit in it; the SAM text pool
struct FastqDev {
    DevBuf<uint8_t> text;
    DevBuf<int64_t> ls1, ls2, off1, off2, seq_pos, name_pos, qual_pos, unit_off;
    DevBuf<uint32_t> cnt1, cnt2, unit_bytes;
    DevBuf<int32_t> name_len, qual_len;
    DevBuf<char> sam;
    PinBuf<char> h_sam;
    int64_t sent_sam = 0, last_sam = -1;
    int rlen_seen = 0;                      // longest read seen so far on this context (sizes scratch; verified on the device)
};
} // namespace dartgpu

struct dartgpu_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, own_stream = nullptr;
    dartgpu_params prm{};
    std::string err;

    // ---- the index: device tables + the sequence table, shared by every context on the same device ----
    std::shared_ptr<dartgpu::SharedIndex> shared;
    dartgpu::DevIndex ix{};
    int64_t G = 0;

    // ---- read batch on the device ----
    int n_reads = 0, max_rlen = 0, cap_rec = 0;
    int64_t n_code_bytes = 0;
    dartgpu::PinBuf<uint8_t> h_raw;          // the caller's bases, staged for DMA
    dartgpu::PinBuf<int64_t> h_off;
    dartgpu::DevBuf<uint8_t> d_raw;
    dartgpu::DevBuf<int64_t> d_off;
    dartgpu::DevBuf<uint32_t> d_padded;
    dartgpu::DevBuf<uint8_t> d_codes;
    dartgpu::DevBuf<uint2> d_packed;         // the search kernel's 2-bit view of the batch
    dartgpu::DevBuf<int64_t> d_dev_off;
    dartgpu::DevBuf<int32_t> d_rlen;

    // ---- seeding buffers ----
    dartgpu::DevBuf<dartgpu::SearchRec> d_recs;
    dartgpu::DevBuf<uint32_t> d_nrec, d_nhits, d_ncand, d_meta, d_big_list, d_mid_list;
    dartgpu::DevBuf<int64_t> d_seed_off;
    dartgpu::DevBuf<uint64_t> d_keys, d_big_scratch;
    dartgpu::DevBuf<int32_t> d_cand_begin, d_cand_count, d_cand_score;
    dartgpu::DevBuf<uint8_t> d_scan_tmp;
    dartgpu::DevBuf<dartgpu::BatchCtl> d_ctl;     // the batch's control block on the device ...
    dartgpu::PinBuf<dartgpu::BatchCtl> h_ctl;     // ... and its copy, read once per batch
    dartgpu::Caps caps;
    int64_t total_seeds = 0;
    // ---- the batch in flight (dartgpu_submit .. dartgpu_wait) ----
    bool in_flight = false, whole_path = false, timed_upload = false;
    bool whole_path_enqueue = false;              // set around enqueue_seeding by the whole-path calls: k_search runs without work counters
    bool results_on_device = false;               // dartgpu_set_result_location
    bool from_fastq = false, emit_sam = false;    // the batch came as FASTQ text / leaves as SAM text (dartgpu_submit_fastq)
    dartgpu::FastqDev fq;
    int attempts = 0;
    cudaEvent_t done = nullptr;                   // blocking-sync event recorded behind the batch
    cudaEvent_t compute_done = nullptr;           // recorded behind the batch's last kernel, before its result copies
    double t_submit_ms = 0;

    // seeding results on the host
    dartgpu::PinBuf<int64_t> h_seed_off;
    dartgpu::PinBuf<uint64_t> h_keys;
    dartgpu::PinBuf<int32_t> h_cand_begin, h_cand_count, h_cand_score;
    dartgpu::PinBuf<uint32_t> h_ncand;
    std::vector<int64_t> o_seed_gpos, o_cand_off;
    std::vector<int32_t> o_seed_rpos, o_seed_len, o_cand_begin, o_cand_count, o_cand_score;

    // ---- k-mer / NW job buffers ----
    dartgpu::DevBuf<uint8_t> d_job_codes;         // fragment bases of the stage entry points
    dartgpu::PinBuf<uint8_t> h_job_codes;
    dartgpu::PinBuf<dartgpu::KmerJobDev> h_kjobs;
    dartgpu::DevBuf<dartgpu::KmerJobDev> d_kjobs;
    dartgpu::DevBuf<dartgpu_kmer_hit> d_khits;
    dartgpu::KmerScratch kscratch;
    dartgpu::NwScratch nwscratch;
    dartgpu::PinBuf<dartgpu_kmer_hit> h_khits;
    dartgpu::PinBuf<dartgpu::NwJobDev> h_njobs;
    dartgpu::DevBuf<dartgpu::NwJobDev> d_njobs;
    dartgpu::DevBuf<uint32_t> d_nw_flags;
    dartgpu::DevBuf<int32_t> d_nw_rowbuf, d_nw_nops;
    dartgpu::DevBuf<uint8_t> d_nw_ops;
    dartgpu::PinBuf<uint8_t> h_nw_ops;
    dartgpu::PinBuf<int32_t> h_nw_nops;
    std::vector<int64_t> o_op_off;
    std::vector<uint8_t> o_ops;

    // ---- whole-path results ----
    std::vector<dartgpu_read_result> o_reads;
    std::vector<dartgpu_report> o_reports;
    std::vector<char> o_cigars;
    std::vector<dartgpu_junction> o_junctions;

    // ---- device orchestration buffers (report_kernels.cu) ----
    void *dpipe = nullptr;

    // ---- measurement ----
    dartgpu_stats stats{};
    cudaEvent_t ev[20] = {};
};

namespace dartgpu {

void stats_begin(dartgpu_ctx *c);
void add_ms(dartgpu_ctx *c, double *slot, cudaEvent_t a, cudaEvent_t b);

// stage runners; all throw CudaError / std::exception on failure
void upload_reads(dartgpu_ctx *c, const dartgpu_reads *reads);          // encode + H2D
void enqueue_seeding(dartgpu_ctx *c);                                     // search, locate, sort+cluster: enqueued, nothing waits
// k-mer jobs whose fragments live in `codes_dev` (device). Results in c->h_khits (valid after return).
void run_kmer(dartgpu_ctx *c, const uint8_t *codes_dev, const KmerJobDev *jobs, int n_jobs, int max_len1);
// NW jobs (op_off / flag_off are assigned on the device). Results: c->o_op_off / c->o_ops (compacted, left-to-right columns).
void run_nw(dartgpu_ctx *c, const uint8_t *codes_dev, NwJobDev *jobs, int n_jobs);
// the whole per-read path over the uploaded batch: device orchestration, enqueued behind the seeding kernels
void enqueue_pipeline(dartgpu_ctx *c);
// sam_kernels.cu
void upload_fastq(dartgpu_ctx *c, const dartgpu_fastq_block *b);          // H2D of the raw text + parse + encode on the device
void enqueue_sam(dartgpu_ctx *c, const dartgpu_read_result *rr, const dartgpu_report *rep, const char *cigars);
// Kernels of different contexts of one device take turns (FIFO) instead of time-sharing the SMs: see capi.cu
void compute_turn_begin(dartgpu_ctx *c);
void compute_turn_end(dartgpu_ctx *c);
void finish_pipeline(dartgpu_ctx *c, dartgpu_map_result *out);          // after the batch's synchronisation
void finish_sam(dartgpu_ctx *c, dartgpu_sam_result *out);
void free_device_pipe(void *p);

struct Timer {
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double ms() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); }
};

} // namespace dartgpu
