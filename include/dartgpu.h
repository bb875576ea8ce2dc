/* dartgpu.h — the C-ABI boundary of the B200 mapping hot path (libdartgpu.so).
 *
 * The reference has no plugin/FFI interface for this path: the boundary is the set of C++ free
 * functions declared in /root/reference/src/structure.h:192-233 and called from ReadMapping()
 * (/root/reference/src/Mapping.cpp:600-639).  Each entry point below replaces one of those call sites,
 * batched over many reads (the reference's own chunk is <=4000 reads, src/GetData.cpp:176 — far too
 * small for a B200, so the caller aggregates chunks).  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns DARTGPU_OK (0) or a negative DARTGPU_ERR_* code; dartgpu_last_error(ctx)
 *     has the text.  There is NO CPU fallback: without a usable CUDA device every call fails.
 *   - a context = one batch in flight on one GPU: its own CUDA stream, device pools and pinned result buffers.
 *     Contexts created from the same index on the same device share ONE resident copy of the index tables; a host
 *     thread may drive several contexts (dartgpu_submit on each, then dartgpu_wait on each) so that the copies and
 *     kernels of consecutive batches overlap.  A context must not be used from two threads at once.
 *   - result buffers are owned by the context (pinned host memory) and stay valid until the next call
 *     on the same context.
 *   - genome coordinates are the reference's: [0,G) forward strand, [G,2G) reverse complement
 *     (src/bwt_index.cpp:234, src/AlignmentCandidates.cpp:88-108); int64 everywhere.
 *   - read bases arrive as the ASCII the reference keeps in ReadItem_t.seq (src/structure.h:149-164);
 *     mate 2 already reverse-complemented as GetNextChunk leaves it (src/GetData.cpp:157-168).
 */
#ifndef DARTGPU_H
#define DARTGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DARTGPU_OK                 0
#define DARTGPU_ERR_NO_DEVICE     -1   /* no CUDA device / driver: the product path refuses to run */
#define DARTGPU_ERR_CUDA          -2   /* a CUDA call or kernel failed */
#define DARTGPU_ERR_INDEX         -3   /* index files missing or inconsistent */
#define DARTGPU_ERR_ARG           -4   /* bad argument (NULL, negative size, ...) */
#define DARTGPU_ERR_READ_TOO_LONG -5   /* a read exceeds DARTGPU_MAX_RLEN */
#define DARTGPU_ERR_NOMEM         -6
#define DARTGPU_MAX_RLEN        1024   /* the reference's .gz reader caps lines at 1023 chars (src/GetData.cpp:186) */

typedef struct dartgpu_ctx dartgpu_ctx;

/* The hot-path globals of the reference (src/structure.h:182-185) as one POD; defaults = src/main.cpp:101-117. */
typedef struct {
    int32_t  max_gaps;      /* MaxGaps        = 5                                   */
    int32_t  max_intron;    /* MaxIntronSize  = 500000 (flag parser clamps >=100000) */
    int32_t  min_intron;    /* MinIntronSize  = 5                                   */
    int32_t  max_mismatch;  /* MaxMismatch    = 0  (-mis)             */
    uint32_t max_dup;       /* MaxDupNum      = 100 (-max_dup, clamped 100..10000)  */
    int32_t  multi_hit;     /* bMultiHit (-m)                                        */
    int32_t  pair_end;      /* bPairEnd: reads 2i,2i+1 are mates                      */
    int32_t  all_sj;        /* bFindAllJunction (-all_sj)                             */
    int32_t  unique;        /* bUnique (-unique)                                      */
    int32_t  host_threads;  /* OpenMP threads that stage PAGEABLE caller buffers into pinned memory; 0 = all cores */
} dartgpu_params;

void dartgpu_default_params(dartgpu_params *p);

/* The loaded index exactly as the reference holds it after bwa_idx_load()+RestoreReferenceInfo()
 * (bwt_t / bntseq_t, src/structure.h:29-69): what a patched Mapping.cpp passes once at start-up. */
typedef struct {
    uint64_t        primary;      /* bwt_t.primary                                  */
    uint64_t        L2[5];        /* bwt_t.L2                                       */
    uint64_t        seq_len;      /* bwt_t.seq_len (= 2G)                           */
    uint64_t        bwt_size;     /* bwt_t.bwt_size, in 32-bit words                */
    const uint32_t *bwt;          /* bwt_t.bwt: 64-byte blocks, 4 x u64 Occ + 8 x u32 symbols */
    uint64_t        sa_intv;      /* bwt_t.sa_intv (32)                             */
    uint64_t        n_sa;         /* bwt_t.n_sa                                     */
    const uint64_t *sa;           /* bwt_t.sa, sa[0] = (uint64_t)-1                 */
    int64_t         l_pac;        /* bntseq_t.l_pac = G                             */
    const uint8_t  *pac;          /* bwaidx_t.pac: 2-bit, forward strand, l_pac/4+1 bytes */
    int32_t         n_seqs;       /* bntseq_t.n_seqs                                */
    const int64_t  *seq_len_arr;  /* bntann1_t.len for each sequence                */
    const char *const *seq_names; /* bntann1_t.name                                 */
} dartgpu_index_view;

/* Replaces: the once-per-run index hand-over (src/main.cpp:220-223).  The tables are re-laid-out on the GPU (one-sector
 * Occ32 blocks: 4 x u32 counts + two 64-bit planes per 64 BWT symbols; a densified suffix array; a search-start table;
 * the 2-bit reference over both strands) and stay resident in HBM, one copy per device.  Texts of 2^33 symbols or more
 * (genomes over 4.29 Gbp) are rejected with DARTGPU_ERR_INDEX: a seed's coordinate has 33 bits. */
int  dartgpu_create(dartgpu_ctx **out, int device, const dartgpu_index_view *idx, const dartgpu_params *p);
/* Convenience: read <prefix>.bwt/.sa/.pac/.ann written by bwt_index / `dart index` / bwa index
 * (formats: src/bwt_index.cpp:15-35, :37-89, :102-121) and call dartgpu_create. */
int  dartgpu_create_from_files(dartgpu_ctx **out, int device, const char *prefix, const dartgpu_params *p);
/* Replaces: the BWT / Occ / SA construction of `dart index` / bwt_index (bwa_idx_build steps 2-5,
 * src/BWT_Index/bwtindex.c:96-144: bwt_bwtgen2, bwt_bwtupdate_core, bwt_cal_sa(32) and the two dumps, src/BWT_Index/bwt.c:174-196).
 * pac = the forward strand, 2 bits per base, first base in the top bits (what bns_fasta2bntseq leaves in <prefix>.pac,
 * src/BWT_Index/bntseq.c:158-211), l_pac bases.  Writes <prefix>.bwt and <prefix>.sa byte-identical to the reference's
 * (the BWT of a text is unique).  max_suffixes_per_pass = 0 sizes the sort passes from the free HBM. */
int  dartgpu_index_build(int device, const uint8_t *pac, int64_t l_pac, const char *prefix, uint64_t max_suffixes_per_pass);
void dartgpu_destroy(dartgpu_ctx *ctx);
int  dartgpu_set_params(dartgpu_ctx *ctx, const dartgpu_params *p);
const char *dartgpu_last_error(const dartgpu_ctx *ctx);   /* ctx may be NULL: error of the last failed create */
int64_t dartgpu_genome_size(const dartgpu_ctx *ctx);
int  dartgpu_num_sequences(const dartgpu_ctx *ctx);
const char *dartgpu_sequence_name(const dartgpu_ctx *ctx, int i);
int64_t dartgpu_sequence_length(const dartgpu_ctx *ctx, int i);
/* Where dartgpu_wait / dartgpu_map_reads leave the records: 0 (default) = the context's page-locked host buffers;
 * 1 = device memory: the pointers of dartgpu_map_result are then DEVICE pointers and nothing crosses PCIe (for callers
 * that consume the records on the GPU, and for device-resident timing: bench.py `value`). */
int  dartgpu_set_result_location(dartgpu_ctx *ctx, int on_device);
/* Pins the calling host thread to the CPUs of `device`'s NUMA node (call it before creating the thread's contexts, so that
 * their page-locked buffers are allocated next to the GPU). */
int  dartgpu_bind_host_thread(int device);
/* Work on `stream` (a cudaStream_t) instead of the context's own stream, e.g. the caller's current stream. */
int  dartgpu_set_stream(dartgpu_ctx *ctx, void *cuda_stream);

/* ---- a batch of reads ------------------------------------------------------------------------------- */
typedef struct {
    int32_t        n_reads;
    const char    *bases;     /* concatenated ReadItem_t.seq, no terminators                     */
    const int64_t *offsets;   /* n_reads+1 offsets into bases; rlen = offsets[i+1]-offsets[i]     */
} dartgpu_reads;

/* ---- stage 1: IdentifySeedPairs + GenerateAlignmentCandidate ------------------------------------------
 * Replaces, for every read of the batch: IdentifySeedPairs(rlen, EncodeSeq) (src/AlignmentCandidates.cpp:181-215,
 * with BWT_Search/bwt_sa, src/bwt_search.cpp:127-182) and GenerateAlignmentCandidate(rlen, seeds)
 * (src/AlignmentCandidates.cpp:241-288).
 * Seeds of read i are seeds[seed_off[i] .. seed_off[i+1]), sorted by (gPos,rPos) as the reference sorts them.
 * Candidate c of read i covers seeds seed_off[i]+cand_begin[k] .. +cand_count[k], k = cand_off[i]+c; its
 * PosDiff is max(seed.gPos-seed.rPos, 0) of its first seed. */
typedef struct {
    const int64_t *seed_off;    /* n_reads+1 */
    const int64_t *seed_gpos;   /* SeedPair_t.gPos */
    const int32_t *seed_rpos;   /* SeedPair_t.rPos */
    const int32_t *seed_len;    /* SeedPair_t.rLen = gLen */
    const int64_t *cand_off;    /* n_reads+1 */
    const int32_t *cand_begin;  /* first seed, relative to the read's first seed */
    const int32_t *cand_count;  /* number of seeds */
    const int32_t *cand_score;  /* AlignmentCandidate_t.Score */
} dartgpu_seeds;

int dartgpu_seed_and_cluster(dartgpu_ctx *ctx, const dartgpu_reads *reads, dartgpu_seeds *out);

/* ---- stage 2: 8-mer re-seeding inside a (read gap, genome window) --------------------------------------
 * Replaces GenerateLongestSimplePairsFromFragmentPair(len1, frag1, len2, frag2) (src/KmerAnalysis.cpp:134-166)
 * as called by ReseedingWithSpecificRegion (src/AlignmentCandidates.cpp:596-624): frag1 = read bases
 * [frag_off, frag_off+frag_len) of `bases`, frag2 = RefSequence[gpos, gpos+glen).
 * Result: SeedPair_t {rPos, gPos, rLen} in fragment-local coordinates; rLen = 0 when nothing was found. */
typedef struct { int64_t frag_off; int32_t frag_len; int32_t glen; int64_t gpos; } dartgpu_kmer_job;
typedef struct { int32_t rpos; int32_t gpos; int32_t len; } dartgpu_kmer_hit;

int dartgpu_kmer_reseed(dartgpu_ctx *ctx, const char *bases, int64_t n_bases,
                        const dartgpu_kmer_job *jobs, int32_t n_jobs, const dartgpu_kmer_hit **out);

/* ---- stage 3: Needleman-Wunsch gap fill ----------------------------------------------------------------
 * Replaces nw_alignment(m, s1, n, s2) (src/nw_alignment.cpp:18-82) at its call sites
 * (src/AlignmentCandidates.cpp:395, :420; src/tools.cpp:156, :220, :268): s1 = read bases
 * [frag_off, frag_off+m) of `bases`, s2 = RefSequence[gpos, gpos+n).
 * Result per job: the alignment columns left to right, one byte each:
 *   0 = s1 and s2 both advance, 1 = '-' inserted into s1 (s2 advances), 2 = '-' inserted into s2.
 * Columns of job j are ops[op_off[j] .. op_off[j+1]). */
typedef struct { int64_t frag_off; int32_t m; int32_t n; int64_t gpos; } dartgpu_nw_job;
typedef struct { const int64_t *op_off; const uint8_t *ops; } dartgpu_nw_result;

int dartgpu_nw_align(dartgpu_ctx *ctx, const char *bases, int64_t n_bases,
                     const dartgpu_nw_job *jobs, int32_t n_jobs, dartgpu_nw_result *out);

/* ---- the whole per-read path -----------------------------------------------------------------------------
 * Replaces the body of the per-read loop of ReadMapping() (src/Mapping.cpp:600-639): seeds, candidates, mate
 * pairing of candidates, GenMappingReport (src/AlignmentCandidates.cpp:1079-1207), CheckPairedFinalAlignments,
 * Set*AlignmentFlag, EvaluateMAPQ and the junctions UpdateLocalSJMap would record.  The caller keeps its own
 * reader and its own OutputPaired/SingledAlignments + OutputSpliceJunctions.
 * With params.pair_end the batch holds mates at 2i, 2i+1 (n_reads even). */
/* The two result records are what crosses PCIe for every read (56 bytes per read with one report, 80 before): the fields are the
 * reference's, the types as narrow as their ranges allow (scores and lengths are bounded by DARTGPU_MAX_RLEN). */
typedef struct {            /* = the report fields of ReadItem_t (src/structure.h:156-163); 24 bytes */
    int64_t report_off;     /* first AlignmentReport of this read in reports[] */
    int32_t n_reports;      /* CanNum */
    int32_t best;           /* iBestAlnCanIdx */
    int16_t score, sub_score, mis_num;
    uint8_t mapq;
    uint8_t reserved;
} dartgpu_read_result;

typedef struct {            /* = AlignmentReport_t + Coordinate_t (src/structure.h:117-141); 32 bytes */
    int64_t pos;            /* coor.gPos (1-based on the chromosome) */
    int32_t cigar_off;      /* coor.CIGAR = cigars[cigar_off .. cigar_off+cigar_len) */
    int32_t flag;           /* iFrag (SAM FLAG); meaningful where the reference assigns it */
    int32_t paired_idx;     /* PairedAlnCanIdx */
    int32_t chr_idx;        /* coor.ChromosomeIdx */
    int16_t aln_score;      /* AlnScore */
    int16_t cigar_len;
    int8_t  sj_type;        /* SJtype, -1 = none */
    uint8_t dir;            /* coor.bDir (1 forward); valid when aln_score > 0 */
    uint8_t reserved[2];
} dartgpu_report;

typedef struct {            /* one UpdateLocalSJMap increment (src/Mapping.cpp:532-565) */
    int64_t g1, g2;         /* absolute coordinates as the map keys them */
    int32_t type;           /* SJtype of the alignment */
    int32_t read;           /* read index in the batch */
} dartgpu_junction;

/* reads[] and reports[] are in input order.  The slices of cigars[] and junctions[] are laid out in the order the GPU's warps
 * finished (every report carries its cigar_off, every junction record its read): contents are deterministic, pool layouts
 * are not. */
typedef struct {
    const dartgpu_read_result *reads;   int32_t n_reads;
    const dartgpu_report      *reports; int64_t n_reports;
    const char                *cigars;  int64_t n_cigar_bytes;
    const dartgpu_junction    *junctions; int64_t n_junctions;
} dartgpu_map_result;

int dartgpu_map_reads(dartgpu_ctx *ctx, const dartgpu_reads *reads, dartgpu_map_result *out);

/* The same in two halves, so that ONE host thread keeps several batches in flight per GPU (one per context) — what the
 * reference does with iThreadNum pthreads pulling chunks under LibraryLock (src/Mapping.cpp:591-595, :792-793).
 * dartgpu_submit stages the reads (page-locked caller buffers are handed to the DMA engine as they are), enqueues every
 * kernel of the batch and the copies of the results on the context's stream and returns without waiting for the GPU;
 * `reads` may be reused as soon as it returns unless it is page-locked (then: after dartgpu_wait).
 * dartgpu_wait sleeps (no spinning) until the batch is done — the batch's only host synchronisation in steady state —
 * and hands out the results.  dartgpu_map_reads = submit + wait.  At most one batch per context is in flight. */
int dartgpu_submit(dartgpu_ctx *ctx, const dartgpu_reads *reads);
int dartgpu_wait(dartgpu_ctx *ctx, dartgpu_map_result *out);

/* ---- read ingest and SAM text on the device ----------------------------------------------
 * Replaces, for FASTQ input, GetNextChunk (src/GetData.cpp:134-179: record parsing, header cut, nst_nt4_table encoding,
 * mate-2 reverse complement + quality reversal) and OutputPairedAlignments / OutputSingledAlignments
 * (src/Mapping.cpp:208-369: the SAM text of every record): the caller hands over raw blocks of the FASTQ file(s) and gets
 * back complete SAM lines in input order plus the junction records — the host only reads and writes files.
 * text1 must start at a record and end with the '\n' that ends a record (dartgpu_fastq_cut finds that point); with two
 * files (-f / -f2) text2 holds the same number of records and read 2i is record i of text1, read 2i+1 record i of text2;
 * with one file and params.pair_end (-p) mates alternate in text1.  Page-locked buffers are DMA'd as they are. */
typedef struct {
    const char *text1; int64_t len1;
    const char *text2; int64_t len2;    /* NULL / 0: one file */
    int32_t     n_records;              /* records in text1 (and in text2) */
    int32_t     fastq;                  /* FastQFormat; FASTA input is not handled on this path (DARTGPU_ERR_ARG) */
    int32_t     max_read_len;           /* hint: longest read (0 = take the first record's); a longer read costs one retry */
    int32_t     reserved;
} dartgpu_fastq_block;

typedef struct {
    const char *sam; int64_t n_bytes;   /* SAM lines of the batch, input order, each ending in '\n' */
    int64_t n_reads, n_unmapped, n_unique, n_paired;   /* the counters behind the reference's summary (Mapping.cpp:803-813) */
    const dartgpu_junction *junctions; int64_t n_junctions;
} dartgpu_sam_result;

int dartgpu_submit_fastq(dartgpu_ctx *ctx, const dartgpu_fastq_block *block);
int dartgpu_wait_sam(dartgpu_ctx *ctx, dartgpu_sam_result *out);
/* Page-locked host memory for the blocks (so that a caller that links nothing but this library can have its file reads
 * DMA'd without a staging copy). */
void *dartgpu_alloc_pinned(uint64_t bytes);
void  dartgpu_free_pinned(void *p);
/* Host helper: how many bytes of `text` hold at most max_records complete 4-line records (0 = no limit)? */
int64_t dartgpu_fastq_cut(const char *text, int64_t len, int32_t max_records, int32_t *n_records);

/* ---- measurement ------------------------------------------------------------------------------------------
 * Device time (CUDA events on the context's stream) and algorithmic work of the kernels launched by the LAST
 * call, for roofline reporting. */
typedef struct {
    double   ms_search, ms_locate, ms_sort_cluster, ms_kmer, ms_nw, ms_h2d, ms_d2h, ms_total_device;
    double   ms_host;                 /* wall time of the whole call on the host */
    double   ms_report;               /* device orchestration kernels (candidate pairing, repair phases, records) */
    uint64_t kernel_launches;
    uint64_t ext_steps, ext_blocks;   /* forward-extension steps and the 64-byte Occ blocks they touched */
    uint64_t lf_steps, hits, seeds;   /* LF-mapping steps (one block each), SA reads, seeds written */
    uint64_t read_bases;
    uint64_t nw_jobs, nw_cells;
    uint64_t kmer_jobs, kmer_window_bases, kmer_read_bases;
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t search_sector_loads;     /* 32-byte Occ / start-table sectors the search kernel really requested */
    double   ms_submit;               /* host time of dartgpu_submit (staging + enqueueing the batch), part of ms_host */
} dartgpu_stats;

int dartgpu_get_stats(const dartgpu_ctx *ctx, dartgpu_stats *out);
/* INT32 add/max operations per second of the context's GPU, measured with a saturating microbenchmark: the denominator
 * of the NW kernels' integer roofline. */
int dartgpu_measure_int32_peak(dartgpu_ctx *ctx, double *ops_per_second);
/* Bytes per second of independent 32-byte (one sector) gathers at random addresses of a table of `table_bytes` that fits
 * L2 — the access pattern of a rank query: the roof of the search kernel on an L2-resident index. */
int dartgpu_measure_l2_peak(dartgpu_ctx *ctx, uint64_t table_bytes, double *bytes_per_second);

/* Device-resident variant of stage 1 for kernel-only timing: upload once, run the seeding kernels many times
 * without any host<->device copy in between (bench.py `value`). */
int dartgpu_upload_reads(dartgpu_ctx *ctx, const dartgpu_reads *reads);
int dartgpu_seed_and_cluster_resident(dartgpu_ctx *ctx);   /* results stay on the device */
/* dartgpu_map_reads over the batch previously uploaded with dartgpu_upload_reads (same `reads`): no read H2D. */
int dartgpu_map_reads_resident(dartgpu_ctx *ctx, const dartgpu_reads *reads, dartgpu_map_result *out);
int dartgpu_submit_resident(dartgpu_ctx *ctx);            /* dartgpu_submit over the uploaded batch; dartgpu_wait completes it */
int dartgpu_synchronize(dartgpu_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* DARTGPU_H */
