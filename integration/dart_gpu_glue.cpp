// dart_gpu_glue.cpp — the reference-side binding of libdartgpu.so (include/dartgpu.h).
//
// Compiled WITH the reference's sources (it includes their structure.h) and linked into their `dart` binary by
// integration/dart_gpu.patch.  What stays the reference's: the CLI (main.cpp), the FASTA/FASTQ/.gz reader
// (GetNextChunk / gzGetNextChunk, GetData.cpp:134-247), the SAM text of every record (OutputPairedAlignments /
// OutputSingledAlignments, Mapping.cpp:208-369), the htslib BAM writer, the junction table
// (UpdateGlobalSJMap / OutputSpliceJunctions, Mapping.cpp:567-577, :697-716) and the summary lines.
// What moves to the GPU: the per-read loop body of ReadMapping() (Mapping.cpp:598-640) — IdentifySeedPairs,
// GenerateAlignmentCandidate, candidate pairing / pruning, GenMappingReport, CheckPairedFinalAlignments, Set*AlignmentFlag,
// EvaluateMAPQ and the increments of UpdateLocalSJMap — through ONE call per batch, dartgpu_map_reads().
//
// A worker thread (the reference starts iThreadNum of them, Mapping.cpp:792) owns one GPU context.  Instead of mapping a
// chunk of <= 4000 reads at a time it pulls chunks under LibraryLock until DART_GPU_BATCH reads (default 262144) are
// gathered — a B200 needs batches, not chunks — maps them, copies the report fields into the reference's ReadItem_t and
// lets the reference format and write them exactly as before.
// Environment: DART_GPU_DEVICES=0,1,...  GPUs to use (thread t works on entry t mod n; default "0")
//              DART_GPU_BATCH=<reads>    reads gathered per GPU call
#include <pthread.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "structure.h"
#include "htslib/htslib/sam.h"
#include "htslib/htslib/kstring.h"
#include "dart_gpu_glue.h"
#include "dartgpu.h"

// file-scope state of the reference's Mapping.cpp that its worker threads share (Mapping.cpp:11-20)
extern FILE *sam_out;
extern samFile *bam_out;
extern bam_hdr_t *header;
extern bool bSepLibrary;
extern pthread_mutex_t LibraryLock, OutputLock;
extern FILE *ReadFileHandler1, *ReadFileHandler2;
extern gzFile gzReadFileHandler1, gzReadFileHandler2;
extern int64_t iTotalReadNum, iUniqueMapping, iUnMapping, iPaired;
extern void OutputPairedAlignments(ReadItem_t& read1, ReadItem_t& read2, int& u, int& un, int& p, vector<string>& out);
extern void OutputSingledAlignments(ReadItem_t& read, int& u, int& un, vector<string>& out);
extern void UpdateGlobalSJMap(map<pair<int64_t, int64_t>, SpliceJunction_t>& LocalSJMap);

static std::vector<int> g_devices;
static std::vector<dartgpu_ctx *> g_ctx;           // one per worker thread
static int g_next_worker = 0;
static pthread_mutex_t g_worker_lock = PTHREAD_MUTEX_INITIALIZER;
static int g_batch = 262144;
static bool g_active = false;

static dartgpu_params current_params()
{   // the hot-path globals of structure.h:182-185 as they stand NOW (bPairEnd is only known once Mapping() has looked
    // at the library, Mapping.cpp:773-787)
    dartgpu_params p; dartgpu_default_params(&p);
    p.max_gaps = MaxGaps; p.max_intron = MaxIntronSize; p.min_intron = MinIntronSize; p.max_mismatch = MaxMismatch;
    p.max_dup = MaxDupNum; p.multi_hit = bMultiHit; p.pair_end = bPairEnd; p.all_sj = bFindAllJunction; p.unique = bUnique;
    p.host_threads = 1;
    return p;
}

bool DartGpuInit()
{
    const char *e = getenv("DART_GPU_DEVICES");
    std::string list = e && *e ? e : "0";
    for (char *tok = strtok(&list[0], ","); tok; tok = strtok(NULL, ",")) g_devices.push_back(atoi(tok));
    if (const char *b = getenv("DART_GPU_BATCH")) g_batch = std::max(2 * ReadChunkSize, atoi(b));
    // the loaded index exactly as the reference holds it (bwt_t / bntseq_t / pac, structure.h:29-69)
    dartgpu_index_view v; memset(&v, 0, sizeof v);
    const bwt_t *bwt = RefIdx->bwt; const bntseq_t *bns = RefIdx->bns;
    v.primary = bwt->primary; for (int i = 0; i < 5; i++) v.L2[i] = bwt->L2[i];
    v.seq_len = bwt->seq_len; v.bwt_size = bwt->bwt_size; v.bwt = bwt->bwt;
    v.sa_intv = bwt->sa_intv; v.n_sa = bwt->n_sa; v.sa = bwt->sa;
    v.l_pac = bns->l_pac; v.pac = RefIdx->pac; v.n_seqs = bns->n_seqs;
    std::vector<int64_t> lens(bns->n_seqs); std::vector<const char *> names(bns->n_seqs);
    for (int i = 0; i < bns->n_seqs; i++) { lens[i] = bns->anns[i].len; names[i] = bns->anns[i].name; }
    v.seq_len_arr = lens.data(); v.seq_names = names.data();
    dartgpu_params p = current_params();
    g_ctx.assign(iThreadNum, (dartgpu_ctx *)NULL);
    for (int t = 0; t < iThreadNum; t++) {          // contexts of one device share ONE resident copy of the index
        int rc = dartgpu_create(&g_ctx[t], g_devices[t % g_devices.size()], &v, &p);
        if (rc != DARTGPU_OK) { fprintf(stderr, "Error! GPU context %d: %s\n", t, dartgpu_last_error(NULL)); exit(1); }
    }
    g_active = true;
    return true;
}

bool DartGpuActive() { return g_active; }

void DartGpuShutdown()
{
    for (size_t t = 0; t < g_ctx.size(); t++) dartgpu_destroy(g_ctx[t]);
    g_ctx.clear(); g_active = false;
}

static void free_reads(ReadItem_t *arr, int n)
{
    for (int i = 0; i < n; i++) {
        delete[] arr[i].header; delete[] arr[i].seq; delete[] arr[i].EncodeSeq;
        if (FastQFormat) delete[] arr[i].qual;
        delete[] arr[i].AlnReportArr;
    }
}

// maps ReadArr[0..n) (one or more whole chunks) on the GPU and fills the report fields of every ReadItem_t
static void map_on_gpu(dartgpu_ctx *ctx, ReadItem_t *ReadArr, int n, bool paired, std::vector<char> &bases, std::vector<int64_t> &off,
                       map<pair<int64_t, int64_t>, SpliceJunction_t> &LocalSJMap)
{
    off.resize(n + 1); off[0] = 0;
    for (int i = 0; i < n; i++) off[i + 1] = off[i] + ReadArr[i].rlen;
    bases.resize(off[n] + 1);
    for (int i = 0; i < n; i++) memcpy(&bases[off[i]], ReadArr[i].seq, ReadArr[i].rlen);   // mate 2 is already flipped (GetData.cpp:157-168)
    dartgpu_params p = current_params(); p.pair_end = paired;
    dartgpu_set_params(ctx, &p);
    dartgpu_reads rd; rd.n_reads = n; rd.bases = bases.data(); rd.offsets = off.data();
    dartgpu_map_result res;
    int rc = dartgpu_map_reads(ctx, &rd, &res);
    if (rc != DARTGPU_OK) { fprintf(stderr, "Error! dartgpu_map_reads: %s\n", dartgpu_last_error(ctx)); exit(1); }
    for (int i = 0; i < n; i++) {
        const dartgpu_read_result &r = res.reads[i];
        ReadItem_t &read = ReadArr[i];
        read.mapq = r.mapq; read.score = r.score; read.sub_score = r.sub_score; read.mis_num = r.mis_num;
        read.CanNum = r.n_reports; read.iBestAlnCanIdx = r.best;
        read.AlnReportArr = new AlignmentReport_t[r.n_reports];
        for (int k = 0; k < r.n_reports; k++) {
            const dartgpu_report &g = res.reports[r.report_off + k];
            AlignmentReport_t &a = read.AlnReportArr[k];
            a.AlnScore = g.aln_score; a.SJtype = g.sj_type; a.iFrag = g.flag; a.PairedAlnCanIdx = g.paired_idx;
            a.coor.bDir = g.dir != 0; a.coor.gPos = g.pos; a.coor.ChromosomeIdx = g.chr_idx;
            if (g.cigar_len > 0) a.coor.CIGAR.assign(res.cigars + g.cigar_off, g.cigar_len);
        }
    }
    for (int64_t k = 0; k < res.n_junctions; k++) {          // the increments UpdateLocalSJMap would have made
        const dartgpu_junction &j = res.junctions[k];
        map<pair<int64_t, int64_t>, SpliceJunction_t>::iterator it = LocalSJMap.find(make_pair(j.g1, j.g2));
        if (it != LocalSJMap.end()) it->second.iCount++;
        else { SpliceJunction_t sj; sj.iCount = 1; sj.type = j.type; LocalSJMap.insert(make_pair(make_pair(j.g1, j.g2), sj)); }
    }
}

void *DartGpuReadMapping(void *arg)
{
    (void)arg;
    pthread_mutex_lock(&g_worker_lock);
    dartgpu_ctx *ctx = g_ctx[g_next_worker++ % g_ctx.size()];
    pthread_mutex_unlock(&g_worker_lock);

    ReadItem_t *ReadArr = new ReadItem_t[g_batch + ReadChunkSize + 2];
    std::vector<int> chunk_begin;                       // chunk boundaries inside the batch
    std::vector<std::string> SamOutputVec;
    std::vector<char> bases; std::vector<int64_t> off;
    map<pair<int64_t, int64_t>, SpliceJunction_t> LocalSJMap;
    bool eof = false;
    while (!eof) {
        // ---- gather chunks (the reference's reader, under its lock) ----
        int n = 0;
        chunk_begin.clear();
        pthread_mutex_lock(&LibraryLock);
        while (n + ReadChunkSize <= g_batch) {
            int got = gzCompressed ? gzGetNextChunk(bSepLibrary, gzReadFileHandler1, gzReadFileHandler2, ReadArr + n)
                                   : GetNextChunk(bSepLibrary, ReadFileHandler1, ReadFileHandler2, ReadArr + n);
            if (got == 0) { eof = true; break; }
            chunk_begin.push_back(n); n += got;
            if (bPairEnd && (got & 1)) break;            // an odd chunk is mapped single-end (Mapping.cpp:598, :625): keep it apart
        }
        pthread_mutex_unlock(&LibraryLock);
        if (n == 0) break;
        chunk_begin.push_back(n);
        // ---- map: all even chunks in one call; a trailing odd chunk on its own, as single-end reads ----
        int n_even = n;
        const int last = (int)chunk_begin.size() - 2;
        const bool odd_tail = bPairEnd && ((chunk_begin[last + 1] - chunk_begin[last]) & 1);
        if (odd_tail) n_even = chunk_begin[last];
        if (n_even > 0) map_on_gpu(ctx, ReadArr, n_even, bPairEnd, bases, off, LocalSJMap);
        if (odd_tail) map_on_gpu(ctx, ReadArr + n_even, n - n_even, false, bases, off, LocalSJMap);
        // ---- format with the reference's own functions, chunk by chunk, and write under its lock ----
        int myUniqueMapping = 0, myUnMapping = 0, myPairing = 0;
        SamOutputVec.clear();
        for (int i = 0; i < n_even; i += bPairEnd ? 2 : 1) {
            if (bPairEnd) OutputPairedAlignments(ReadArr[i], ReadArr[i + 1], myUniqueMapping, myUnMapping, myPairing, SamOutputVec);
            else OutputSingledAlignments(ReadArr[i], myUniqueMapping, myUnMapping, SamOutputVec);
        }
        for (int i = n_even; i < n; i++) OutputSingledAlignments(ReadArr[i], myUniqueMapping, myUnMapping, SamOutputVec);
        pthread_mutex_lock(&OutputLock);
        iTotalReadNum += n; iUniqueMapping += myUniqueMapping; iUnMapping += myUnMapping; iPaired += myPairing;
        if (OutputFileFormat == 0) {
            for (size_t k = 0; k < SamOutputVec.size(); k++) { fputs(SamOutputVec[k].c_str(), sam_out); fputc('\n', sam_out); }
        } else {
            bam1_t *b = bam_init1();
            kstring_t str = {0, 0, NULL};
            for (size_t k = 0; k < SamOutputVec.size(); k++) {
                str.s = (char *)SamOutputVec[k].c_str(); str.l = SamOutputVec[k].length();
                if (sam_parse1(&str, header, b) >= 0) (void)sam_write1(bam_out, header, b);
            }
            bam_destroy1(b);
        }
        pthread_mutex_unlock(&OutputLock);
        free_reads(ReadArr, n);
    }
    delete[] ReadArr;
    pthread_mutex_lock(&OutputLock);
    UpdateGlobalSJMap(LocalSJMap);
    pthread_mutex_unlock(&OutputLock);
    return (void *)(1);
}
