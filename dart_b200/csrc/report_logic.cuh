// The per-candidate repair pipeline of GenMappingReport and the per-pair logic around it, phase-split and written
// over raw arrays so the SAME code runs as device code (report_kernels.cu: one thread per candidate / per pair) and,
// compiled for the host by tests/host/logic_harness.cpp, on a CPU-only box where it is checked against the reference
// before any GPU time is spent.  Nothing here allocates; every candidate owns a fixed-capacity slice of a seed pool.
//
// Reference functions restated here (all in /root/reference/src):
//   Mapping.cpp:371-477            RemoveRedundantCandidates, CheckPairedAlignmentCandidates, RemoveUnMatedAlignmentCandidates
//   AlignmentCandidates.cpp:817-902   RemoveTandemRepeatSeeds, RemoveTranslocatedSeeds
//   AlignmentCandidates.cpp:685-700, :596-624   IdentifyMissingSeeds, ReseedingWithSpecificRegion
//   AlignmentCandidates.cpp:577-594, :547-575, :385-467   SeedExtension, FillGapsBetweenAdjacentSeeds, IdentifyBestGappedPartition
//   AlignmentCandidates.cpp:702-815   CheckSeqFragment, IdentifySpliceJunction, CheckSpliceJunction
//   AlignmentCandidates.cpp:904-1035  CheckSeedOverlapping, CheckOverlappingSeeds, IdentifyNormalPairs
//   AlignmentCandidates.cpp:136-163, :1052-1064, :83-116, :37-61   CheckCoordinateValidity, CheckMinIntronSize, GenCoordinateInfo, GenerateCIGAR
//   AlignmentCandidates.cpp:1079-1207 GenMappingReport
//   tools.cpp:49-104, :130-300     AddNewCigarElements, ProcessNormal/Head/TailSequencePair, CheckLocalAlignmentQuality
//   Mapping.cpp:74-206, :479-565   Set*AlignmentFlag, EvaluateMAPQ, CheckPairedFinalAlignments, UpdateLocalSJMap
#pragma once
#include <stdint.h>

#include "dartgpu_internal.h"

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define HDN __host__ __device__
#else
#define HD inline
#define HDN inline
#endif

namespace dartgpu {

struct RSeed {            // SeedPair_t (structure.h:106-115) + the id of the batched job attached to it
    int64_t gPos, PosDiff;
    int32_t rPos, rLen, gLen;
    int32_t job;          // phase A: k-mer job between seed i-1 and i; phase B: first of two NW jobs; phase C: pair NW job
    uint8_t simple, acceptor, pad0, pad1;
    int32_t pad2;
};
static_assert(sizeof(RSeed) == 40, "RSeed layout");

struct CandState {        // AlignmentCandidate_t + AlignmentReport_t of one candidate
    int64_t PosDiff;
    int64_t pos;          // coor.gPos
    int64_t sv_off;       // slice of the seed pool
    int64_t cig_off;      // slice of the CIGAR pair pool
    int64_t text_off;     // CIGAR text
    int32_t read;
    int32_t seed_begin, seed_count;   // run inside the read's sorted seed keys, relative to the read's first seed (Env::seed_off)
    int32_t Score, PairedIdx, SJtype;
    int32_t sv_n, sv_cap;
    int32_t cig_cap, cig_n, text_len;
    int32_t AlnScore, mis, chr;
    int32_t n_ext;        // gaps sent to NW by phase B (bounds the seeds phase C can add)
    uint8_t live, skip, dir, pad;
};

struct ReadOut {          // per read scratch of the final pass
    int32_t score, sub_score, mis_num, mapq, best, n_reports;
};

struct PhaseParams {
    int max_gaps, max_intron, min_intron, max_mismatch, multi_hit, pair_end, all_sj;
};

// reference bases over [0,2G): device = 2-bit packed words, host = .pac
struct RefView {
    const uint32_t *ref2;  // device layout (16 bases per word, first base in the top bits), or nullptr
    const uint8_t *pac;    // host layout (forward strand, 4 bases per byte), or nullptr
    int64_t G;
    HD int code(int64_t p) const
    {
        if (p < 0 || p >= 2 * G) return 0;
        if (ref2) return (int)((ref2[p >> 4] >> (30 - 2 * (int)(p & 15))) & 3);
        if (p < G) return (pac[p >> 2] >> ((~p & 3) << 1)) & 3;
        int64_t q = 2 * G - 1 - p;
        return 3 - ((pac[q >> 2] >> ((~q & 3) << 1)) & 3);
    }
    // 32 bases starting at p as one u64, base k at bits 62-2k; positions outside [0,2G) read as 0 like code().
    // Round-1 ncu (k_phase<2>, full-size config[2]): a third of the phase's instructions and most of its stalls were
    // single-base code() calls — a dependent load plus two range checks per base.
    HD uint64_t window32(int64_t p) const
    {
        if (ref2 && p >= 0 && p + 32 <= 2 * G) {         // the device table has guard words behind 2G
            const uint32_t *w = ref2 + (p >> 4);
            const int sh = 2 * (int)(p & 15);
            const uint64_t hi = (uint64_t)w[0] << 32 | w[1];
            return sh ? (hi << sh) | ((uint64_t)w[2] >> (32 - sh)) : hi;
        }
        uint64_t x = 0;
        for (int k = 0; k < 32; k++) x |= (uint64_t)code(p + k) << (62 - 2 * k);
        return x;
    }
};

// bases [at, at+n) of a window32 value, right-aligned (n <= 31)
HD uint64_t win_bases(uint64_t w, int at, int n) { return n <= 0 ? 0 : (w << (2 * at)) >> (64 - 2 * n); }

// Sequential reader: keeps the 16-base word under the cursor in a register (gapped_partition, simple_enough, the
// CIGAR builders walk the genome base by base).
struct RefCursor {
    const RefView &rv; int64_t wi; uint32_t w;
    HD explicit RefCursor(const RefView &r) : rv(r), wi(-1), w(0) {}
    HD int code(int64_t p)
    {
        if (!rv.ref2 || p < 0 || p >= 2 * rv.G) return rv.code(p);
        const int64_t i = p >> 4;
        if (i != wi) { wi = i; w = rv.ref2[i]; }
        return (int)((w >> (30 - 2 * (int)(p & 15))) & 3);
    }
};

// read codes: 0..3 = ACGT, 8..11 = acgt, 4 = other, 5 = 'N'.  The reference compares raw characters against the
// upper-case genome in several places (tools.cpp:44, AlignmentCandidates.cpp:405,430): only upper-case ACGT can be equal.
HD bool raw_eq(uint8_t read_code, int gcode) { return read_code < 4 && (int)read_code == gcode; }

struct Env {
    PhaseParams P;
    RefView ref;
    int64_t G;
    const int64_t *ends; const int32_t *end_chr; int n_ends;   // sorted ChrLocMap
    const int64_t *chr_fwd;                                     // ChromosomeVec[i].FowardLocation
    const uint8_t *codes; const int64_t *code_off; const int32_t *rlen;
    const uint64_t *keys; const int64_t *seed_off;              // sorted seed keys of the batch; first seed of every read
    BatchCtl *ctl;                                              // device kernels only: counts / abort / error flags of the batch
    CandState *cs; RSeed *pool;
    // job queues
    KmerJobDev *kjobs; int32_t *kjob_count; const dartgpu_kmer_hit *khits;
    NwJobDev *njobs; int32_t *njob_count;       // queue of the phase that runs (B: gap pairs, C: non-simple pairs)
    NwJobDev *njobs_c; int32_t *njob_count_c;    // phase C's queue, for candidates that run A-B-C in one go
    const uint8_t *ops; const int32_t *nops;     // NW results of the round being consumed (right-aligned per job)
    const NwJobDev *done_jobs;                   // the jobs those results belong to
    int32_t *xscratch;                           // Rvec/Lvec scratch of the gap-extension jobs
    int32_t *cig;                                // CIGAR pair pool: len << 8 | op
};

// Claim n consecutive job slots (n is the same constant at every call site).  On the device the claims of the lanes
// that reach the call together are merged into one atomic (a million single-address atomics per batch otherwise:
// ~0.5 ms on B200); slot order is arbitrary either way and no result depends on it.
HD int alloc_slots(int32_t *counter, int n)
{
#if defined(__CUDA_ARCH__)
    const unsigned active = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(active) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, n * __popc(active));
    base = __shfl_sync(active, base, leader);
    return base + n * __popc(active & ((1u << lane) - 1u));
#else
    int v = *counter; *counter += n; return v;
#endif
}

HD int chr_lookup(const Env &E, int64_t g, int64_t *end_out)
{   // ChrLocMap.lower_bound(g)
    int lo = 0, hi = E.n_ends;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (E.ends[mid] < g) lo = mid + 1; else hi = mid; }
    if (lo >= E.n_ends) lo = E.n_ends - 1;
    if (end_out) *end_out = E.ends[lo];
    return E.end_chr[lo];
}

HD bool seed_less(const RSeed &a, const RSeed &b) { return a.gPos == b.gPos ? a.rPos < b.rPos : a.gPos < b.gPos; }

// stable insertion sort by (gPos,rPos) — the order of CompByGenomePos; keys are unique on live data
HDN void sort_seeds(RSeed *sv, int n)
{
    for (int i = 1; i < n; i++) {
        RSeed v = sv[i];
        int j = i - 1;
        while (j >= 0 && seed_less(v, sv[j])) { sv[j + 1] = sv[j]; j--; }
        sv[j + 1] = v;
    }
}

HDN void remove_null(RSeed *sv, int32_t &n)
{
    int w = 0;
    for (int i = 0; i < n; i++) if (sv[i].rLen != 0) { if (w != i) sv[w] = sv[i]; w++; }
    n = w;
}

// scratch: 2*n ints
HDN void sort_by_rpos(const RSeed *sv, int n, int32_t *key, int32_t *idx)
{
    for (int i = 0; i < n; i++) {
        int k = sv[i].rPos, j = i - 1;
        while (j >= 0 && key[j] > k) { key[j + 1] = key[j]; idx[j + 1] = idx[j]; j--; }
        key[j + 1] = k; idx[j + 1] = i;
    }
}

HDN void remove_tandem(RSeed *sv, int32_t &n, int32_t *scratch)
{
    if (n < 2) return;
    int32_t *key = scratch, *idx = scratch + n;
    sort_by_rpos(sv, n, key, idx);
    bool any = false;
    for (int i = 0; i < n;) {
        int j = i + 1;
        while (j < n && key[j] == key[i]) j++;
        if (j - i > 1) { any = true; for (int k = i; k < j; k++) sv[idx[k]].rLen = sv[idx[k]].gLen = 0; }
        i = j;
    }
    if (any) remove_null(sv, n);
}

HDN void remove_translocated(RSeed *sv, int32_t &n, int32_t *scratch)
{
    if (n < 2) return;
    int32_t *key = scratch, *idx = scratch + n;
    sort_by_rpos(sv, n, key, idx);
    bool any = false;
    const int num = n;
    for (int i = 0; i < num; i++) {
        if (key[i] != sv[i].rPos) {
            any = true;
            int j = idx[i];
            for (int t = i + 1; t <= j; t++) if (idx[t] > j) j = idx[t];
            int s1 = 0, s2 = 0;
            for (int k = i; k <= j; k++) { if (k < idx[k]) s1 += sv[idx[k]].rLen; else s2 += sv[idx[k]].rLen; }
            if (s1 > s2) { for (int k = i; k <= j; k++) if (k > idx[k]) sv[idx[k]].rLen = sv[idx[k]].gLen = 0; }
            else { for (int k = i; k <= j; k++) if (k < idx[k]) sv[idx[k]].rLen = sv[idx[k]].gLen = 0; }
            i = j;
        }
    }
    if (any) remove_null(sv, n);
}

// ---------------------------------------------------------------------------------------------------
// candidate pairing / pruning, one read (single-end) or one pair
// ---------------------------------------------------------------------------------------------------
HDN void remove_redundant(CandState *v, int n)
{
    if (n <= 1) return;
    int score1 = 0, score2 = 0;
    for (int i = 0; i < n; i++) {
        int s = v[i].Score;
        if (s > score2) { if (s >= score1) { score2 = score1; score1 = s; } else score2 = s; }
        else if (s == score2) score2 = score1;
    }
    int thr = (score1 == score2 || score1 - score2 > 20) ? score1 : score2;
    for (int i = 0; i < n; i++) if (v[i].Score < thr) v[i].Score = 0;
}

HDN void pair_and_prune(CandState *v1, int n1, CandState *v2, int n2, bool paired)
{
    if (!paired) { remove_redundant(v1, n1); return; }
    bool pairing = false;
    if (n1 * n2 > 1000) { remove_redundant(v1, n1); remove_redundant(v2, n2); }
    for (int i = 0; i != n1; i++) {
        if (v1[i].Score == 0) continue;
        int best = -1;
        int64_t min_dist = 2000000;
        for (int j = 0; j != n2; j++) {
            if (v2[j].Score == 0 || v2[j].PosDiff < v1[i].PosDiff) continue;
            int64_t d = v2[j].PosDiff - v1[i].PosDiff;
            if (d < 0) d = -d;
            if (d < min_dist) { best = j; min_dist = d; }
        }
        if (best != -1) {
            int j = best;
            if (v2[j].PairedIdx == -1) { pairing = true; v1[i].PairedIdx = j; v2[j].PairedIdx = i; }
            else if (v1[i].Score > v1[v2[j].PairedIdx].Score) { v1[v2[j].PairedIdx].PairedIdx = -1; v1[i].PairedIdx = j; v2[j].PairedIdx = i; }
        }
    }
    if (pairing) {
        for (int i = 0; i != n1; i++) {
            if (v1[i].PairedIdx == -1) v1[i].Score = 0;
            else { int j = v1[i].PairedIdx; v1[i].Score = v2[j].Score = v1[i].Score + v2[j].Score; }
        }
        for (int j = 0; j != n2; j++) if (v2[j].PairedIdx == -1) v2[j].Score = 0;
    }
    remove_redundant(v1, n1);
    remove_redundant(v2, n2);
}

// seeds one live candidate can ever hold: n after the filters, + (n-1) re-seeds, + 2 per gap, + one pair per gap
HD int seed_capacity(int count) { return count <= 1 ? 1 : 12 * count + 4; }   // a lone seed has no gaps to fill

// Upper bound of the seeds (incl. scratch) a phase touches, from what is known when it starts: lets the kernels run the
// phase on a small shared-memory copy of the candidate's seeds instead of its HBM slice (ncu, round 1: phase C spent
// 95 % of its time on store-to-load round trips through L2).
HD int phase_seed_bound(const CandState &c, int which)
{
    if (which == 0) return c.seed_count + (8 * c.seed_count + 39) / 40;      // + the rPos sort scratch behind the seeds
    if (which == 1) return 2 * c.sv_n;                                       // one re-seed per gap
    if (which == 2) return 2 * (c.sv_n + 2 * c.n_ext);                       // two seeds per aligned gap, one pair per gap
    return c.sv_n;
}

// ---------------------------------------------------------------------------------------------------
// phase A: filters + windows for 8-mer re-seeding
// ---------------------------------------------------------------------------------------------------
HDN void phase_a(const Env &E, CandState &c, RSeed *sv)
{
    if (!c.live) return;
    int32_t n = c.seed_count;
    for (int s = 0; s < n; s++) {
        uint64_t key = E.keys[E.seed_off[c.read] + c.seed_begin + s];
        RSeed &d = sv[s];
        d.gPos = key_gpos(key); d.rPos = key_rpos(key); d.rLen = d.gLen = key_len(key);
        d.PosDiff = d.gPos - d.rPos; d.simple = 1; d.acceptor = 0; d.job = -1; d.pad0 = d.pad1 = 0; d.pad2 = 0;
    }
    int32_t *scratch = reinterpret_cast<int32_t *>(sv + n);   // the unused tail of the slice
    remove_tandem(sv, n, scratch);
    remove_translocated(sv, n, scratch);
    const int64_t coff = E.code_off[c.read];
    for (int i = 1; i < n; i++) {
        sv[i].job = -1;
        int rGaps;
        if ((int)(sv[i].PosDiff - sv[i - 1].PosDiff) > E.P.max_gaps && (rGaps = sv[i].rPos - sv[i - 1].rPos - sv[i - 1].rLen) > 20) {
            int rBegin = sv[i - 1].rPos + sv[i - 1].rLen;
            int64_t Lb = sv[i - 1].gPos + sv[i - 1].gLen, Rb = sv[i].gPos;
            int id = alloc_slots(E.kjob_count, 1);
            KmerJobDev j; j.s1_off = coff + rBegin; j.gpos = Lb; j.len1 = rGaps; j.len2 = (int32_t)(Rb - Lb > 0 ? Rb - Lb : 0);
            E.kjobs[id] = j;
            sv[i].job = id;
        }
    }
    c.sv_n = n;
}

// ---------------------------------------------------------------------------------------------------
// phase B: accept re-seeds; gaps to be aligned against both flanks
// ---------------------------------------------------------------------------------------------------
HDN void phase_b(const Env &E, CandState &c, RSeed *sv)
{
    if (!c.live) return;
    int32_t n = c.sv_n;
    const int64_t coff = E.code_off[c.read];
    const int num = n;
    bool added = false;
    for (int i = 1; i < num; i++) {
        int id = sv[i].job;
        if (id < 0) continue;
        const KmerJobDev &J = E.kjobs[id];
        const dartgpu_kmer_hit &h = E.khits[id];
        int thr = (int)(J.len1 * 0.85); if (thr < 8) thr = 8;
        if (h.len >= thr && h.len > 0) {
            RSeed d; d.simple = 1; d.acceptor = 0; d.job = -1; d.pad0 = d.pad1 = 0; d.pad2 = 0;
            d.rPos = h.rpos + (int)(J.s1_off - coff); d.gPos = (int64_t)h.gpos + J.gpos; d.rLen = d.gLen = h.len;
            d.PosDiff = d.gPos - d.rPos;
            sv[n++] = d; added = true;
        }
    }
    for (int i = 0; i < n; i++) sv[i].job = -1;
    if (added) sort_seeds(sv, n);
    int n_ext = 0;
    for (int i = 1; i < n; i++) {
        if ((int)(sv[i].PosDiff - sv[i - 1].PosDiff) > E.P.min_intron && sv[i].rPos > (sv[i - 1].rPos + sv[i - 1].rLen)) {
            int rGaps = sv[i].rPos - (sv[i - 1].rPos + sv[i - 1].rLen);
            int64_t s1 = coff + sv[i - 1].rPos + sv[i - 1].rLen;
            int id = alloc_slots(E.njob_count, 2);
            NwJobDev a; a.s1_off = s1; a.gpos = sv[i - 1].gPos + sv[i - 1].gLen; a.op_off = 0; a.flag_off = 0; a.aux_off = 0; a.m = rGaps; a.n = rGaps;
            NwJobDev b = a; b.gpos = sv[i].gPos - rGaps;
            E.njobs[id] = a; E.njobs[id + 1] = b;
            sv[i].job = id;
            n_ext++;
        }
    }
    c.sv_n = n;
    c.n_ext = n_ext;
}

// ---------------------------------------------------------------------------------------------------
// phase C helpers
// ---------------------------------------------------------------------------------------------------
struct Aln {              // one NW result: columns left to right
    const uint8_t *ops; int k;
};
HD Aln job_alignment(const Env &E, int id)
{
    const NwJobDev &J = E.done_jobs[id];
    int k = E.nops[id];
    Aln a; a.k = k; a.ops = E.ops + J.op_off + J.m + J.n - k;
    return a;
}

// IdentifyBestGappedPartition over the two alignments of a gap; scratch: 2*(rGaps+1) ints
HDN void gapped_partition(const Env &E, const uint8_t *gap, int rGaps, const RSeed &L, const RSeed &Rt, const Aln &A1, int64_t g1,
                          const Aln &A2, int64_t g2, int32_t *scratch, int &P_out, int &left_ext, int &right_ext)
{
    int32_t *Rv = scratch, *Lv = scratch + rGaps + 1;
    for (int t = 0; t <= rGaps; t++) Rv[t] = Lv[t] = 0;
    // alignment 1: genome gaps at the right end are re-filled with the bases that follow the window
    int last1 = A1.k - 1;
    while (last1 >= 0 && A1.ops[last1] == 2) last1--;          // a2[t] == '-'  <=> op 2
    {
        int i = 0, p = 0, s = 0; int64_t g = g1, fill = L.gPos + L.gLen + rGaps;
        RefCursor ref(E.ref), reff(E.ref);
        for (int t = 0; t < A1.k; t++) {
            int op = A1.ops[t];
            bool rgap = op == 1;                                 // '-' in the read string
            int gc = -1;
            if (op != 2) gc = ref.code(g++);
            else if (t > last1) gc = reff.code(fill++);
            if (!rgap && gc >= 0 && raw_eq(gap[i], gc)) s++;
            if (!rgap) { p++; i++; }
            Rv[p] = s;
        }
    }
    // alignment 2: genome gaps at the left end are re-filled walking LEFT from the window start (sic, :424-425)
    int first2 = 0;
    while (first2 < A2.k && A2.ops[first2] == 2) first2++;
    {
        // backward pass needs, for column t, the read index and genome index: precompute totals
        int ri = 0; int64_t gj = g2;
        for (int t = 0; t < A2.k; t++) { if (A2.ops[t] != 1) ri++; if (A2.ops[t] != 2) gj++; }
        int p = 0, s = 0;
        RefCursor ref(E.ref), reff(E.ref);
        for (int t = A2.k - 1; t >= 0; t--) {
            int op = A2.ops[t];
            bool rgap = op == 1;
            if (op != 2) gj--;
            if (!rgap) ri--;
            int gc = -1;
            if (op != 2) gc = ref.code(gj);
            else if (t < first2) gc = reff.code(g2 - (first2 - 1 - t));   // a4[i] for i = first2-1 .. 0 gets ref[g2], ref[g2-1], ..
            if (!rgap && gc >= 0 && raw_eq(gap[ri], gc)) s++;
            if (!rgap) p++;
            Lv[rGaps - p] = s;
        }
    }
    int best = 0, P = 0;
    for (int t = 0; t <= rGaps; t++) if (Rv[t] + Lv[t] > best) { best = Rv[t] + Lv[t]; P = t; }
    right_ext = left_ext = 0;
    if (!(best < (int)(rGaps * 0.8) || (rGaps - best) > E.P.max_mismatch)) {
        for (int p = P, t = 0; p > 0; t++) { int op = A1.ops[t]; if (op != 1) p--; if (op != 2 || t > last1) right_ext++; }
        for (int p = rGaps - P, t = A2.k - 1; p > 0; t--) { int op = A2.ops[t]; if (op != 1) p--; if (op != 2 || t < first2) left_ext++; }
    }
    P_out = P;
}

// motifs GT/AG, CT/AC, GC/AG, CT/GC as codes (main.cpp:18), donor pair << 4 | acceptor pair
HD int motif_pairs(int type)
{
    const int m[4] = {(2 << 2 | 3) << 4 | (0 << 2 | 2), (1 << 2 | 3) << 4 | (0 << 2 | 1), (2 << 2 | 1) << 4 | (0 << 2 | 2), (1 << 2 | 3) << 4 | (2 << 2 | 1)};
    return m[type];
}
HD int shift_of(int i) { return i == 0 ? 0 : ((i & 1) ? (i + 1) / 2 : -(i / 2)); }   // 0,1,-1,2,-2,..,9,-9

// IdentifySpliceJunction (AlignmentCandidates.cpp:732-756) + CheckSeqFragment (:702-730).  All bases it can look at lie in
// ref[L-9, L+11) and ref[R-11, R+9): two 32-base windows are loaded once and every comparison is a register operation.
struct JunctionWindows { uint64_t WL, WR; };    // WL starts at L-9, WR at R-11
HD JunctionWindows junction_windows(const Env &E, const RSeed &l, const RSeed &r)
{
    JunctionWindows w; w.WL = E.ref.window32(l.gPos + l.gLen - 9); w.WR = E.ref.window32(r.gPos - 11);
    return w;
}
HDN int find_junction(const JunctionWindows &W, int type, const RSeed &l, const RSeed &r)
{
    int i = l.rLen < r.rLen ? l.rLen : r.rLen, j = l.gLen < r.gLen ? l.gLen : r.gLen;
    if (i < j) j = i;
    if (j > 9) j = 9;
    j <<= 1;
    const int mp = motif_pairs(type);
    int shift = 0;
    for (i = 0; i <= j; i++) {
        shift = shift_of(i);
        // same_fragment: ref[L, L+shift) == ref[R, R+shift)  (shift > 0)  or  ref[L+shift, L) == ref[R+shift, R)  (shift < 0)
        if (shift > 0 && win_bases(W.WL, 9, shift) != win_bases(W.WR, 11, shift)) continue;
        if (shift < 0 && win_bases(W.WL, 9 + shift, -shift) != win_bases(W.WR, 11 + shift, -shift)) continue;
        // donor pair at L+shift, acceptor pair at R-2+shift
        if ((int)win_bases(W.WL, 9 + shift, 2) == (mp >> 4) && (int)win_bases(W.WR, 9 + shift, 2) == (mp & 15)) break;
    }
    return i > j ? 10 : shift;
}

// CheckSpliceJunction. The reference keeps the per-type (index, shift) lists; only the best type's list is applied,
// so the scan is done twice: once to choose the type, once to apply it.
HDN int check_splice_junction(const Env &E, RSeed *sv, int n)
{
    int min_cost = 1000, best_type = -1;
    for (int type = 0; type < 4; type++) {
        int mis = 0, cost = 0, found = 0;
        for (int i = 1; i < n; i++) {
            if ((sv[i].PosDiff - sv[i - 1].PosDiff) > E.P.min_intron && sv[i - 1].simple && sv[i].simple) {
                int shift = find_junction(junction_windows(E, sv[i - 1], sv[i]), type, sv[i - 1], sv[i]);
                if (shift != 10) found++; else mis++;
                cost += shift < 0 ? -shift : shift;
            }
        }
        if (found > 0 && cost < min_cost) { min_cost = cost; best_type = type; }
        if (mis == 0) break;
    }
    if (best_type != -1) {
        // shifts are evaluated on the untouched seeds (as the reference's stored list was), then applied in order
        // — applying one junction changes the lengths the next evaluation would see, so evaluate all first.
        int32_t cnt = 0;
        for (int i = 1; i < n; i++) sv[i].job = 10;
        for (int i = 1; i < n; i++)
            if ((sv[i].PosDiff - sv[i - 1].PosDiff) > E.P.min_intron && sv[i - 1].simple && sv[i].simple) { sv[i].job = find_junction(junction_windows(E, sv[i - 1], sv[i]), best_type, sv[i - 1], sv[i]); cnt++; }
        for (int j = 1; j < n; j++) {
            int shift = sv[j].job;
            if (shift != 10) {
                sv[j].acceptor = 1;
                if (shift != 0) {
                    sv[j - 1].rLen += shift; sv[j - 1].gLen += shift;
                    sv[j].rLen -= shift; sv[j].gLen -= shift;
                    sv[j].rPos += shift; sv[j].gPos += shift;
                }
            }
        }
        (void)cnt;
    }
    for (int i = 0; i < n; i++) sv[i].job = -1;
    return best_type;
}

HD bool seed_overlap(RSeed &p1, RSeed &p2)
{
    int ov;
    bool master = true;
    if ((ov = p1.rPos + p1.rLen - p2.rPos) > 0) {
        if (p1.rLen < p2.rLen) { master = false; if (p1.rLen > ov) p1.gLen = (p1.rLen -= ov); else p1.rLen = p1.gLen = 0; }
        else { if (p2.rLen > ov) { p2.rPos += ov; p2.gPos += ov; p2.gLen = (p2.rLen -= ov); } else p2.rLen = p2.gLen = 0; }
    }
    if ((p1.rLen > 0 && p2.rLen > 0) && (ov = (int)(p1.gPos + p1.gLen - p2.gPos)) > 0) {
        if (p1.gLen < p2.gLen) { master = false; if (p1.rLen > ov) p1.gLen = (p1.rLen -= ov); else p1.rLen = p1.gLen = 0; }
        else { if (p2.rLen > ov) { p2.rPos += ov; p2.gPos += ov; p2.gLen = (p2.rLen -= ov); } else p2.rLen = p2.gLen = 0; }
    }
    return master;
}

HDN void check_overlapping(RSeed *sv, int32_t &n)
{
    const int num = n;
    if (num < 2) return;
    bool null_seed = false;
    for (int i = 0; i < num;) {
        if (sv[i].rLen > 0) {
            int rEnd = sv[i].rPos + sv[i].rLen - 1;
            int64_t gEnd = sv[i].gPos + sv[i].gLen - 1;
            for (int j = i + 1; j < num; j++) {
                if (sv[j].rLen == 0) continue;
                if (rEnd < sv[j].rPos && gEnd < sv[j].gPos) break;
                if (!seed_overlap(sv[i], sv[j])) break;
            }
            if (sv[i].rLen == 0) {
                null_seed = true;
                int k = i - 1;
                while (k > 0 && sv[k].rLen == 0) k--;
                i = k < 0 ? 0 : k;
            } else i++;
        } else { null_seed = true; i++; }
    }
    if (null_seed) remove_null(sv, n);
}

HDN void identify_normal_pairs(RSeed *sv, int32_t &n)
{
    if (n <= 1) return;
    check_overlapping(sv, n);
    const int num = n;
    for (int i = 0, j = 1; j < num; i++, j++) {
        if (sv[j].rPos - sv[i].rPos - sv[i].rLen == 0) continue;
        int rGaps = sv[j].rPos - (sv[i].rPos + sv[i].rLen); if (rGaps < 0) rGaps = 0;
        int gGaps = (int)(sv[j].gPos - (sv[i].gPos + sv[i].gLen)); if (gGaps < 0) gGaps = 0; else if (gGaps > 30 && gGaps > (rGaps << 1)) gGaps = 0;
        if (rGaps > 0 || gGaps > 0) {
            RSeed d; d.simple = 0; d.acceptor = 0; d.job = -1; d.pad0 = d.pad1 = 0; d.pad2 = 0;
            d.rPos = sv[i].rPos + sv[i].rLen; d.gPos = sv[i].gPos + sv[i].gLen; d.PosDiff = d.gPos - d.rPos;
            d.rLen = rGaps; d.gLen = gGaps;
            sv[n++] = d;
        }
    }
    if (n > num) sort_seeds(sv, n);   // inplace_merge of two sorted runs == stable sort of the whole
}

HD bool coordinates_valid(const Env &E, const RSeed *sv, int n)
{
    int64_t g1 = 0, g2 = 2 * E.G;
    for (int i = 0; i < n; i++) if (sv[i].gLen > 0) { g1 = sv[i].gPos; break; }
    for (int i = n - 1; i >= 0; i--) if (sv[i].gLen > 0) { g2 = sv[i].gPos + sv[i].gLen - 1; break; }
    return !((g1 < E.G && g2 >= E.G) || (g1 >= E.G && g2 < E.G));
}

// the <=2-mismatch fast path of tools.cpp:149, :213, :261
HD bool simple_enough(const Env &E, const uint8_t *rc, const RSeed &sp, int *n_out)
{
    if (sp.rLen != sp.gLen) return false;
    int nm = 0;
    RefCursor ref(E.ref);
    for (int i = 0; i < sp.rLen; i++) if (!raw_eq(rc[sp.rPos + i], ref.code(sp.gPos + i))) nm++;
    *n_out = nm;
    return nm <= 2 && nm <= (int)(sp.rLen * 0.2);
}

// nw_alignment (nw_alignment.cpp:18-82) for blocks of at most NW_TINY x NW_TINY, in the calling thread: the integer half-unit
// recurrence and the traceback priority of nw_kernel.cu (S == R first, then S == T, else diagonal), flags in two 64-bit
// registers.  On BASELINE config[1] half of all candidates carry exactly one non-simple pair of a base or two (a mismatch
// between two exact seeds fails the "<= 2 mismatches and <= 20 %" shortcut of tools.cpp:149 when the block is shorter than 5
// bases): round 1 sent a million 1..10-cell alignments per batch through the job queue, the shape sort and the NW kernel, and
// every one of those candidates through one more pass of the phase kernels.  They are now aligned where they are needed.
// ops: columns left to right (0 both advance, 1 gap in the read string, 2 gap in the genome string); returns their number.
constexpr int NW_TINY = 8;
constexpr int JOB_NONE = -1, JOB_TINY = -2;
HDN int nw_tiny(const Env &E, const uint8_t *s1, int m, int64_t gpos, int n, uint8_t *ops)
{
    if (m <= 0 || n <= 0) {                       // one empty side: all gaps (nw_alignment.cpp:61-74 with i or j at 0)
        const int k = (m > 0 ? m : 0) + (n > 0 ? n : 0);
        for (int t = 0; t < k; t++) ops[t] = (uint8_t)(m <= 0 ? 1 : 2);
        return k;
    }
    constexpr int NEG = -30000;
    int g[NW_TINY], S[NW_TINY + 1], T[NW_TINY + 1];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < NW_TINY; j++) { g[j] = j < n ? E.ref.code(gpos + j) : 0; S[j + 1] = -3 - j; T[j + 1] = NEG; }
    S[0] = 0; T[0] = NEG;
    uint64_t fR = 0, fT = 0;                      // cell (i, j) at bit (i - 1) * 8 + (j - 1)
    for (int i = 1; i <= m; i++) {
        int a = (int)s1[i - 1];
        a = (a & 4) ? 7 : (a & 3);                // 8..11 are lower-case ACGT: same base for the table compare
        int Sl = -2 - i, Rl = NEG, Sd = S[0];
        S[0] = Sl;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 1; j <= NW_TINY; j++) {
            if (j <= n) {
                const int Su = S[j], Tu = T[j];
                const int R = (Rl - 1 > Sl - 3) ? Rl - 1 : Sl - 3;
                const int Tt = (Tu - 1 > Su - 3) ? Tu - 1 : Su - 3;
                int h = Sd + (a == g[j - 1] ? 3 : -3);
                if (R > h) h = R;
                if (Tt > h) h = Tt;
                const int Sc = (h / 2) * 2;       // the reference's max(short, short, short) truncates toward zero
                const int bit = (i - 1) * 8 + (j - 1);
                if (Sc == R) fR |= 1ull << bit;
                if (Sc == Tt) fT |= 1ull << bit;
                Sd = Su; Sl = Sc; Rl = R; S[j] = Sc; T[j] = Tt;
            }
        }
    }
    uint8_t rev[2 * NW_TINY];
    int ti = m, tj = n, k = 0;
    while (ti > 0 || tj > 0) {
        int op;
        if (ti == 0) op = 1;
        else if (tj == 0) op = 2;
        else { const int bit = (ti - 1) * 8 + (tj - 1); op = ((fR >> bit) & 1) ? 1 : (((fT >> bit) & 1) ? 2 : 0); }
        rev[k++] = (uint8_t)op;
        if (op == 1) tj--; else if (op == 2) ti--; else { ti--; tj--; }
    }
    for (int t = 0; t < k; t++) ops[t] = rev[k - 1 - t];
    return k;
}

// ---------------------------------------------------------------------------------------------------
// phase C: gapped partitions -> seeds, splice-motif snapping, normal pairs, pair alignments to run
// ---------------------------------------------------------------------------------------------------
HDN void phase_c(const Env &E, CandState &c, RSeed *sv)
{
    if (!c.live) return;
    int32_t n = c.sv_n;
    const uint8_t *rc = E.codes + E.code_off[c.read];
    const int num = n;
    for (int i = 1; i < num; i++) {
        int id = sv[i].job;
        if (id < 0) continue;
        const RSeed &L = sv[i - 1], &Rt = sv[i];
        const NwJobDev &J1 = E.done_jobs[id], &J2 = E.done_jobs[id + 1];
        const int rGaps = J1.m;
        int P, le, re;
        gapped_partition(E, rc + L.rPos + L.rLen, rGaps, L, Rt, job_alignment(E, id), J1.gpos, job_alignment(E, id + 1), J2.gpos,
                         E.xscratch + J1.aux_off, P, le, re);
        RSeed d; d.simple = 0; d.acceptor = 0; d.job = -1; d.pad0 = d.pad1 = 0; d.pad2 = 0;
        int rest = rGaps;
        if (P > 0) {
            d.rPos = L.rPos + L.rLen; d.gPos = L.gPos + L.gLen; d.PosDiff = d.gPos - d.rPos; d.rLen = P; d.gLen = re;
            sv[n++] = d;
        }
        if ((rest -= P) > 0) {
            d.rLen = rest; d.gLen = le; d.rPos = Rt.rPos - d.rLen; d.gPos = Rt.gPos - d.gLen; d.PosDiff = d.gPos - d.rPos;
            sv[n++] = d;
        }
    }
    for (int i = 0; i < n; i++) sv[i].job = -1;
    if (n > num) sort_seeds(sv, n);
    c.SJtype = check_splice_junction(E, sv, n);
    identify_normal_pairs(sv, n);
    c.sv_n = n;
    c.cig_cap = 0;
    if (n > 1 && !coordinates_valid(E, sv, n)) { c.skip = 1; return; }
    int cap = 2 * n + 4;
    const int64_t coff = E.code_off[c.read];
    for (int j = 0; j < n; j++) {
        RSeed &sp = sv[j];
        sp.job = -1;
        if ((sp.rLen == 0 && sp.gLen == 0) || sp.simple) continue;
        bool middle = !(j == 0 || j == n - 1);
        if (middle && (sp.PosDiff == -1 || sp.rLen == 0 || sp.gLen == 0)) continue;
        int nm;
        if (simple_enough(E, rc, sp, &nm)) continue;
        cap += sp.rLen + sp.gLen + 2;
        if (sp.rLen <= NW_TINY && sp.gLen <= NW_TINY) { sp.job = JOB_TINY; continue; }       // aligned in place by phase D (nw_tiny)
        int id = alloc_slots(E.njob_count, 1);
        NwJobDev a; a.s1_off = coff + sp.rPos; a.gpos = sp.gPos; a.op_off = 0; a.flag_off = 0; a.aux_off = 0; a.m = sp.rLen; a.n = sp.gLen;
        E.njobs[id] = a;
        sp.job = id;
    }
    c.cig_cap = cap;
}

// ---------------------------------------------------------------------------------------------------
// phase D: CIGAR, score, coordinates
// ---------------------------------------------------------------------------------------------------
struct CigW {             // appends (len, op) pairs
    int32_t *p; int n, cap;
    HD void push(int len, char op) { if (n < cap) p[n] = len << 8 | (uint8_t)op; n++; }
};

// AddNewCigarElements over columns [t0,t1) of an alignment; i/g are the read / genome indices at t0
HDN int add_cigar(const Env &E, const uint8_t *s1, int64_t gpos, const Aln &A, int t0, int t1, int i, int64_t g, CigW &cv)
{
    char state = '*';
    int c = 0, score = 0;
    RefCursor ref(E.ref);
    for (int t = t0; t < t1; t++) {
        int op = A.ops[t];
        char want;
        if (op == 1) { want = 'D'; g++; }
        else if (op == 2) { want = 'I'; i++; }
        else { want = 'M'; if (raw_eq(s1[i], ref.code(g))) score++; i++; g++; }
        if (state == want) c++;
        else { if (c > 0) cv.push(c, state); c = 1; state = want; }
    }
    if (c > 0) cv.push(c, state);
    (void)gpos;
    return score;
}

HDN bool local_quality_ok(const Env &E, const uint8_t *s1, int64_t gpos, const Aln &A)
{
    int type = -1, nn = 0, mis = 0, status = 0, i = 0;
    int64_t g = gpos;
    RefCursor ref(E.ref);
    for (int t = 0; t < A.k; t++) {
        int op = A.ops[t], ty;
        if (op == 1) { ty = 0; g++; }
        else if (op == 2) { ty = 1; i++; }
        else { ty = 2; nn++; if (!raw_eq(s1[i], ref.code(g))) mis++; i++; g++; }
        if (type != ty) { type = ty; status++; }
    }
    return !(status >= 4 || (mis >= 3 && mis >= (int)(nn * 0.3)));
}

HD int digits10(int v) { int d = 1; while (v >= 10) { v /= 10; d++; } return d; }

HDN void phase_d(const Env &E, CandState &c, RSeed *sv)
{
    c.AlnScore = 0; c.cig_n = 0; c.text_len = 0;
    if (!c.live || c.skip) return;
    const int n = c.sv_n;
    const uint8_t *rc = E.codes + E.code_off[c.read];
    CigW cv; cv.p = E.cig + c.cig_off; cv.n = 0; cv.cap = c.cig_cap;
    int first_rpos = n > 0 ? sv[0].rPos : 0;   // the leading soft clip is inserted at the front: reserve slot 0
    cv.push(0, 'S');
    int mis = 0, aln = 0;
    for (int j = 0; j != n; j++) {
        RSeed &sp = sv[j];
        if (sp.rLen == 0 && sp.gLen == 0) continue;
        int g;
        if (j > 0 && (g = (int)(sp.gPos - (sv[j - 1].gPos + sv[j - 1].gLen))) > 0) cv.push(g, 'N');
        if (sp.simple) { cv.push(sp.rLen, 'M'); aln += sp.rLen; continue; }
        const bool head = j == 0, tail = !head && j == n - 1;
        int score = 0;
        if (!head && !tail && sp.PosDiff == -1) cv.push(sp.rLen, 'S');
        else if (!head && !tail && (sp.rLen == 0 || sp.gLen == 0)) {
            if (sp.rLen > 0) cv.push(sp.rLen, 'I'); else if (sp.gLen > 0) cv.push(sp.gLen, 'D');
        } else if (sp.job == JOB_NONE) {
            int nm = 0;
            simple_enough(E, rc, sp, &nm);
            score = sp.rLen - nm;
            cv.push(sp.rLen, 'M');
        } else {
            uint8_t tiny_ops[2 * NW_TINY];
            Aln A;
            if (sp.job == JOB_TINY) { A.k = nw_tiny(E, rc + sp.rPos, sp.rLen, sp.gPos, sp.gLen, tiny_ops); A.ops = tiny_ops; }
            else A = job_alignment(E, sp.job);
            const uint8_t *s1 = rc + sp.rPos;
            const int64_t g0 = sp.gPos;
            if (!head && !tail) score = add_cigar(E, s1, g0, A, 0, A.k, 0, g0, cv);
            else if (!local_quality_ok(E, s1, g0, A)) { cv.push(sp.rLen, 'S'); score = 0; }
            else if (head) {
                int t = 0, p = 0;
                while (t < A.k && A.ops[t] == 1) { t++; p++; }          // leading '-' in the read block: shrink the genome block
                int i0 = 0; int64_t gg = g0 + p;
                if (p > 0) { sp.gPos += p; sp.gLen -= p; }
                int q = 0;
                while (t < A.k && A.ops[t] == 2) { t++; q++; }          // then leading '-' in the genome block: soft clip
                if (q > 0) { sp.rPos += q; sp.rLen -= q; cv.push(q, 'S'); i0 = q; }
                score = add_cigar(E, s1, g0, A, t, A.k, i0, gg, cv);
            } else {
                int t1 = A.k, cnt = 0;
                while (t1 > 0 && A.ops[t1 - 1] == 1) { t1--; cnt++; }   // trailing '-' in the read block
                if (cnt > 0) sp.gLen -= cnt;
                cnt = 0;
                while (t1 > 0 && A.ops[t1 - 1] == 2) { t1--; cnt++; }   // then trailing '-' in the genome block
                if (cnt > 0) sp.rLen -= cnt;
                score = add_cigar(E, s1, g0, A, 0, t1, 0, g0, cv);
                if (cnt > 0) cv.push(cnt, 'S');
            }
        }
        aln += score;
        mis += sp.rLen - score;
    }
    int lead = 0;
    if (n > 0) {
        (void)first_rpos;
        if ((lead = sv[0].rPos) > 0) cv.p[0] = lead << 8 | 'S';
        int tailclip = E.rlen[c.read] - (sv[n - 1].rPos + sv[n - 1].rLen);
        if (tailclip > 0) cv.push(tailclip, 'S');
    }
    // drop the reserved slot when there is no leading clip
    int start = lead > 0 ? 0 : 1;
    int cn = cv.n - start;
    if (cv.n > cv.cap) { c.AlnScore = 0; c.cig_n = -1; return; }     // capacity bug guard (never expected)
    if (mis > E.P.max_mismatch || cn == 0) aln = 0;
    for (int t = start; t < cv.n; t++) if ((cv.p[t] & 0xff) == 'N' && (cv.p[t] >> 8) < E.P.min_intron) { aln = 0; break; }
    c.mis = mis;
    if (aln > 0) {
        const bool first = !E.P.pair_end || (c.read & 1) == 0;
        int64_t gPos = sv[0].gPos, end_gPos = sv[n - 1].gPos + sv[n - 1].gLen - 1, endc;
        c.chr = chr_lookup(E, gPos, &endc);
        if (gPos < E.G) { c.dir = first ? 1 : 0; c.pos = gPos + 1 - E.chr_fwd[c.chr]; }
        else { c.dir = first ? 0 : 1; c.pos = endc - end_gPos + 1; }
        if (c.pos <= 0) aln = 0;
        else {
            // GenerateCIGAR: reverse for the reverse strand, merge equal neighbours; compacted to the front of the slice
            int32_t *p = cv.p;
            if (gPos >= E.G) for (int a = start, b = cv.n - 1; a < b; a++, b--) { int32_t t = p[a]; p[a] = p[b]; p[b] = t; }
            int w = 0, tl = 0;
            char state = '\0'; int acc = 0;
            for (int t = start; t < cv.n; t++) {
                char op = (char)(p[t] & 0xff); int len = p[t] >> 8;
                if (op != state) { if (acc > 0) { p[w++] = acc << 8 | (uint8_t)state; tl += digits10(acc) + 1; } acc = len; state = op; }
                else acc += len;
            }
            if (acc > 0) { p[w++] = acc << 8 | (uint8_t)state; tl += digits10(acc) + 1; }
            c.cig_n = w; c.text_len = tl;
        }
    }
    c.AlnScore = aln;
}

// ---------------------------------------------------------------------------------------------------
// final pass, one read (single-end) or one pair: best / second best, mate rescue, flags, MAPQ
// ---------------------------------------------------------------------------------------------------
// rep: the read's report records (n_reports entries, already holding aln_score / sj_type / paired_idx / dir ...)
HDN void read_best(const CandState *cs, int ncand, dartgpu_report *rep, ReadOut &r)
{
    r.score = r.best = 0; r.sub_score = 0; r.mis_num = 0; r.mapq = 0;
    r.n_reports = ncand > 0 ? ncand : 1;
    if (ncand == 0) {
        rep[0].aln_score = 0; rep[0].sj_type = -1; rep[0].flag = 0; rep[0].paired_idx = -1; rep[0].dir = 0; rep[0].chr_idx = 0;
        rep[0].pos = 0; rep[0].cigar_off = 0; rep[0].cigar_len = 0; rep[0].reserved[0] = rep[0].reserved[1] = 0;
        return;
    }
    for (int k = 0; k < ncand; k++) {
        const CandState &a = cs[k];
        dartgpu_report &p = rep[k];
        p.aln_score = 0; p.sj_type = -1; p.flag = 0; p.paired_idx = a.PairedIdx; p.dir = 0; p.chr_idx = 0; p.pos = 0;
        p.cigar_off = 0; p.cigar_len = 0; p.reserved[0] = p.reserved[1] = 0;
        if (!a.live) continue;
        p.sj_type = a.SJtype;
        if (a.skip) continue;
        p.aln_score = a.AlnScore;
        if (a.AlnScore > 0) {
            p.dir = a.dir; p.chr_idx = a.chr; p.pos = a.pos; p.cigar_len = (int16_t)a.text_len;
            if (a.AlnScore > r.score) { r.best = k; r.mis_num = a.mis; r.sub_score = r.score; r.score = a.AlnScore; }
            else if (a.AlnScore == r.score) r.sub_score = r.score;
        }
    }
}

HD void eval_mapq(ReadOut &r, const dartgpu_report *rep)
{
    if (r.score == 0 || r.score == r.sub_score) r.mapq = 0;
    else if (r.sub_score == 0 || r.score > r.sub_score) r.mapq = 50;
    else {
        int m = 0;
        for (int k = 0; k < r.n_reports; k++) if (rep[k].aln_score == r.score) m++;
        r.mapq = m >= 10 ? 0 : m >= 4 ? 1 : m == 3 ? 2 : m == 2 ? 3 : 50;
    }
}

// `dir` of a report whose score was later zeroed is still read by the flag logic: it is kept in cs[k].dir
HDN void finish_single(ReadOut &r, dartgpu_report *rep, const CandState *cs, int ncand)
{
    auto dir_of = [&](int k) { return k < ncand ? (int)cs[k].dir : 0; };
    if (r.score > r.sub_score) { int k = r.best; rep[k].flag = dir_of(k) ? 0 : 0x10; }
    else if (r.score > 0) { for (int k = 0; k < r.n_reports; k++) if (rep[k].aln_score > 0) rep[k].flag = dir_of(k) ? 0 : 0x10; }
    else rep[0].flag = 0x4;
    eval_mapq(r, rep);
}

HDN void finish_pair(ReadOut &r1, dartgpu_report *p1, const CandState *c1, int n1, ReadOut &r2, dartgpu_report *p2,
                     const CandState *c2, int n2, bool multi_hit)
{
    auto d1 = [&](int k) { return k < n1 ? (int)c1[k].dir : 0; };
    auto d2 = [&](int k) { return k < n2 ? (int)c2[k].dir : 0; };
    {   // CheckPairedFinalAlignments (Mapping.cpp:479-530)
        bool mated = p1[r1.best].paired_idx == r2.best;
        if (!(!multi_hit && mated)) {
            if (!mated && r1.score > 0 && r2.score > 0) {
                int s = 0;
                for (int i = 0; i != r1.n_reports; i++) {
                    int j;
                    if (p1[i].aln_score > 0 && (j = p1[i].paired_idx) != -1 && p2[j].aln_score > 0) {
                        mated = true;
                        if (s < p1[i].aln_score + p2[j].aln_score) {
                            s = p1[i].aln_score + p2[j].aln_score;
                            r1.best = i; r1.score = p1[i].aln_score;
                            r2.best = j; r2.score = p2[j].aln_score;
                        }
                    }
                }
            }
            if (mated) {
                for (int i = 0; i != r1.n_reports; i++) {
                    int j;
                    if (p1[i].aln_score != r1.score || ((j = p1[i].paired_idx) != -1 && p2[j].aln_score != r2.score)) { p1[i].aln_score = 0; p1[i].paired_idx = -1; }
                }
            } else {
                for (int i = 0; i != r1.n_reports; i++) { if (p1[i].paired_idx != -1) p1[i].paired_idx = -1; if (p1[i].aln_score > 0 && p1[i].aln_score != r1.score) p1[i].aln_score = 0; }
                for (int j = 0; j != r2.n_reports; j++) { if (p2[j].paired_idx != -1) p2[j].paired_idx = -1; if (p2[j].aln_score > 0 && p2[j].aln_score != r2.score) p2[j].aln_score = 0; }
            }
        }
    }
    int i, j;   // SetPairedAlignmentFlag (Mapping.cpp:101-186)
    if (r1.score > r1.sub_score && r2.score > r2.sub_score) {
        i = r1.best; j = r2.best;
        p1[i].flag = 0x41; p2[j].flag = 0x81;
        if (j == p1[i].paired_idx) { p1[i].flag |= 0x2; p2[j].flag |= 0x2; }
        p1[i].flag |= d1(i) ? 0x20 : 0x10;
        p2[j].flag |= d2(j) ? 0x20 : 0x10;
    } else {
        if (r1.score > r1.sub_score) {
            i = r1.best; p1[i].flag = 0x41 | (d1(i) ? 0x20 : 0x10);
            if ((j = p1[i].paired_idx) != -1 && p2[j].aln_score > 0) p1[i].flag |= 0x2; else p1[i].flag |= 0x8;
        } else if (r1.score > 0) {
            for (i = 0; i < r1.n_reports; i++) if (p1[i].aln_score > 0) {
                p1[i].flag = 0x41 | (d1(i) ? 0x20 : 0x10);
                if ((j = p1[i].paired_idx) != -1 && p2[j].aln_score > 0) p1[i].flag |= 0x2; else p1[i].flag |= 0x8;
            }
        } else { p1[0].flag = 0x41 | 0x4; if (r2.score == 0) p1[0].flag |= 0x8; else p1[0].flag |= d2(r2.best) ? 0x10 : 0x20; }
        if (r2.score > r2.sub_score) {
            j = r2.best; p2[j].flag = 0x81 | (d2(j) ? 0x20 : 0x10);
            if ((i = p2[j].paired_idx) != -1 && p1[i].aln_score > 0) p2[j].flag |= 0x2; else p2[j].flag |= 0x8;
        } else if (r2.score > 0) {
            for (j = 0; j < r2.n_reports; j++) if (p2[j].aln_score > 0) {
                p2[j].flag = 0x81 | (d2(j) ? 0x20 : 0x10);
                if ((i = p2[j].paired_idx) != -1 && p1[i].aln_score > 0) p2[j].flag |= 0x2; else p2[j].flag |= 0x8;
            }
        } else { p2[0].flag = 0x81 | 0x4; if (r1.score == 0) p2[0].flag |= 0x8; else p2[0].flag |= d1(r1.best) ? 0x10 : 0x20; }
    }
    eval_mapq(r1, p1); eval_mapq(r2, p2);
}

// junction records of the read's best alignment (UpdateLocalSJMap). out == nullptr: count only.
HDN int emit_junctions(const Env &E, const ReadOut &r, const CandState *cs, int ncand, int read, dartgpu_junction *out)
{
    if (!((r.mapq == 50 || (E.P.all_sj && r.score > 0)) && r.best < ncand)) return 0;
    const CandState &a = cs[r.best];
    if (!a.live || a.SJtype == -1) return 0;
    const RSeed *sv = E.pool + a.sv_off;
    int k = 0;
    for (int s = 1; s < a.sv_n; s++) {
        if (!sv[s].acceptor) continue;
        int64_t g1, g2;
        if (a.PosDiff < E.G) { g1 = sv[s - 1].gPos + sv[s - 1].gLen; g2 = sv[s].gPos - 1; }
        else { g1 = 2 * E.G - sv[s].gPos; g2 = 2 * E.G - 1 - (sv[s - 1].gPos + sv[s - 1].gLen); }
        int64_t d = g2 - g1; if (d < 0) d = -d;
        if (d < E.P.min_intron) continue;
        if (out) { out[k].g1 = g1; out[k].g2 = g2; out[k].type = a.SJtype; out[k].read = read; }
        k++;
    }
    return k;
}

// CIGAR text of one candidate from its merged pairs
HDN void write_cigar_text(const int32_t *pairs, int n, char *out)
{
    for (int t = 0; t < n; t++) {
        int len = pairs[t] >> 8, d = digits10(len);
        for (int k = d - 1; k >= 0; k--) { out[k] = (char)('0' + len % 10); len /= 10; }
        out[d] = (char)(pairs[t] & 0xff);
        out += d + 1;
    }
}

} // namespace dartgpu
