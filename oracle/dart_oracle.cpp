// TEST INFRASTRUCTURE — the CPU oracle (see dart_oracle.h). Never part of the product path.
//
// Every function names the reference lines it restates. The code is written from the algorithm,
// in the reformulated shape the CUDA kernels use (popcount ranks instead of the byte table,
// adjacent-pair cluster boundaries instead of the greedy scan, per-diagonal aggregation instead
// of two sorts, integer half-unit NW instead of float + truncating max), so that agreement with
// oracle/_ref/libdartref.so proves the reformulations, not just a transcription.
#include "dart_oracle.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct or_index {
    uint64_t primary = 0, L2[5] = {0, 0, 0, 0, 0}, seq_len = 0;
    std::vector<uint32_t> bwt;   // 16 words per 128-base block: 4 x u64 counts, 8 x u32 of 2-bit symbols
    std::vector<uint64_t> sa;    // sampled every sa_intv; sa[0] = -1
    uint64_t sa_intv = 32;
    int64_t G = 0;               // forward genome length; text = forward + reverse complement
    std::vector<uint8_t> pac;    // 2-bit, MSB first, forward strand only
    std::vector<std::string> chr_name;
    std::vector<int64_t> chr_len, chr_fwd;
    std::vector<int64_t> ends;   // sorted last-base coordinates of every chromosome on both strands
    or_counters ctr{};
};

static bool read_file(const std::string &fn, std::vector<uint8_t> &buf)
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

// File formats: /root/reference/src/bwt_index.cpp:15-35 (.sa), :37-70 (.ann), :102-121 (.bwt), :229-253 (.pac)
extern "C" or_index *or_load(const char *prefix)
{
    or_index *ix = new or_index;
    std::vector<uint8_t> b;
    std::string p(prefix);
    if (!read_file(p + ".bwt", b) || b.size() < 40) { delete ix; return nullptr; }
    memcpy(&ix->primary, b.data(), 8);
    memcpy(&ix->L2[1], b.data() + 8, 32);
    ix->seq_len = ix->L2[4];
    ix->bwt.resize((b.size() - 40) / 4);
    memcpy(ix->bwt.data(), b.data() + 40, ix->bwt.size() * 4);

    if (!read_file(p + ".sa", b) || b.size() < 56) { delete ix; return nullptr; }
    memcpy(&ix->sa_intv, b.data() + 40, 8);
    uint64_t n_sa = (ix->seq_len + ix->sa_intv) / ix->sa_intv;
    ix->sa.assign(n_sa, 0);
    ix->sa[0] = (uint64_t)-1;
    memcpy(ix->sa.data() + 1, b.data() + 56, std::min<size_t>((n_sa - 1) * 8, b.size() - 56));

    FILE *fp = fopen((p + ".ann").c_str(), "r");
    if (!fp) { delete ix; return nullptr; }
    long long lpac; int nseq; unsigned seed;
    if (fscanf(fp, "%lld%d%u", &lpac, &nseq, &seed) != 3) { fclose(fp); delete ix; return nullptr; }
    ix->G = lpac;
    int64_t acc = 0;
    for (int i = 0; i < nseq; i++) {
        unsigned gi; char name[1024]; long long off; int len, namb;
        if (fscanf(fp, "%u%1023s", &gi, name) != 2) break;
        int c; while ((c = fgetc(fp)) != '\n' && c != EOF) {}
        if (fscanf(fp, "%lld%d%d", &off, &len, &namb) != 3) break;
        ix->chr_name.push_back(name); ix->chr_len.push_back(len); ix->chr_fwd.push_back(acc);
        acc += len;
        ix->ends.push_back(ix->chr_fwd.back() + len - 1);
        ix->ends.push_back(2 * ix->G - acc + len - 1);
    }
    fclose(fp);
    std::sort(ix->ends.begin(), ix->ends.end());
    if (!read_file(p + ".pac", ix->pac)) { delete ix; return nullptr; }
    return ix;
}
extern "C" void or_free(or_index *ix) { delete ix; }
extern "C" int64_t or_genome_size(const or_index *ix) { return ix->G; }
extern "C" int or_num_chromosomes(const or_index *ix) { return (int)ix->chr_len.size(); }
extern "C" void or_counters_get(const or_index *ix, or_counters *o) { *o = ix->ctr; }
extern "C" void or_counters_reset(or_index *ix) { ix->ctr = or_counters{}; }

// RefSequence restated as codes: forward from .pac, reverse strand = complement read backwards
// (/root/reference/src/bwt_index.cpp:193-212).
static inline uint8_t ref_code(const or_index *ix, int64_t p)
{
    if (p < 0 || p >= 2 * ix->G) return 0; // the reference would read out of bounds here
    if (p < ix->G) return ix->pac[p >> 2] >> ((~p & 3) << 1) & 3;
    int64_t q = 2 * ix->G - 1 - p;
    return 3 - (ix->pac[q >> 2] >> ((~q & 3) << 1) & 3);
}
extern "C" void or_ref_codes(const or_index *ix, int64_t pos, int len, uint8_t *out)
{
    for (int i = 0; i < len; i++) out[i] = ref_code(ix, pos + i);
}

// ---- FM-index ranks -------------------------------------------------------------------------
// Counts of A,C,G,T among the first `n` (0..16) symbols of a 2-bit word, symbol 0 in the top bits.
static inline void count16(uint32_t w, int n, uint64_t cnt[4])
{
    if (n <= 0) return;
    uint32_t keep = n >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * n));
    uint32_t lo = w & 0x55555555u & keep, hi = (w >> 1) & 0x55555555u & keep;
    uint32_t valid = 0x55555555u & keep;
    int t = __builtin_popcount(hi & lo), g = __builtin_popcount(hi & ~lo & valid), c = __builtin_popcount(~hi & lo & valid);
    cnt[3] += t; cnt[2] += g; cnt[1] += c; cnt[0] += n - t - g - c;
}

// bwt_occ4 (/root/reference/src/bwt_search.cpp:67-84): occurrences of each symbol in BWT[0..k].
extern "C" void or_rank4(const or_index *ix, uint64_t k, uint64_t cnt[4])
{
    if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
    k -= (k >= ix->primary); // the sentinel is not stored
    const uint32_t *blk = ix->bwt.data() + ((k >> 7) << 4);
    memcpy(cnt, blk, 32);
    int upto = (int)(k & 127) + 1; // symbols of this block to include
    for (int w = 0; w < 8; w++) count16(blk[8 + w], std::min(16, upto - 16 * w), cnt);
}

// bwt_invPsi (/root/reference/src/bwt_search.cpp:119-125): one LF step; symbol and rank come from one block.
static uint64_t lf_step(or_index *ix, uint64_t k)
{
    ix->ctr.lf_steps++;
    if (k == ix->primary) return 0;
    uint64_t x = k - (k > ix->primary);
    uint32_t w = ix->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)];
    int c = (w >> ((~x & 15) << 1)) & 3;
    uint64_t cnt[4];
    or_rank4(ix, k, cnt);
    return ix->L2[c] + cnt[c];
}

// bwt_sa (/root/reference/src/bwt_search.cpp:127-137)
extern "C" uint64_t or_locate(or_index *ix, uint64_t k)
{
    uint64_t steps = 0;
    while (k & (ix->sa_intv - 1)) { steps++; k = lf_step(ix, k); }
    ix->ctr.hits++;
    return steps + ix->sa[k / ix->sa_intv];
}

// BWT_Search (/root/reference/src/bwt_search.cpp:139-182).  mirrored = false: locate from the forward interval as the
// reference does.  mirrored = true: locate from the reverse-complement interval and mirror (the CUDA path's route).
static int search_impl(or_index *ix, const uint8_t *s, int start, int stop, int max_dup, int *len_out, uint64_t *locs, int cap,
                       bool mirrored, bool count)
{
    if (count) ix->ctr.searches++;
    int c0 = s[start];
    uint64_t x0 = ix->L2[c0] + 1, x1 = ix->L2[3 - c0] + 1, x2 = ix->L2[c0 + 1] - ix->L2[c0];
    int pos;
    for (pos = start + 1; pos < stop; pos++) {
        if (s[pos] > 3) break;
        uint64_t k = x1 - 1, l = x1 - 1 + x2, tk[4], tl[4];
        if (count) {
            ix->ctr.ext_steps++;
            ix->ctr.ext_blocks += ((k - (k >= ix->primary)) >> 7) == ((l - (l >= ix->primary)) >> 7) ? 1 : 2;
        }
        or_rank4(ix, k, tk);
        or_rank4(ix, l, tl);
        int c = 3 - s[pos];
        uint64_t n2 = tl[c] - tk[c];
        if (n2 == 0) break;
        // forward-strand interval start: skip the sub-intervals of the symbols above c, and the sentinel
        uint64_t n0 = x0 + ((x1 <= ix->primary && x1 + x2 - 1 >= ix->primary) ? 1 : 0);
        for (int j = 3; j > c; j--) n0 += tl[j] - tk[j];
        x0 = n0; x1 = ix->L2[c] + 1 + tk[c]; x2 = n2;
    }
    *len_out = 0;
    if (x2 <= (uint64_t)(unsigned)max_dup && (*len_out = pos - start) >= 16) {
        for (uint64_t j = 0; j < x2; j++) {
            uint64_t g;
            if (!mirrored) {
                g = or_locate(ix, x0 + j);
                if (count) {   // the walk the CUDA path does for the same hit, counted without touching the other counters
                    or_counters keep = ix->ctr;
                    or_locate(ix, x1 + j);
                    uint64_t walked = ix->ctr.lf_steps - keep.lf_steps;
                    ix->ctr = keep;
                    ix->ctr.lf_steps_rc += walked;
                }
            } else {
                or_counters keep = ix->ctr;
                g = 2 * (uint64_t)ix->G - or_locate(ix, x1 + j) - (uint64_t)*len_out;
                ix->ctr = keep;
            }
            if ((int)j < cap) locs[j] = g;
        }
        return (int)x2;
    }
    return 0;
}

extern "C" int or_search(or_index *ix, const uint8_t *s, int start, int stop, int max_dup, int *len_out, uint64_t *locs, int cap)
{
    return search_impl(ix, s, start, stop, max_dup, len_out, locs, cap, false, true);
}

extern "C" int or_search_mirrored(or_index *ix, const uint8_t *s, int start, int stop, int max_dup, int *len_out, uint64_t *locs, int cap)
{
    return search_impl(ix, s, start, stop, max_dup, len_out, locs, cap, true, false);
}

// IdentifySeedPairs (/root/reference/src/AlignmentCandidates.cpp:181-215)
extern "C" int or_seed_read(or_index *ix, const uint8_t *s, int rlen, int max_dup,
                            int32_t *rpos, int64_t *gpos, int32_t *len, int cap)
{
    struct Seed { int64_t g; int32_t r, l; };
    std::vector<Seed> v;
    std::vector<uint64_t> locs(std::max(max_dup, 1));
    ix->ctr.read_bases += rlen;
    for (int pos = 0; pos < rlen - 13;) {
        if (s[pos] > 3) { pos++; continue; }
        int l, f = or_search(ix, s, pos, rlen, max_dup, &l, locs.data(), (int)locs.size());
        if (f > 0) {
            for (int j = 0; j < f; j++) v.push_back({(int64_t)locs[j], pos, l});
            pos += l;
        } else pos++;
    }
    std::sort(v.begin(), v.end(), [](const Seed &a, const Seed &b) { return a.g != b.g ? a.g < b.g : a.r < b.r; });
    ix->ctr.seeds += v.size();
    for (size_t i = 0; i < v.size() && (int)i < cap; i++) { rpos[i] = v[i].r; gpos[i] = v[i].g; len[i] = v[i].l; }
    return (int)v.size();
}

// ChrLocMap.lower_bound(g)->first (/root/reference/src/bwt_index.cpp:249-250): last coordinate of g's chromosome
static inline int64_t chr_end(const or_index *ix, int64_t g)
{
    auto it = std::lower_bound(ix->ends.begin(), ix->ends.end(), g);
    return it == ix->ends.end() ? ix->ends.back() : *it;
}

// GenerateAlignmentCandidate (/root/reference/src/AlignmentCandidates.cpp:241-288).
// The greedy scan only ever compares a seed with its predecessor in the sorted list (j == k-1 at every
// test), so clusters are the maximal runs between adjacent pairs that fail the chaining test.
extern "C" int or_cluster_read(const or_index *ix, int rlen, int n, const int32_t *rpos, const int64_t *gpos,
                               const int32_t *len, int max_gaps, int max_intron,
                               int32_t *c_begin, int32_t *c_count, int32_t *c_score, int cap)
{
    int thr = (int)(rlen * 0.3), nc = 0, i = 0;
    while (i < n && gpos[i] - rpos[i] < 0) i++;
    while (i < n) {
        int k = i + 1, score = len[i];
        for (; k < n; k++) {
            int64_t d = llabs((gpos[k] - rpos[k]) - (gpos[k - 1] - rpos[k - 1]));
            bool chain = d < max_gaps ||
                         (d < max_intron && gpos[k] < chr_end(ix, gpos[k - 1]) && rpos[k] > rpos[k - 1]);
            if (!chain) break;
            score += len[k];
        }
        if (score > thr) {
            if (nc < cap) { c_begin[nc] = i; c_count[nc] = k - i; c_score[nc] = score; }
            nc++;
        }
        i = k;
    }
    return nc;
}

// ---- 8-mer re-seeding -------------------------------------------------------------------------
static const unsigned char nt4(unsigned char c)
{
    switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2;
                 case 'T': case 't': return 3; case '-': return 5; default: return 4; }
}

// CreateKmerVecFromReadSeq + CreateKmerID (/root/reference/src/KmerAnalysis.cpp:25-80), unsorted.
// Quirks kept: only a literal 'N' breaks a word; other non-ACGT symbols add 4 into the rolling id; the
// first id after a (re)start is not masked; after an 'N' restart the window is one base late.
static void kmer_list(int len, const char *seq, std::vector<std::pair<uint32_t, uint32_t>> &out)
{
    out.clear();
    int tail = 0, count = 0, head;
    while (count < 8 && tail < len) { if (seq[tail++] != 'N') count++; else count = 0; }
    if (count != 8) return;
    auto fresh = [&](int h) { uint32_t id = 0; for (int i = h; i < h + 8; i++) id = (id << 2) + nt4(seq[i]); return id; };
    head = tail - 8;
    uint32_t wid = fresh(head);
    out.push_back({wid, (uint32_t)head});
    for (head += 1; tail < len; head++, tail++) {
        if (seq[tail] != 'N') {
            wid = ((wid & 0x3FFF) << 2) + nt4(seq[tail]);
            out.push_back({wid, (uint32_t)head});
        } else {
            count = 0; tail++;
            while (count < 8 && tail < len) { if (seq[tail++] != 'N') count++; else count = 0; }
            if (count != 8) break;
            head = tail - 8;
            wid = fresh(head);
            out.push_back({wid, (uint32_t)head});
        }
    }
}

// GenerateLongestSimplePairsFromFragmentPair (/root/reference/src/KmerAnalysis.cpp:82-106, :134-166).
// The reference sorts all common 8-mer pairs by (PosDiff, rPos) and walks equal-PosDiff runs; only the
// size, first rPos and last rPos of each run are used, so aggregate per diagonal and walk diagonals upward.
extern "C" void or_kmer_pair(or_index *ix, int len1, const char *f1, int len2, const char *f2, int64_t out3[3])
{
    if (ix) { ix->ctr.kmer_calls++; ix->ctr.kmer_window_bases += len2; ix->ctr.kmer_read_bases += len1; }
    out3[0] = out3[1] = out3[2] = 0;
    std::vector<std::pair<uint32_t, uint32_t>> a, b;
    kmer_list(len1, f1, a);
    kmer_list(len2, f2, b);
    if (a.empty() || b.empty()) return;
    std::sort(a.begin(), a.end());
    struct Diag { int count = 0; uint32_t rmin = 0, rmax = 0; };
    std::vector<Diag> diag((size_t)len1 + len2 + 2);
    for (auto &gk : b) {
        auto it = std::lower_bound(a.begin(), a.end(), std::make_pair(gk.first, 0u));
        for (; it != a.end() && it->first == gk.first; ++it) {
            int d = (int)(gk.second - it->second);
            Diag &D = diag[(size_t)(d + len1)];
            if (D.count == 0) D.rmin = D.rmax = it->second;
            else { D.rmin = std::min(D.rmin, it->second); D.rmax = std::max(D.rmax, it->second); }
            D.count++;
        }
    }
    int s = 1, max_len = 0;
    for (size_t di = 0; di < diag.size(); di++) {
        const Diag &D = diag[di];
        if (D.count == 0) continue;
        s += D.count - 1;
        int l = 8 + (int)(D.rmax - D.rmin);
        if (l > max_len && s > (l - 8) / 2) {
            out3[0] = D.rmin; out3[1] = (int64_t)D.rmin + ((int64_t)di - len1); out3[2] = l;
            max_len = l; s = 1;
        }
    }
}

// ---- Needleman-Wunsch ---------------------------------------------------------------------------
// nw_alignment (/root/reference/src/nw_alignment.cpp:18-82) in integer half-units: the
// reference's 3-way max resolves to double max(short,short,short), so each S is truncated toward zero to
// a whole unit while R/T keep half units; traceback tests S==R, then S==T, else diagonal.
extern "C" int or_nw(or_index *ix, int m, const char *s1, int n, const char *s2, uint8_t *ops)
{
    if (ix) { ix->ctr.nw_calls++; ix->ctr.nw_cells += (uint64_t)m * n; }
    const int W = n + 1, NEG = -131072;
    std::vector<int> S((size_t)(m + 1) * W), R(S.size()), T(S.size());
    S[0] = R[0] = T[0] = 0;
    for (int i = 1; i <= m; i++) { R[(size_t)i * W] = NEG; S[(size_t)i * W] = T[(size_t)i * W] = -2 - i; }
    for (int j = 1; j <= n; j++) { T[j] = NEG; S[j] = R[j] = -2 - j; }
    for (int i = 1; i <= m; i++)
        for (int j = 1; j <= n; j++) {
            size_t c = (size_t)i * W + j;
            R[c] = std::max(R[c - 1] - 1, S[c - 1] - 3);
            T[c] = std::max(T[c - W] - 1, S[c - W] - 3);
            unsigned char a = nt4(s1[i - 1]), b = nt4(s2[j - 1]);
            int h = std::max(S[c - W - 1] + (a == b ? 3 : -3), std::max(R[c], T[c]));
            S[c] = (h / 2) * 2; // truncation toward zero to a whole unit
        }
    int i = m, j = n, k = 0;
    while (i > 0 || j > 0) {
        size_t c = (size_t)i * W + j;
        if (S[c] == R[c]) { ops[k++] = 1; j--; }
        else if (S[c] == T[c]) { ops[k++] = 2; i--; }
        else { ops[k++] = 0; i--; j--; }
    }
    std::reverse(ops, ops + k);
    return k;
}

// ops -> the two gapped strings the reference produces in place
static void gapped_strings(int m, const char *s1, int n, const char *s2, const uint8_t *ops, int k,
                           std::string &a, std::string &b)
{
    a.clear(); b.clear();
    int i = 0, j = 0;
    for (int c = 0; c < k; c++) {
        if (ops[c] == 0) { a += s1[i++]; b += s2[j++]; }
        else if (ops[c] == 1) { a += '-'; b += s2[j++]; }
        else { a += s1[i++]; b += '-'; }
    }
    (void)m; (void)n;
}

// IdentifyBestGappedPartition (/root/reference/src/AlignmentCandidates.cpp:385-467)
extern "C" void or_gapped_partition(or_index *ix, const char *seq, int rGaps, int l_rpos, int l_rlen,
                                    int64_t l_gpos, int l_glen, int r_rpos, int64_t r_gpos, int max_mismatch,
                                    int out3[3])
{
    (void)r_rpos;
    static const char B[] = "ACGT";
    auto ref = [&](int64_t p) { return B[ref_code(ix, p)]; };
    std::string gap(seq + l_rpos + l_rlen, rGaps), g1(rGaps, 'A'), g2(rGaps, 'A'), a1, a2, a3, a4;
    for (int i = 0; i < rGaps; i++) { g1[i] = ref(l_gpos + l_glen + i); g2[i] = ref(r_gpos - rGaps + i); }
    std::vector<uint8_t> ops(2 * rGaps + 2);

    int k = or_nw(ix, rGaps, gap.data(), rGaps, g1.data(), ops.data());
    gapped_strings(rGaps, gap.data(), rGaps, g1.data(), ops.data(), k, a1, a2);
    { // genome gaps at the right end are re-filled with the bases that follow the window
        int i = k - 1; while (a2[i] == '-') i--;
        int64_t g = l_gpos + l_glen + rGaps;
        for (i += 1; i < k; i++, g++) a2[i] = ref(g);
    }
    std::vector<int> Rv(rGaps + 1, 0), Lv(rGaps + 1, 0);
    for (int p = 0, s = 0, i = 0; i < k; i++) { if (a1[i] == a2[i]) s++; if (a1[i] != '-') p++; Rv[p] = s; }

    int k2 = or_nw(ix, rGaps, gap.data(), rGaps, g2.data(), ops.data());
    gapped_strings(rGaps, gap.data(), rGaps, g2.data(), ops.data(), k2, a3, a4);
    { // genome gaps at the left end are re-filled walking left from the window start (sic)
        int i = 0; while (a4[i] == '-') i++;
        int64_t g = r_gpos - rGaps;
        for (i -= 1; i >= 0; i--, g--) a4[i] = ref(g);
    }
    for (int p = 0, s = 0, i = k2 - 1; i >= 0; i--) { if (a3[i] == a4[i]) s++; if (a3[i] != '-') p++; Lv[rGaps - p] = s; }

    int best = 0, P = 0;
    for (int i = 0; i <= rGaps; i++) if (Rv[i] + Lv[i] > best) { best = Rv[i] + Lv[i]; P = i; }
    int right_ext = 0, left_ext = 0;
    if (!(best < (int)(rGaps * 0.8) || (rGaps - best) > max_mismatch)) {
        for (int p = P, i = 0; p > 0; i++) { if (a1[i] != '-') p--; if (a2[i] != '-') right_ext++; }
        for (int p = rGaps - P, i = k2 - 1; p > 0; i--) { if (a3[i] != '-') p--; if (a4[i] != '-') left_ext++; }
    }
    out3[0] = P; out3[1] = left_ext; out3[2] = right_ext;
}
