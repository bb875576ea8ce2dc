// Whole-path orchestration behind dartgpu_map_reads(): everything ReadMapping() does per read between the
// reader and the SAM writer (/root/reference/src/Mapping.cpp:600-639), phase-split so that the nested,
// data-dependent NW and 8-mer calls of GenMappingReport (/root/reference/src/AlignmentCandidates.cpp:1079-1207)
// become three batched GPU launches over ALL candidates of the batch:
//
//   GPU  seeds + candidates (seed_kernels.cu)                       IdentifySeedPairs, GenerateAlignmentCandidate
//   host pair candidates, drop redundant ones                       Mapping.cpp:371-477
//   host phase A: tandem / translocation filters, collect windows   AlignmentCandidates.cpp:817-902, :685-700
//   GPU  8-mer re-seeding of every (read gap, genome window)        KmerAnalysis.cpp:134-166
//   host phase B: accept re-seeds, collect gap pairs                AlignmentCandidates.cpp:596-624, :577-594
//   GPU  NW of every gap against both flanks                        nw_alignment.cpp (call sites :395, :420)
//   host phase C: gapped partitions, splice-motif snapping, normal pairs, collect pair alignments
//                                                                   AlignmentCandidates.cpp:385-467, :732-815, :904-1035
//   GPU  NW of every non-simple pair                                tools.cpp:156, :220, :268
//   host phase D: CIGAR, score, coordinates; then per read: best/second best, mate rescue, flags, MAPQ,
//                 junction records                                  AlignmentCandidates.cpp:1119-1197, Mapping.cpp:74-206, :479-565
//
// The host phases are control flow over a handful of seeds per candidate; all per-base arithmetic on the hot
// path (FM search, locate, sort, clustering, 8-mer join, NW) runs in the CUDA kernels.  Host work is spread
// over OpenMP threads; candidates are independent until the final per-read pass.
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "context.h"

namespace dartgpu {

namespace {

struct Seed {           // SeedPair_t (/root/reference/src/structure.h:106-115)
    bool simple, acceptor;
    int rPos;
    int64_t gPos;
    int rLen, gLen;
    int64_t PosDiff;
};

inline bool by_genome_pos(const Seed &a, const Seed &b) { return a.gPos == b.gPos ? a.rPos < b.rPos : a.gPos < b.gPos; }

typedef std::vector<std::pair<int, char>> Cigar;

struct Cand {           // AlignmentCandidate_t + the state GenMappingReport keeps per candidate
    int Score = 0, SJtype = -1, PairedIdx = -1;
    int64_t PosDiff = 0;
    int64_t seed_begin = 0; int seed_count = 0;   // run inside the batch's sorted seed array
    std::vector<Seed> sv;
    bool live = false;            // went through the repair pipeline (Score > 0 on entry)
    bool skip = false;            // CheckCoordinateValidity failed
    // batched jobs
    std::vector<std::pair<int, int>> kjobs;   // (i, job id): re-seed between sv[i-1] and sv[i]
    std::vector<std::pair<int, int>> xjobs;   // (i, first of two NW jobs): gap between sv[i-1] and sv[i]
    std::vector<int> pjob;                    // per seed: NW job id, or -1
    int tidA = 0, tidB = 0, tidC = 0;         // which thread-local job list holds this candidate's jobs
    // result
    int AlnScore = 0, mis = 0;
    bool dir = false; int chr = 0; int64_t pos = 0;
    std::string cigar;
};

struct ReadState {
    int rlen = 0;
    const char *seq = nullptr;
    int64_t code_off = 0;          // offset of the read's codes on the device
    std::vector<Cand> cands;
    int score = 0, sub_score = 0, mis_num = 0, mapq = 0, CanNum = 0, best = 0;
    std::vector<int> AlnScore, SJtype, flag, paired;   // AlnReportArr
};

inline int chr_lookup(const dartgpu_ctx *c, int64_t g, int64_t *end_out)
{   // ChrLocMap.lower_bound(g)
    size_t k = std::lower_bound(c->ends.begin(), c->ends.end(), g) - c->ends.begin();
    if (k >= c->ends.size()) k = c->ends.size() - 1;
    if (end_out) *end_out = c->ends[k];
    return c->end_chr[k];
}

// ---- Mapping.cpp:371-401 ----
void remove_redundant(std::vector<Cand> &v)
{
    if (v.size() <= 1) return;
    int score1 = 0, score2 = 0;
    for (auto &a : v) {
        if (a.Score > score2) {
            if (a.Score >= score1) { score2 = score1; score1 = a.Score; }
            else score2 = a.Score;
        } else if (a.Score == score2) score2 = score1;
    }
    int thr = (score1 == score2 || score1 - score2 > 20) ? score1 : score2;
    for (auto &a : v) if (a.Score < thr) a.Score = 0;
}

// ---- Mapping.cpp:403-450 ----
bool pair_candidates(std::vector<Cand> &v1, std::vector<Cand> &v2)
{
    bool pairing = false;
    int num1 = (int)v1.size(), num2 = (int)v2.size();
    if (num1 * num2 > 1000) { remove_redundant(v1); remove_redundant(v2); }
    for (int i = 0; i != num1; i++) {
        if (v1[i].Score == 0) continue;
        int best_mate = -1;
        int64_t min_dist = 2000000;
        for (int j = 0; j != num2; j++) {
            if (v2[j].Score == 0 || v2[j].PosDiff < v1[i].PosDiff) continue;
            int64_t dist = std::llabs(v2[j].PosDiff - v1[i].PosDiff);
            if (dist < min_dist) { best_mate = j; min_dist = dist; }
        }
        if (best_mate != -1) {
            int j = best_mate;
            if (v2[j].PairedIdx == -1) { pairing = true; v1[i].PairedIdx = j; v2[j].PairedIdx = i; }
            else if (v1[i].Score > v1[v2[j].PairedIdx].Score) {
                v1[v2[j].PairedIdx].PairedIdx = -1;
                v1[i].PairedIdx = j; v2[j].PairedIdx = i;
            }
        }
    }
    return pairing;
}

// ---- Mapping.cpp:452-468 ----
void remove_unmated(std::vector<Cand> &v1, std::vector<Cand> &v2)
{
    for (auto &a : v1) {
        if (a.PairedIdx == -1) a.Score = 0;
        else { int j = a.PairedIdx; a.Score = v2[j].Score = a.Score + v2[j].Score; }
    }
    for (auto &b : v2) if (b.PairedIdx == -1) b.Score = 0;
}

void remove_null(std::vector<Seed> &sv)
{
    sv.erase(std::remove_if(sv.begin(), sv.end(), [](const Seed &s) { return s.rLen == 0; }), sv.end());
}

inline bool by_first(const std::pair<int, int> &a, const std::pair<int, int> &b) { return a.first < b.first; }

// ---- AlignmentCandidates.cpp:817-842 ----
void remove_tandem(std::vector<Seed> &sv)
{
    int num = (int)sv.size();
    if (num < 2) return;
    std::vector<std::pair<int, int>> vec(num);
    for (int i = 0; i < num; i++) vec[i] = {sv[i].rPos, i};
    std::sort(vec.begin(), vec.end(), by_first);
    bool any = false;
    for (int i = 0; i < num;) {
        int j = i + 1;
        while (j < num && vec[j].first == vec[i].first) j++;
        if (j - i > 1) { any = true; for (int k = i; k < j; k++) sv[vec[k].second].rLen = sv[vec[k].second].gLen = 0; }
        i = j;
    }
    if (any) remove_null(sv);
}

// ---- AlignmentCandidates.cpp:844-902 ----
void remove_translocated(std::vector<Seed> &sv)
{
    int num = (int)sv.size();
    if (num < 2) return;
    std::vector<std::pair<int, int>> vec(num);
    for (int i = 0; i < num; i++) vec[i] = {sv[i].rPos, i};
    std::sort(vec.begin(), vec.end(), by_first);
    bool any = false;
    for (int i = 0; i < num; i++) {
        if (vec[i].first != sv[i].rPos) {
            any = true;
            int j = vec[i].second;
            for (int t = i + 1; t <= j; t++) if (vec[t].second > j) j = vec[t].second;
            int s1 = 0, s2 = 0;
            for (int k = i; k <= j; k++) { if (k < vec[k].second) s1 += sv[vec[k].second].rLen; else s2 += sv[vec[k].second].rLen; }
            if (s1 > s2) { for (int k = i; k <= j; k++) if (k > vec[k].second) sv[vec[k].second].rLen = sv[vec[k].second].gLen = 0; }
            else { for (int k = i; k <= j; k++) if (k < vec[k].second) sv[vec[k].second].rLen = sv[vec[k].second].gLen = 0; }
            i = j;
        }
    }
    if (any) remove_null(sv);
}

// ---- AlignmentCandidates.cpp:702-756 ----
bool same_fragment(const dartgpu_ctx *c, int64_t L, int64_t R, int shift)
{
    if (shift > 0) { for (int i = 0; i < shift; i++, L++, R++) if (host_ref_code(c, L) != host_ref_code(c, R)) return false; }
    else { shift = -shift; L -= shift; R -= shift; for (int i = 0; i < shift; i++, L++, R++) if (host_ref_code(c, L) != host_ref_code(c, R)) return false; }
    return true;
}

const char *const kMotif[4] = {"GT/AG", "CT/AC", "GC/AG", "CT/GC"};   // main.cpp:18
const int kShift[19] = {0, 1, -1, 2, -2, 3, -3, 4, -4, 5, -5, 6, -6, 7, -7, 8, -8, 9, -9};

int find_junction(const dartgpu_ctx *c, int type, const Seed &l, const Seed &r)
{
    int i = std::min(l.rLen, r.rLen), j = std::min(l.gLen, r.gLen);
    if (i < j) j = i;
    if (j > 9) j = 9;
    j <<= 1;
    int64_t L = l.gPos + l.gLen, R = r.gPos;
    int shift = 0;
    for (i = 0; i <= j; i++) {
        shift = kShift[i];
        if (shift != 0 && !same_fragment(c, L, R, shift)) continue;
        int64_t g1 = L + shift, g2 = R - 2 + shift;
        if (host_ref_char(c, g1) == kMotif[type][0] && host_ref_char(c, g1 + 1) == kMotif[type][1] &&
            host_ref_char(c, g2) == kMotif[type][3] && host_ref_char(c, g2 + 1) == kMotif[type][4]) break;
    }
    return i > j ? 10 : shift;
}

// ---- AlignmentCandidates.cpp:758-815 ----
int check_splice_junction(const dartgpu_ctx *c, std::vector<Seed> &sv)
{
    int num = (int)sv.size(), min_cost = 1000, best_type = -1;
    std::vector<std::pair<int, int>> vec, best_vec;
    for (int type = 0; type < 4; type++) {
        vec.clear();
        int mis = 0, cost = 0;
        for (int i = 1; i < num; i++) {
            if ((sv[i].PosDiff - sv[i - 1].PosDiff) > c->prm.min_intron && sv[i - 1].simple && sv[i].simple) {
                int shift = find_junction(c, type, sv[i - 1], sv[i]);
                if (shift != 10) vec.push_back({i, shift}); else mis++;
                cost += std::abs(shift);
            }
        }
        if (!vec.empty() && cost < min_cost) { min_cost = cost; best_type = type; best_vec = vec; }
        if (mis == 0) break;
    }
    if (best_type != -1) {
        for (auto &p : best_vec) {
            int j = p.first, shift = p.second;
            if (shift != 10) {
                sv[j].acceptor = true;
                if (shift != 0) {
                    sv[j - 1].rLen += shift; sv[j - 1].gLen += shift;
                    sv[j].rLen -= shift; sv[j].gLen -= shift;
                    sv[j].rPos += shift; sv[j].gPos += shift;
                }
            }
        }
    }
    return best_type;
}

// ---- AlignmentCandidates.cpp:904-954 ----
bool seed_overlap(Seed &p1, Seed &p2)
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

// ---- AlignmentCandidates.cpp:956-999 ----
void check_overlapping(std::vector<Seed> &sv)
{
    int num = (int)sv.size();
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
    if (null_seed) remove_null(sv);
}

// ---- AlignmentCandidates.cpp:1001-1035 ----
void identify_normal_pairs(std::vector<Seed> &sv)
{
    if (sv.size() <= 1) return;
    check_overlapping(sv);
    int num = (int)sv.size();
    Seed sp; sp.acceptor = sp.simple = false;
    for (int i = 0, j = 1; j < num; i++, j++) {
        if (sv[j].rPos - sv[i].rPos - sv[i].rLen == 0) continue;
        int rGaps = sv[j].rPos - (sv[i].rPos + sv[i].rLen); if (rGaps < 0) rGaps = 0;
        int gGaps = (int)(sv[j].gPos - (sv[i].gPos + sv[i].gLen)); if (gGaps < 0) gGaps = 0; else if (gGaps > 30 && gGaps > (rGaps << 1)) gGaps = 0;
        if (rGaps > 0 || gGaps > 0) {
            sp.rPos = sv[i].rPos + sv[i].rLen; sp.gPos = sv[i].gPos + sv[i].gLen;
            sp.PosDiff = sp.gPos - sp.rPos; sp.rLen = rGaps; sp.gLen = gGaps;
            sv.push_back(sp);
        }
    }
    if ((int)sv.size() > num) std::inplace_merge(sv.begin(), sv.begin() + num, sv.end(), by_genome_pos);
}

// ---- AlignmentCandidates.cpp:136-163 ----
bool coordinates_valid(const dartgpu_ctx *c, const std::vector<Seed> &sv)
{
    int64_t g1 = 0, g2 = 2 * c->G;
    for (auto it = sv.begin(); it != sv.end(); ++it) if (it->gLen > 0) { g1 = it->gPos; break; }
    for (auto it = sv.rbegin(); it != sv.rend(); ++it) if (it->gLen > 0) { g2 = it->gPos + it->gLen - 1; break; }
    return !((g1 < c->G && g2 >= c->G) || (g1 >= c->G && g2 < c->G));
}

// gapped strings of one NW job: a = read side, b = genome side ('-' marks gaps)
void gapped(const dartgpu_ctx *c, const char *s1, int64_t gpos, const uint8_t *ops, int k, std::string &a, std::string &b)
{
    a.resize(k); b.resize(k);
    int i = 0; int64_t g = gpos;
    for (int t = 0; t < k; t++) {
        if (ops[t] == 0) { a[t] = s1[i++]; b[t] = host_ref_char(c, g++); }
        else if (ops[t] == 1) { a[t] = '-'; b[t] = host_ref_char(c, g++); }
        else { a[t] = s1[i++]; b[t] = '-'; }
    }
}

// ---- tools.cpp:49-104 ----
int add_cigar(const std::string &s1, const std::string &s2, Cigar &cv)
{
    char state = '*';
    int c = 0, score = 0, len = (int)s1.length();
    for (int i = 0; i < len; i++) {
        char want;
        if (s1[i] == '-') want = 'D';
        else if (s2[i] == '-') want = 'I';
        else { want = 'M'; if (s1[i] == s2[i]) score++; }
        if (state == want) c++;
        else { if (c > 0) cv.push_back({c, state}); c = 1; state = want; }
    }
    if (c > 0) cv.push_back({c, state});
    return score;
}

// ---- tools.cpp:166-201 ----
bool local_quality_ok(const std::string &a1, const std::string &a2)
{
    int type = -1, n = 0, mis = 0, status = 0, len = (int)a1.length();
    for (int i = 0; i < len; i++) {
        int t;
        if (a1[i] == '-') t = 0;
        else if (a2[i] == '-') t = 1;
        else { t = 2; n++; if (a1[i] != a2[i]) mis++; }
        if (type != t) { type = t; status++; }
    }
    return !(status >= 4 || (mis >= 3 && mis >= (int)(n * 0.3)));
}

// the <=2-mismatch fast path shared by tools.cpp:149, :213, :261
inline bool simple_enough(const dartgpu_ctx *c, const char *seq, const Seed &sp, int *n_out)
{
    if (sp.rLen != sp.gLen) return false;
    int n = 0;
    for (int i = 0; i < sp.rLen; i++) if (seq[sp.rPos + i] != host_ref_char(c, sp.gPos + i)) n++;
    *n_out = n;
    return n <= 2 && n <= (int)(sp.rLen * 0.2);
}

// ---- AlignmentCandidates.cpp:37-61 ----
std::string cigar_string(const Cigar &cv)
{
    std::string out;
    char state = '\0', buf[16];
    int c = 0;
    for (size_t i = 0; i < cv.size(); i++) {
        if (cv[i].second != state) {
            if (c > 0) { snprintf(buf, sizeof buf, "%d%c", c, state); out += buf; }
            c = cv[i].first; state = cv[i].second;
        } else c += cv[i].first;
    }
    if (c > 0) { snprintf(buf, sizeof buf, "%d%c", c, state); out += buf; }
    return out;
}

} // namespace

// =====================================================================================================
void run_pipeline(dartgpu_ctx *c, const dartgpu_reads *reads, dartgpu_map_result *out)
{
    const int n = c->n_reads;
    const dartgpu_params &P = c->prm;
    const int threads = P.host_threads > 0 ? P.host_threads : omp_get_max_threads();
    const bool paired = P.pair_end != 0;
    std::vector<ReadState> R(n);

    // ---- candidates from the GPU stage; pairing and pruning (Mapping.cpp:603-610, :631-633) ----
#pragma omp parallel for schedule(dynamic, 256) num_threads(threads)
    for (int u = 0; u < (paired ? n / 2 : n); u++) {
        for (int m = 0; m < (paired ? 2 : 1); m++) {
            int i = paired ? 2 * u + m : u;
            ReadState &r = R[i];
            r.rlen = c->h_rlen.p[i];
            r.seq = reads->bases + reads->offsets[i];
            r.code_off = c->h_dev_off.p[i];
            int64_t so = c->h_seed_off.p[i];
            uint32_t nc = c->h_ncand.p[i];
            r.cands.resize(nc);
            for (uint32_t k = 0; k < nc; k++) {
                Cand &a = r.cands[k];
                a.Score = c->h_cand_score.p[so + k];
                a.seed_begin = so + c->h_cand_begin.p[so + k];
                a.seed_count = c->h_cand_count.p[so + k];
                uint64_t key = c->h_keys.p[a.seed_begin];
                a.PosDiff = std::max<int64_t>(key_gpos(key) - key_rpos(key), 0);
            }
        }
        if (paired) {
            auto &v1 = R[2 * u].cands, &v2 = R[2 * u + 1].cands;
            if (pair_candidates(v1, v2)) remove_unmated(v1, v2);
            remove_redundant(v1); remove_redundant(v2);
        } else remove_redundant(R[u].cands);
    }

    // live candidates, flattened
    std::vector<std::pair<int, int>> live;
    for (int i = 0; i < n; i++)
        for (int k = 0; k < (int)R[i].cands.size(); k++)
            if (R[i].cands[k].Score != 0) live.push_back({i, k});
    const int64_t nlive = (int64_t)live.size();

    // ---- phase A ----
    std::vector<std::vector<KmerJobDev>> kj_t(threads);
#pragma omp parallel num_threads(threads)
    {
        auto &kj = kj_t[omp_get_thread_num()];
#pragma omp for schedule(static)
        for (int64_t w = 0; w < nlive; w++) {
            ReadState &r = R[live[w].first];
            Cand &a = r.cands[live[w].second];
            a.live = true; a.tidA = omp_get_thread_num();
            a.sv.resize(a.seed_count);
            for (int s = 0; s < a.seed_count; s++) {
                uint64_t key = c->h_keys.p[a.seed_begin + s];
                Seed &sd = a.sv[s];
                sd.simple = true; sd.acceptor = false;
                sd.rPos = key_rpos(key); sd.gPos = key_gpos(key); sd.rLen = sd.gLen = key_len(key);
                sd.PosDiff = sd.gPos - sd.rPos;
            }
            remove_tandem(a.sv);
            remove_translocated(a.sv);
            // IdentifyMissingSeeds (AlignmentCandidates.cpp:685-700): windows between distant adjacent seeds
            int num = (int)a.sv.size();
            for (int i = 1; i < num; i++) {
                int rGaps;
                if ((int)(a.sv[i].PosDiff - a.sv[i - 1].PosDiff) > P.max_gaps &&
                    (rGaps = a.sv[i].rPos - a.sv[i - 1].rPos - a.sv[i - 1].rLen) > 20) {
                    int rBegin = a.sv[i - 1].rPos + a.sv[i - 1].rLen;
                    int64_t Lb = a.sv[i - 1].gPos + a.sv[i - 1].gLen, Rb = a.sv[i].gPos;
                    KmerJobDev j; j.s1_off = r.code_off + rBegin; j.gpos = Lb; j.len1 = rGaps; j.len2 = (int32_t)(Rb - Lb);
                    a.kjobs.push_back({i, (int)kj.size()});   // thread-local id, rebased below
                    kj.push_back(j);
                }
            }
        }
    }
    // concatenate per-thread job lists (static schedule => thread t owns a contiguous range of `live`)
    std::vector<int> kbase(threads + 1, 0);
    for (int t = 0; t < threads; t++) kbase[t + 1] = kbase[t] + (int)kj_t[t].size();
    c->h_kjobs.reserve(kbase[threads] + 1);
    int max_len1 = 8;
    for (int t = 0; t < threads; t++)
        for (size_t k = 0; k < kj_t[t].size(); k++) {
            c->h_kjobs.p[kbase[t] + k] = kj_t[t][k];
            if (kj_t[t][k].len2 < 0) c->h_kjobs.p[kbase[t] + k].len2 = 0;
            max_len1 = std::max(max_len1, kj_t[t][k].len1);
        }
    run_kmer(c, c->d_codes.p, c->h_kjobs.p, kbase[threads], max_len1);

    // ---- phase B ----
    std::vector<std::vector<NwJobDev>> nj_t(threads);
#pragma omp parallel num_threads(threads)
    {
        const int tid = omp_get_thread_num();
        auto &nj = nj_t[tid];
#pragma omp for schedule(static)
        for (int64_t w = 0; w < nlive; w++) {
            ReadState &r = R[live[w].first];
            Cand &a = r.cands[live[w].second];
            // ReseedingWithSpecificRegion's acceptance test (AlignmentCandidates.cpp:609-618)
            bool added = false;
            a.tidB = tid;
            for (auto &kjb : a.kjobs) {
                const KmerJobDev &J = c->h_kjobs.p[kbase[a.tidA] + kjb.second];
                const dartgpu_kmer_hit &h = c->h_khits.p[kbase[a.tidA] + kjb.second];
                int thr = (int)(J.len1 * 0.85); if (thr < 8) thr = 8;
                if (h.len >= thr && h.len > 0) {
                    Seed sd; sd.simple = true; sd.acceptor = false;
                    sd.rPos = h.rpos + (int)(J.s1_off - r.code_off);
                    sd.gPos = (int64_t)h.gpos + J.gpos;
                    sd.rLen = sd.gLen = h.len;
                    sd.PosDiff = sd.gPos - sd.rPos;
                    a.sv.push_back(sd); added = true;
                }
            }
            if (added) std::sort(a.sv.begin(), a.sv.end(), by_genome_pos);
            // SeedExtension (AlignmentCandidates.cpp:577-594): every gap is aligned against both flanks
            int num = (int)a.sv.size();
            for (int i = 1; i < num; i++) {
                if ((int)(a.sv[i].PosDiff - a.sv[i - 1].PosDiff) > P.min_intron && a.sv[i].rPos > (a.sv[i - 1].rPos + a.sv[i - 1].rLen)) {
                    int rGaps = a.sv[i].rPos - (a.sv[i - 1].rPos + a.sv[i - 1].rLen);
                    int64_t s1 = r.code_off + a.sv[i - 1].rPos + a.sv[i - 1].rLen;
                    a.xjobs.push_back({i, (int)nj.size()});
                    nj.push_back(NwJobDev{s1, a.sv[i - 1].gPos + a.sv[i - 1].gLen, 0, 0, 0, rGaps, rGaps});
                    nj.push_back(NwJobDev{s1, a.sv[i].gPos - rGaps, 0, 0, 0, rGaps, rGaps});
                }
            }
        }
    }
    std::vector<int> nbase(threads + 1, 0);
    for (int t = 0; t < threads; t++) nbase[t + 1] = nbase[t] + (int)nj_t[t].size();
    c->h_njobs.reserve(nbase[threads] + 1);
    for (int t = 0; t < threads; t++)
        if (!nj_t[t].empty()) memcpy(c->h_njobs.p + nbase[t], nj_t[t].data(), nj_t[t].size() * sizeof(NwJobDev));
    run_nw(c, c->d_codes.p, c->h_njobs.p, nbase[threads]);
    // keep this round's results: phase C reads them while collecting the next round's jobs
    std::vector<int64_t> x_off = c->o_op_off;
    std::vector<uint8_t> x_ops = c->o_ops;
    std::vector<NwJobDev> x_jobs(c->h_njobs.p, c->h_njobs.p + nbase[threads]);
    const std::vector<int> xbase = nbase;

    // ---- phase C ----
    for (auto &v : nj_t) v.clear();
#pragma omp parallel num_threads(threads)
    {
        const int tid = omp_get_thread_num();
        auto &nj = nj_t[tid];
        std::string a1, a2, a3, a4;
        std::vector<int> Rv, Lv;
#pragma omp for schedule(static)
        for (int64_t w = 0; w < nlive; w++) {
            ReadState &r = R[live[w].first];
            Cand &a = r.cands[live[w].second];
            // FillGapsBetweenAdjacentSeeds / IdentifyBestGappedPartition (AlignmentCandidates.cpp:385-467, :547-575)
            std::vector<Seed> extra;
            a.tidC = tid;
            for (auto &xj : a.xjobs) {
                const int i = xj.first, j1 = xbase[a.tidB] + xj.second, j2 = j1 + 1;
                const Seed &L = a.sv[i - 1], &Rt = a.sv[i];
                const int rGaps = x_jobs[j1].m;
                const char *gap = r.seq + L.rPos + L.rLen;
                int k1 = (int)(x_off[j1 + 1] - x_off[j1]), k2 = (int)(x_off[j2 + 1] - x_off[j2]);
                gapped(c, gap, x_jobs[j1].gpos, x_ops.data() + x_off[j1], k1, a1, a2);
                gapped(c, gap, x_jobs[j2].gpos, x_ops.data() + x_off[j2], k2, a3, a4);
                { int t = k1 - 1; while (a2[t] == '-') t--; int64_t g = L.gPos + L.gLen + rGaps; for (t += 1; t < k1; t++, g++) a2[t] = host_ref_char(c, g); }
                Rv.assign(rGaps + 1, 0); Lv.assign(rGaps + 1, 0);
                for (int p = 0, s = 0, t = 0; t < k1; t++) { if (a1[t] == a2[t]) s++; if (a1[t] != '-') p++; Rv[p] = s; }
                { int t = 0; while (a4[t] == '-') t++; int64_t g = Rt.gPos - rGaps; for (t -= 1; t >= 0; t--, g--) a4[t] = host_ref_char(c, g); }
                for (int p = 0, s = 0, t = k2 - 1; t >= 0; t--) { if (a3[t] == a4[t]) s++; if (a3[t] != '-') p++; Lv[rGaps - p] = s; }
                int best = 0, Pp = 0;
                for (int t = 0; t <= rGaps; t++) if (Rv[t] + Lv[t] > best) { best = Rv[t] + Lv[t]; Pp = t; }
                int right_ext = 0, left_ext = 0;
                if (!(best < (int)(rGaps * 0.8) || (rGaps - best) > P.max_mismatch)) {
                    for (int p = Pp, t = 0; p > 0; t++) { if (a1[t] != '-') p--; if (a2[t] != '-') right_ext++; }
                    for (int p = rGaps - Pp, t = k2 - 1; p > 0; t--) { if (a3[t] != '-') p--; if (a4[t] != '-') left_ext++; }
                }
                Seed sd; sd.acceptor = false; sd.simple = false;
                int rest = rGaps;
                if (Pp > 0) {
                    sd.rPos = L.rPos + L.rLen; sd.gPos = L.gPos + L.gLen; sd.PosDiff = sd.gPos - sd.rPos;
                    sd.rLen = Pp; sd.gLen = right_ext;
                    extra.push_back(sd);
                }
                if ((rest -= Pp) > 0) {
                    sd.rLen = rest; sd.gLen = left_ext;
                    sd.rPos = Rt.rPos - sd.rLen; sd.gPos = Rt.gPos - sd.gLen; sd.PosDiff = sd.gPos - sd.rPos;
                    extra.push_back(sd);
                }
            }
            if (!extra.empty()) {
                a.sv.insert(a.sv.end(), extra.begin(), extra.end());
                std::sort(a.sv.begin(), a.sv.end(), by_genome_pos);
            }
            a.SJtype = check_splice_junction(c, a.sv);
            identify_normal_pairs(a.sv);
            int num = (int)a.sv.size();
            if (num > 1 && !coordinates_valid(c, a.sv)) { a.skip = true; continue; }
            // which non-simple pairs need NW (tools.cpp:130-164, :203-249, :251-300)
            a.pjob.assign(num, -1);
            for (int j = 0; j < num; j++) {
                const Seed &sp = a.sv[j];
                if ((sp.rLen == 0 && sp.gLen == 0) || sp.simple) continue;
                bool middle = !(j == 0 || j == num - 1);
                if (middle && (sp.PosDiff == -1 || sp.rLen == 0 || sp.gLen == 0)) continue;
                int nm;
                if (simple_enough(c, r.seq, sp, &nm)) continue;
                a.pjob[j] = (int)nj.size();
                nj.push_back(NwJobDev{r.code_off + sp.rPos, sp.gPos, 0, 0, 0, sp.rLen, sp.gLen});
            }
        }
    }
    for (int t = 0; t < threads; t++) nbase[t + 1] = nbase[t] + (int)nj_t[t].size();
    c->h_njobs.reserve(nbase[threads] + 1);
    for (int t = 0; t < threads; t++)
        if (!nj_t[t].empty()) memcpy(c->h_njobs.p + nbase[t], nj_t[t].data(), nj_t[t].size() * sizeof(NwJobDev));
    run_nw(c, c->d_codes.p, c->h_njobs.p, nbase[threads]);

    // ---- phase D: CIGAR / score / coordinates per candidate (AlignmentCandidates.cpp:1119-1184) ----
#pragma omp parallel num_threads(threads)
    {
        std::string f1, f2;
        Cigar cv;
#pragma omp for schedule(static)
        for (int64_t w = 0; w < nlive; w++) {
            ReadState &r = R[live[w].first];
            Cand &a = r.cands[live[w].second];
            if (a.skip) continue;
            const int num = (int)a.sv.size();
            cv.clear();
            int mis = 0, aln = 0;
            for (int j = 0; j != num; j++) {
                Seed &sp = a.sv[j];
                if (sp.rLen == 0 && sp.gLen == 0) continue;
                int g;
                if (j > 0 && (g = (int)(sp.gPos - (a.sv[j - 1].gPos + a.sv[j - 1].gLen))) > 0) cv.push_back({g, 'N'});
                if (sp.simple) { cv.push_back({sp.rLen, 'M'}); aln += sp.rLen; continue; }
                const bool head = j == 0, tail = !head && j == num - 1;
                int score = 0;
                if (!head && !tail && sp.PosDiff == -1) cv.push_back({sp.rLen, 'S'});
                else if (!head && !tail && (sp.rLen == 0 || sp.gLen == 0)) {
                    if (sp.rLen > 0) cv.push_back({sp.rLen, 'I'}); else if (sp.gLen > 0) cv.push_back({sp.gLen, 'D'});
                } else if (a.pjob[j] < 0) {
                    int nm = 0;
                    simple_enough(c, r.seq, sp, &nm);
                    score = sp.rLen - nm;
                    cv.push_back({sp.rLen, 'M'});
                } else {
                    const int jid = nbase[a.tidC] + a.pjob[j];
                    const int k = (int)(c->o_op_off[jid + 1] - c->o_op_off[jid]);
                    gapped(c, r.seq + sp.rPos, sp.gPos, c->o_ops.data() + c->o_op_off[jid], k, f1, f2);
                    if (!head && !tail) score = add_cigar(f1, f2, cv);
                    else if (!local_quality_ok(f1, f2)) { cv.push_back({sp.rLen, 'S'}); score = 0; }
                    else if (head) {   // tools.cpp:228-246
                        int p = 0; while (f1[p] == '-') p++;
                        if (p > 0) { f1.erase(0, p); f2.erase(0, p); sp.gPos += p; sp.gLen -= p; }
                        p = 0; while (f2[p] == '-') p++;
                        if (p > 0) { f1.erase(0, p); f2.erase(0, p); sp.rPos += p; sp.rLen -= p; cv.push_back({p, 'S'}); }
                        score = add_cigar(f1, f2, cv);
                    } else {           // tools.cpp:276-297
                        int p = (int)f1.length() - 1, cnt = 0; while (p >= 0 && f1[p--] == '-') cnt++;
                        if (cnt > 0) { f1.resize(f1.length() - cnt); f2.resize(f2.length() - cnt); sp.gLen -= cnt; }
                        p = (int)f2.length() - 1; cnt = 0; while (p >= 0 && f2[p--] == '-') cnt++;
                        if (cnt > 0) { f1.resize(f1.length() - cnt); f2.resize(f2.length() - cnt); sp.rLen -= cnt; }
                        score = add_cigar(f1, f2, cv);
                        if (cnt > 0) cv.push_back({cnt, 'S'});
                    }
                }
                aln += score;
                mis += sp.rLen - score;   // after the head/tail helper may have shrunk rLen, as the reference does
            }
            if (num > 0) {
                int j;
                if ((j = a.sv.front().rPos) > 0) cv.insert(cv.begin(), {j, 'S'});
                if ((j = r.rlen - (a.sv.back().rPos + a.sv.back().rLen)) > 0) cv.push_back({j, 'S'});
            }
            if (mis > P.max_mismatch || cv.empty()) aln = 0;
            for (auto &e : cv) if (e.second == 'N' && e.first < P.min_intron) { aln = 0; break; }
            a.mis = mis;
            if (aln > 0) {
                // GenCoordinateInfo (AlignmentCandidates.cpp:83-116)
                const bool first = !paired || (live[w].first & 1) == 0;
                int64_t gPos = a.sv.front().gPos, end_gPos = a.sv.back().gPos + a.sv.back().gLen - 1, endc;
                a.chr = chr_lookup(c, gPos, &endc);
                if (gPos < c->G) { a.dir = first; a.pos = gPos + 1 - c->chr_fwd[a.chr]; }
                else { a.dir = !first; a.pos = endc - end_gPos + 1; }
                if (a.pos <= 0) aln = 0;
                else {
                    if (a.sv.front().gPos >= c->G) std::reverse(cv.begin(), cv.end());
                    a.cigar = cigar_string(cv);
                }
            }
            a.AlnScore = aln;
        }
    }

    // ---- per read: best / second best (AlignmentCandidates.cpp:1185-1206) ----
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int i = 0; i < n; i++) {
        ReadState &r = R[i];
        r.score = r.best = 0; r.sub_score = 0; r.mis_num = 0;
        int cn = (int)r.cands.size();
        if (cn > 0) {
            r.CanNum = cn;
            r.AlnScore.assign(cn, 0); r.SJtype.assign(cn, -1); r.flag.assign(cn, 0); r.paired.assign(cn, -1);
            for (int k = 0; k < cn; k++) {
                Cand &a = r.cands[k];
                r.paired[k] = a.PairedIdx;
                if (!a.live) continue;
                r.SJtype[k] = a.SJtype;
                if (a.skip) continue;
                r.AlnScore[k] = a.AlnScore;
                if (a.AlnScore > 0) {
                    if (a.AlnScore > r.score) { r.best = k; r.mis_num = a.mis; r.sub_score = r.score; r.score = a.AlnScore; }
                    else if (a.AlnScore == r.score) r.sub_score = r.score;
                }
            }
        } else {
            r.CanNum = 1;
            r.AlnScore.assign(1, 0); r.SJtype.assign(1, -1); r.flag.assign(1, 0); r.paired.assign(1, -1);
        }
    }

    // ---- pairs: CheckPairedFinalAlignments, flags, MAPQ (Mapping.cpp:74-206, :479-530) ----
    auto mapq_of = [](ReadState &r) {
        if (r.score == 0 || r.score == r.sub_score) r.mapq = 0;
        else if (r.sub_score == 0 || r.score > r.sub_score) r.mapq = 50;
        else {
            int m = 0;
            for (int k = 0; k < r.CanNum; k++) if (r.AlnScore[k] == r.score) m++;
            r.mapq = m >= 10 ? 0 : m >= 4 ? 1 : m == 3 ? 2 : m == 2 ? 3 : 50;
        }
    };
    auto dir_of = [](const ReadState &r, int k) { return k < (int)r.cands.size() ? r.cands[k].dir : false; };
    const int units = paired ? n / 2 : n;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int u = 0; u < units; u++) {
        if (!paired) {
            ReadState &r = R[u];
            if (r.score > r.sub_score) { int k = r.best; r.flag[k] = dir_of(r, k) ? 0 : 0x10; }
            else if (r.score > 0) { for (int k = 0; k < r.CanNum; k++) if (r.AlnScore[k] > 0) r.flag[k] = dir_of(r, k) ? 0 : 0x10; }
            else r.flag[0] = 0x4;
            mapq_of(r);
            continue;
        }
        ReadState &r1 = R[2 * u], &r2 = R[2 * u + 1];
        {   // CheckPairedFinalAlignments
            bool mated = r1.paired[r1.best] == r2.best;
            if (!(!P.multi_hit && mated)) {
                if (!mated && r1.score > 0 && r2.score > 0) {
                    int s = 0;
                    for (int i = 0; i != r1.CanNum; i++) {
                        int j;
                        if (r1.AlnScore[i] > 0 && (j = r1.paired[i]) != -1 && r2.AlnScore[j] > 0) {
                            mated = true;
                            if (s < r1.AlnScore[i] + r2.AlnScore[j]) {
                                s = r1.AlnScore[i] + r2.AlnScore[j];
                                r1.best = i; r1.score = r1.AlnScore[i];
                                r2.best = j; r2.score = r2.AlnScore[j];
                            }
                        }
                    }
                }
                if (mated) {
                    for (int i = 0; i != r1.CanNum; i++) {
                        int j;
                        if (r1.AlnScore[i] != r1.score || ((j = r1.paired[i]) != -1 && r2.AlnScore[j] != r2.score)) {
                            r1.AlnScore[i] = 0; r1.paired[i] = -1;
                        }
                    }
                } else {
                    for (int i = 0; i != r1.CanNum; i++) {
                        if (r1.paired[i] != -1) r1.paired[i] = -1;
                        if (r1.AlnScore[i] > 0 && r1.AlnScore[i] != r1.score) r1.AlnScore[i] = 0;
                    }
                    for (int j = 0; j != r2.CanNum; j++) {
                        if (r2.paired[j] != -1) r2.paired[j] = -1;
                        if (r2.AlnScore[j] > 0 && r2.AlnScore[j] != r2.score) r2.AlnScore[j] = 0;
                    }
                }
            }
        }
        {   // SetPairedAlignmentFlag
            int i, j;
            if (r1.score > r1.sub_score && r2.score > r2.sub_score) {
                i = r1.best; j = r2.best;
                r1.flag[i] = 0x41; r2.flag[j] = 0x81;
                if (j == r1.paired[i]) { r1.flag[i] |= 0x2; r2.flag[j] |= 0x2; }
                r1.flag[i] |= dir_of(r1, i) ? 0x20 : 0x10;
                r2.flag[j] |= dir_of(r2, j) ? 0x20 : 0x10;
            } else {
                if (r1.score > r1.sub_score) {
                    i = r1.best;
                    r1.flag[i] = 0x41 | (dir_of(r1, i) ? 0x20 : 0x10);
                    if ((j = r1.paired[i]) != -1 && r2.AlnScore[j] > 0) r1.flag[i] |= 0x2; else r1.flag[i] |= 0x8;
                } else if (r1.score > 0) {
                    for (i = 0; i < r1.CanNum; i++) if (r1.AlnScore[i] > 0) {
                        r1.flag[i] = 0x41 | (dir_of(r1, i) ? 0x20 : 0x10);
                        if ((j = r1.paired[i]) != -1 && r2.AlnScore[j] > 0) r1.flag[i] |= 0x2; else r1.flag[i] |= 0x8;
                    }
                } else {
                    r1.flag[0] = 0x41 | 0x4;
                    if (r2.score == 0) r1.flag[0] |= 0x8; else r1.flag[0] |= dir_of(r2, r2.best) ? 0x10 : 0x20;
                }
                if (r2.score > r2.sub_score) {
                    j = r2.best;
                    r2.flag[j] = 0x81 | (dir_of(r2, j) ? 0x20 : 0x10);
                    if ((i = r2.paired[j]) != -1 && r1.AlnScore[i] > 0) r2.flag[j] |= 0x2; else r2.flag[j] |= 0x8;
                } else if (r2.score > 0) {
                    for (j = 0; j < r2.CanNum; j++) if (r2.AlnScore[j] > 0) {
                        r2.flag[j] = 0x81 | (dir_of(r2, j) ? 0x20 : 0x10);
                        if ((i = r2.paired[j]) != -1 && r1.AlnScore[i] > 0) r2.flag[j] |= 0x2; else r2.flag[j] |= 0x8;
                    }
                } else {
                    r2.flag[0] = 0x81 | 0x4;
                    if (r1.score == 0) r2.flag[0] |= 0x8; else r2.flag[0] |= dir_of(r1, r1.best) ? 0x10 : 0x20;
                }
            }
        }
        mapq_of(r1); mapq_of(r2);
    }

    // ---- flatten results; junction records in read order (UpdateLocalSJMap, Mapping.cpp:532-565) ----
    c->o_reads.resize(n);
    int64_t nrep = 0;
    for (int i = 0; i < n; i++) { c->o_reads[i].report_off = nrep; nrep += R[i].CanNum; }
    c->o_reports.resize(nrep);
    c->o_cigars.clear();
    c->o_junctions.clear();
    for (int i = 0; i < n; i++) {
        ReadState &r = R[i];
        dartgpu_read_result &o = c->o_reads[i];
        o.mapq = r.mapq; o.score = r.score; o.sub_score = r.sub_score; o.mis_num = r.mis_num;
        o.n_reports = r.CanNum; o.best = r.best;
        for (int k = 0; k < r.CanNum; k++) {
            dartgpu_report &p = c->o_reports[o.report_off + k];
            p.aln_score = r.AlnScore[k]; p.sj_type = r.SJtype[k]; p.flag = r.flag[k]; p.paired_idx = r.paired[k];
            p.dir = 0; p.chr_idx = 0; p.pos = 0; p.cigar_off = (int64_t)c->o_cigars.size(); p.cigar_len = 0; p.reserved = 0;
            if (k < (int)r.cands.size() && r.cands[k].live && !r.cands[k].skip && r.cands[k].AlnScore > 0) {
                const Cand &a = r.cands[k];
                p.dir = a.dir ? 1 : 0; p.chr_idx = a.chr; p.pos = a.pos;
                p.cigar_len = (int32_t)a.cigar.size();
                c->o_cigars.insert(c->o_cigars.end(), a.cigar.begin(), a.cigar.end());
            }
        }
        if ((r.mapq == 50 || (P.all_sj && r.score > 0)) && r.best < (int)r.cands.size()) {
            const Cand &a = r.cands[r.best];
            if (a.SJtype != -1) {
                for (int s = 1; s < (int)a.sv.size(); s++) {
                    if (!a.sv[s].acceptor) continue;
                    int64_t g1, g2;
                    if (a.PosDiff < c->G) { g1 = a.sv[s - 1].gPos + a.sv[s - 1].gLen; g2 = a.sv[s].gPos - 1; }
                    else { g1 = 2 * c->G - a.sv[s].gPos; g2 = 2 * c->G - 1 - (a.sv[s - 1].gPos + a.sv[s - 1].gLen); }
                    if (std::llabs(g2 - g1) < P.min_intron) continue;
                    c->o_junctions.push_back(dartgpu_junction{g1, g2, a.SJtype, i});
                }
            }
        }
    }
    out->reads = c->o_reads.data(); out->n_reads = n;
    out->reports = c->o_reports.data(); out->n_reports = nrep;
    out->cigars = c->o_cigars.data(); out->n_cigar_bytes = (int64_t)c->o_cigars.size();
    out->junctions = c->o_junctions.data(); out->n_junctions = (int64_t)c->o_junctions.size();
}

} // namespace dartgpu
