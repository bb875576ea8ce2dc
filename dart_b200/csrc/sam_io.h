// Reader + SAM-text writer stand-ins shared by dart_b200_map (GPU, through the C-ABI) and the CPU logic harness of
// tests/host: the reference's FASTQ reader (/root/reference/src/GetData.cpp:77-179), its SAM text records
// (/root/reference/src/Mapping.cpp:208-369, :741-751) and its junction table (/root/reference/src/Mapping.cpp:567-577, :683-716).
#pragma once
#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dartgpu.h"

struct Reads {
    std::vector<std::string> name;
    std::string seq, qual;            // concatenated, as held after load (mate 2 already flipped)
    std::vector<int64_t> off;         // n+1
    bool fastq = true;
};

static char comp(char c)
{   // GetComplementaryBase, /root/reference/src/tools.cpp:3-17
    switch (c) {
    case 'A': case 'a': return 'T';
    case 'C': case 'c': return 'G';
    case 'G': case 'g': return 'C';
    case 'T': case 't': return 'A';
    default: return 'N';
    }
}

static void revcomp(const char *s, int n, char *out)
{
    for (int i = 0; i < n; i++) out[i] = comp(s[n - 1 - i]);
}

// header = text between the leading '@'/'>' run and the first ' ', '/' or tab (GetData.cpp:55-75, :89-90)
static std::string parse_header(const char *b, int len)
{
    int p1 = len - 1, p2 = len - 1;
    for (int i = 1; i < len; i++) if (b[i] != '>' && b[i] != '@') { p1 = i; break; }
    for (int i = 1; i < len; i++) if (b[i] == ' ' || b[i] == '/' || b[i] == '\t') { p2 = i; break; }
    return p2 > p1 ? std::string(b + p1, p2 - p1) : std::string();
}

static bool slurp(const char *fn, std::string &buf)
{
    FILE *fp = fopen(fn, "rb");
    if (!fp) return false;
    fseek(fp, 0, SEEK_END);
    long n = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    buf.resize(n);
    size_t got = n ? fread(&buf[0], 1, n, fp) : 0;
    fclose(fp);
    return got == (size_t)n;
}

struct Rec { const char *h; int hl; const char *s; int sl; const char *q; int ql; };

static void parse_fastq(const std::string &buf, std::vector<Rec> &out)
{
    const char *p = buf.data(), *e = p + buf.size();
    auto line = [&](const char *&b, int &l) {
        if (p >= e) return false;
        const char *nl = (const char *)memchr(p, '\n', e - p);
        b = p; l = (int)((nl ? nl : e) - p) + (nl ? 1 : 0); // length including the newline, as getline reports
        p = nl ? nl + 1 : e;
        return true;
    };
    for (;;) {
        Rec r; const char *t; int tl;
        if (!line(r.h, r.hl)) break;
        if (!line(r.s, r.sl)) break;
        if (!line(t, tl)) break;
        if (!line(r.q, r.ql)) break;
        r.sl -= 1; // rlen = getline length - 1 (GetData.cpp:101)
        if (r.sl <= 0) break;
        out.push_back(r);
    }
}

// Interleaves the two files like GetNextChunk (mate 1, mate 2, ...); flips mate 2 when paired.
static bool load_reads(const char *f1, const char *f2, bool interleaved_pairs, Reads &R, bool &paired)
{
    std::string b1, b2;
    std::vector<Rec> r1, r2;
    if (!slurp(f1, b1)) { fprintf(stderr, "Cannot access file:[%s]\n", f1); return false; }
    parse_fastq(b1, r1);
    if (f2) {
        if (!slurp(f2, b2)) { fprintf(stderr, "Cannot access file:[%s]\n", f2); return false; }
        parse_fastq(b2, r2);
    }
    paired = f2 != nullptr || interleaved_pairs;
    std::vector<const Rec *> order;
    if (f2) { size_t n = std::min(r1.size(), r2.size()); for (size_t i = 0; i < n; i++) { order.push_back(&r1[i]); order.push_back(&r2[i]); } }
    else for (auto &r : r1) order.push_back(&r);
    if (paired && (order.size() & 1)) order.pop_back();
    size_t n = order.size();
    R.off.assign(n + 1, 0);
    for (size_t i = 0; i < n; i++) R.off[i + 1] = R.off[i] + order[i]->sl;
    R.seq.resize(R.off[n]); R.qual.resize(R.off[n]); R.name.resize(n);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        const Rec &r = *order[i];
        R.name[i] = parse_header(r.h, r.hl);
        char *s = &R.seq[R.off[i]], *q = &R.qual[R.off[i]];
        int ql = std::min(r.sl, r.ql);
        if (paired && (i & 1)) {   // GetData.cpp:157-168
            revcomp(r.s, r.sl, s);
            memset(q, 0, r.sl);
            for (int k = 0; k < ql; k++) q[r.sl - 1 - k] = r.q[k];
        } else {
            memcpy(s, r.s, r.sl);
            memset(q, 0, r.sl);
            memcpy(q, r.q, ql);
        }
    }
    return true;
}

static const char *XS_A[] = {"", " XS:A:+", " XS:A:-"};

struct Out { std::string sam; std::vector<dartgpu_junction> sj; int64_t n_unmapped = 0, n_unique = 0, n_paired = 0; };

// OutputPairedAlignments / OutputSingledAlignments (Mapping.cpp:208-369), one read (or pair) at a time
static void format_batch(const std::vector<std::string> &seq_names, const Reads &R, int64_t first, const dartgpu_map_result &res, const dartgpu_params &P, Out &o)
{
    const int n = res.n_reads;
    const bool paired = P.pair_end != 0;
    const int units = paired ? n / 2 : n;
    std::vector<std::string> chunk(units);
    std::vector<int> un(units, 0), uq(units, 0), pr(units, 0);
#pragma omp parallel for schedule(dynamic, 512)
    for (int u = 0; u < units; u++) {
        std::string &s = chunk[u];
        char num[64];
        auto emit_unmapped = [&](int64_t gi, const dartgpu_read_result &rr) {
            const dartgpu_report &p = res.reports[rr.report_off];
            int L = (int)(R.off[gi + 1] - R.off[gi]);
            s += R.name[gi]; snprintf(num, sizeof num, "\t%d\t*\t0\t0\t*\t*\t0\t0\t", p.flag); s += num;
            s.append(R.seq.data() + R.off[gi], L); s += '\t';
            if (R.fastq) s.append(R.qual.data() + R.off[gi], L); else s += '*';
            s += "\tAS:i:0\tXS:i:0\n";
        };
        auto emit_mapped = [&](int64_t gi, const dartgpu_read_result &rr, const dartgpu_report &p, bool print_stored,
                               const dartgpu_report *mate, int dist, int xs) {
            int L = (int)(R.off[gi + 1] - R.off[gi]);
            s += R.name[gi];
            snprintf(num, sizeof num, "\t%d\t", p.flag); s += num;
            s += seq_names[p.chr_idx];
            snprintf(num, sizeof num, "\t%lld\t%d\t", (long long)p.pos, rr.mapq); s += num;
            s.append(res.cigars + p.cigar_off, p.cigar_len);
            if (mate) { snprintf(num, sizeof num, "\t=\t%lld\t%d\t", (long long)mate->pos, dist); s += num; }
            else s += "\t*\t0\t0\t";
            const char *sq = R.seq.data() + R.off[gi], *ql = R.qual.data() + R.off[gi];
            if (print_stored) { s.append(sq, L); s += '\t'; if (R.fastq) s.append(ql, L); else s += '*'; }
            else {
                size_t at = s.size(); s.resize(at + L); revcomp(sq, L, &s[at]); s += '\t';
                if (R.fastq) { at = s.size(); s.resize(at + L); for (int k = 0; k < L; k++) s[at + k] = ql[L - 1 - k]; } else s += '*';
            }
            snprintf(num, sizeof num, "\tNM:i:%d\tAS:i:%d\tXS:i:%d%s\n", rr.mis_num, rr.score, rr.sub_score, XS_A[xs]); s += num;
        };
        if (!paired) {
            const int64_t gi = first + u;
            const dartgpu_read_result &rr = res.reads[u];
            if (rr.score == 0) { un[u]++; emit_unmapped(gi, rr); }
            else if (!P.unique || rr.mapq > 3) {
                if (rr.mapq == 50) uq[u]++;
                for (int i = rr.best; i < rr.n_reports; i++) {
                    const dartgpu_report &p = res.reports[rr.report_off + i];
                    if (p.aln_score == rr.score) {
                        int xs = p.sj_type == -1 ? 0 : (p.sj_type == 0 || p.sj_type == 2) ? 1 : 2;
                        emit_mapped(gi, rr, p, p.dir != 0, nullptr, 0, xs);
                        if (!P.multi_hit) break;
                    }
                }
            }
            continue;
        }
        const int64_t g1 = first + 2 * u, g2 = g1 + 1;
        const dartgpu_read_result &r1 = res.reads[2 * u], &r2 = res.reads[2 * u + 1];
        const int L1 = (int)(R.off[g1 + 1] - R.off[g1]), L2 = (int)(R.off[g2 + 1] - R.off[g2]);
        if (r1.score == 0) { un[u]++; emit_unmapped(g1, r1); }
        else if (!P.unique || r1.mapq > 3) {
            if (r1.mapq == 50) uq[u]++;
            for (int i = r1.best; i < r1.n_reports; i++) {
                const dartgpu_report &p = res.reports[r1.report_off + i];
                if (p.aln_score > 0) {
                    int xs = p.sj_type == -1 ? 0 : (p.sj_type == 0 || p.sj_type == 2) ? 1 : 2;
                    int j = p.paired_idx;
                    const dartgpu_report *m = (j != -1 && res.reports[r2.report_off + j].aln_score > 0) ? &res.reports[r2.report_off + j] : nullptr;
                    int dist = 0;
                    if (m) { dist = (int)(m->pos - p.pos + (p.dir ? L2 : 0 - L1)); if (i == r1.best) pr[u] += 2; }
                    emit_mapped(g1, r1, p, p.dir != 0, m, dist, xs);
                }
                if (!P.multi_hit) break;
            }
        }
        if (r2.score == 0) { un[u]++; emit_unmapped(g2, r2); }
        else if (!P.unique || r2.mapq > 3) {
            if (r2.mapq == 50) uq[u]++;
            for (int j = r2.best; j < r2.n_reports; j++) {
                const dartgpu_report &p = res.reports[r2.report_off + j];
                if (p.aln_score > 0) {
                    int xs = p.sj_type == -1 ? 0 : (p.sj_type == 0 || p.sj_type == 2) ? 2 : 1;
                    int i = p.paired_idx;
                    const dartgpu_report *m = (i != -1 && res.reports[r1.report_off + i].aln_score > 0) ? &res.reports[r1.report_off + i] : nullptr;
                    int dist = 0;
                    if (m) dist = 0 - (int)(p.pos - m->pos + (m->dir ? L2 : 0 - L1));
                    // mate 2 is stored flipped: the stored string is what a reverse-strand report prints
                    emit_mapped(g2, r2, p, p.dir == 0, m, dist, xs);
                }
                if (!P.multi_hit) break;
            }
        }
    }
    for (int u = 0; u < units; u++) { o.sam += chunk[u]; o.n_unmapped += un[u]; o.n_unique += uq[u]; o.n_paired += pr[u]; }
    o.sj.insert(o.sj.end(), res.junctions, res.junctions + res.n_junctions);
}


static void write_header(FILE *fo, const std::vector<std::string> &names, const std::vector<int64_t> &lens)
{
    fprintf(fo, "@PG\tID:Dart\tPN:Dart\tVN:1.4.6\n");   // Mapping.cpp:741 (VersionStr, main.cpp:13)
    for (size_t i = 0; i < names.size(); i++) fprintf(fo, "@SQ\tSN:%s\tLN:%lld\n", names[i].c_str(), (long long)lens[i]);
}

// OutputSpliceJunctions + AbsLoc2ChrLoc (Mapping.cpp:683-716); sj: key -> (type, count)
static int write_junctions(const char *sj_fn, const std::vector<std::string> &names, const std::vector<int64_t> &lens, int64_t G,
                           const std::map<std::pair<int64_t, int64_t>, std::pair<int, int>> &sj)
{
    FILE *fj = fopen(sj_fn, "w");
    int nj = 0;
    if (!fj) return 0;
    std::vector<int64_t> fwd(names.size() + 1, 0);
    for (size_t i = 0; i < names.size(); i++) fwd[i + 1] = fwd[i] + lens[i];
    for (auto &kv : sj) {
        int64_t g1 = kv.first.first, g2 = kv.first.second;
        int chr = -1;   // ChrLocMap.lower_bound(g1): forward ends, then reverse-strand ends
        if (g1 < G) { for (int i = 0; i < (int)fwd.size() - 1; i++) if (g1 <= fwd[i + 1] - 1) { chr = i; break; } }
        else if (g1 < 2 * G) { for (int i = (int)fwd.size() - 2; i >= 0; i--) if (g1 <= 2 * G - fwd[i] - 1) { chr = i; break; } }
        if (chr == -1) continue;
        nj++;
        fprintf(fj, "%s\t%lld\t%lld\t%d\n", names[chr].c_str(), (long long)(g1 + 1 - fwd[chr]), (long long)(g2 + 1 - fwd[chr]), kv.second.second);
    }
    fclose(fj);
    return nj;
}

// header + records + junction table, exactly as Mapping() / OutputSpliceJunctions() write them
static int write_outputs(const char *out_fn, const char *sj_fn, const std::vector<std::string> &names, const std::vector<int64_t> &lens,
                         int64_t G, const std::vector<Out> &outs, int64_t *unm, int64_t *uq, int64_t *prd)
{
    FILE *fo = fopen(out_fn, "w");
    if (!fo) { fprintf(stderr, "cannot write %s\n", out_fn); return -1; }
    write_header(fo, names, lens);
    std::map<std::pair<int64_t, int64_t>, std::pair<int, int>> sj; // key -> (type, count); first insert fixes the type
    for (auto &o : outs) {
        fwrite(o.sam.data(), 1, o.sam.size(), fo);
        *unm += o.n_unmapped; *uq += o.n_unique; *prd += o.n_paired;
        for (auto &j : o.sj) {
            auto it = sj.find({j.g1, j.g2});
            if (it != sj.end()) it->second.second++; else sj[{j.g1, j.g2}] = {j.type, 1};
        }
    }
    fclose(fo);
    return write_junctions(sj_fn, names, lens, G, sj);
}
