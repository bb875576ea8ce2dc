// TEST INFRASTRUCTURE — never part of the product.
//
// Runs the device orchestration logic (dart_b200/csrc/report_logic.cuh, the code the report kernels execute one
// thread per candidate) on the CPU, with the oracle standing in for the CUDA kernels (seeds, candidates, 8-mer
// re-seeding, NW), and writes SAM + junctions through the same writer dart_b200_map uses.  tests/test_logic_cpu.py
// compares its output with the canonical reference, so the orchestration is proven on a box without a GPU.
//
//   logic_harness -i idx -f r1.fq [-f2 r2.fq] -o out.sam -j junc [-mis N] [-max_dup N] [-m] [-all_sj] ...
#include <cstdint>
#include <numeric>

#include "../../dart_b200/csrc/report_logic.cuh"
#include "../../dart_b200/csrc/sam_io.h"
#include "../../oracle/dart_oracle.h"

using namespace dartgpu;

static uint8_t code_of(unsigned char ch)
{   // must match dart_b200/csrc/capi.cu
    switch (ch) {
    case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3;
    case 'a': return 8; case 'c': return 9; case 'g': return 10; case 't': return 11;
    case 'N': return 5; default: return 4;
    }
}

int main(int argc, char **argv)
{
    dartgpu_params P{5, 500000, 5, 0, 100, 0, 0, 0, 0, 0};   // the reference defaults (src/main.cpp:101-117); no libdartgpu here
    const char *index = nullptr, *f1 = nullptr, *f2 = nullptr, *out_fn = "out.sam", *sj_fn = "junc.tab";
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { return argv[++i]; };
        if (a == "-i") index = next(); else if (a == "-f") f1 = next(); else if (a == "-f2") f2 = next();
        else if (a == "-o") out_fn = next(); else if (a == "-j") sj_fn = next();
        else if (a == "-mis") P.max_mismatch = atoi(next()); else if (a == "-max_dup") P.max_dup = (uint32_t)atoi(next());
        else if (a == "-m") P.multi_hit = 1; else if (a == "-unique") P.unique = 1; else if (a == "-all_sj") P.all_sj = 1;
        else if (a == "-max_intron") P.max_intron = atoi(next()); else if (a == "-min_intron") P.min_intron = atoi(next());
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 1; }
    }
    if (P.max_dup < 100) P.max_dup = 100; else if (P.max_dup >= 10000) P.max_dup = 10000;
    if (P.max_intron < 100000) P.max_intron = 100000;
    Reads R; bool paired = false;
    if (!load_reads(f1, f2, false, R, paired)) return 1;
    P.pair_end = paired;
    const int n = (int)R.name.size();
    or_index *O = or_load(index);
    if (!O) { fprintf(stderr, "cannot load %s\n", index); return 1; }
    // index side tables
    std::string pre(index), pac;
    slurp((pre + ".pac").c_str(), pac);
    std::vector<std::string> names; std::vector<int64_t> lens, fwd, ends; std::vector<int32_t> end_chr;
    {
        FILE *fp = fopen((pre + ".ann").c_str(), "r");
        long long lpac; int nseq; unsigned seed;
        if (fscanf(fp, "%lld%d%u", &lpac, &nseq, &seed) != 3) return 1;
        for (int i = 0; i < nseq; i++) {
            unsigned gi; char name[1024]; long long off; int len, namb;
            if (fscanf(fp, "%u%1023s", &gi, name) != 2) break;
            int ch; while ((ch = fgetc(fp)) != '\n' && ch != EOF) {}
            if (fscanf(fp, "%lld%d%d", &off, &len, &namb) != 3) break;
            names.push_back(name); lens.push_back(len);
        }
        fclose(fp);
    }
    const int64_t G = or_genome_size(O);
    {
        int64_t acc = 0; std::vector<std::pair<int64_t, int>> e;
        for (size_t i = 0; i < lens.size(); i++) { fwd.push_back(acc); acc += lens[i]; e.push_back({fwd[i] + lens[i] - 1, (int)i}); e.push_back({2 * G - acc + lens[i] - 1, (int)i}); }
        std::sort(e.begin(), e.end());
        for (auto &p : e) { ends.push_back(p.first); end_chr.push_back(p.second); }
    }

    // ---- reads -> codes (16-byte aligned slots, as the device batch) ----
    std::vector<int64_t> code_off(n + 1, 0); std::vector<int32_t> rlen(n);
    for (int i = 0; i < n; i++) { rlen[i] = (int)(R.off[i + 1] - R.off[i]); code_off[i + 1] = code_off[i] + ((rlen[i] + 15) & ~15); }
    std::vector<uint8_t> codes(code_off[n] + 16, 4), plain(code_off[n] + 16, 4);
    std::vector<char> chars(code_off[n] + 16, 'N');
    for (int i = 0; i < n; i++)
        for (int k = 0; k < rlen[i]; k++) {
            char ch = R.seq[R.off[i] + k];
            uint8_t c = code_of((unsigned char)ch);
            codes[code_off[i] + k] = c; plain[code_off[i] + k] = (c & 4) ? 4 : (c & 3); chars[code_off[i] + k] = ch;
        }

    // ---- stage 1 via the oracle: seeds (as keys) and candidates ----
    std::vector<uint64_t> keys; std::vector<int64_t> seed_off(n + 1, 0);
    std::vector<CandState> cs; std::vector<int64_t> cand_off(n + 1, 0);
    {
        const int cap = 1 << 20;
        std::vector<int32_t> rp(cap), ln(cap), cb(cap), cc(cap), sc(cap); std::vector<int64_t> gp(cap);
        for (int i = 0; i < n; i++) {
            int ns = or_seed_read(O, plain.data() + code_off[i], rlen[i], (int)P.max_dup, rp.data(), gp.data(), ln.data(), cap);
            if (ns > cap) { fprintf(stderr, "seed capacity\n"); return 1; }
            for (int s = 0; s < ns; s++) keys.push_back(seed_key((uint64_t)gp[s], (uint32_t)rp[s], (uint32_t)ln[s]));
            int nc = or_cluster_read(O, rlen[i], ns, rp.data(), gp.data(), ln.data(), P.max_gaps, P.max_intron, cb.data(), cc.data(), sc.data(), cap);
            for (int k = 0; k < nc; k++) {
                CandState c{};
                c.read = i; c.seed_begin = cb[k]; c.seed_count = cc[k]; c.Score = sc[k]; c.PairedIdx = -1; c.SJtype = -1;
                int64_t pd = gp[cb[k]] - rp[cb[k]]; c.PosDiff = pd < 0 ? 0 : pd;
                cs.push_back(c);
            }
            seed_off[i + 1] = seed_off[i] + ns; cand_off[i + 1] = cand_off[i] + nc;
        }
    }
    // ---- pairing / pruning ----
    for (int u = 0; u < (paired ? n / 2 : n); u++) {
        if (paired) pair_and_prune(cs.data() + cand_off[2 * u], (int)(cand_off[2 * u + 1] - cand_off[2 * u]), cs.data() + cand_off[2 * u + 1],
                                   (int)(cand_off[2 * u + 2] - cand_off[2 * u + 1]), true);
        else pair_and_prune(cs.data() + cand_off[u], (int)(cand_off[u + 1] - cand_off[u]), nullptr, 0, false);
    }
    const int ncand = (int)cs.size();
    int64_t pool_total = 0;
    for (auto &c : cs) { c.live = c.Score != 0; c.sv_cap = c.live ? seed_capacity(c.seed_count) : 0; c.sv_off = pool_total; pool_total += c.sv_cap; }
    std::vector<RSeed> pool(pool_total + 1);
    std::vector<KmerJobDev> kjobs(pool_total / 12 + 2); std::vector<dartgpu_kmer_hit> khits(kjobs.size());
    std::vector<NwJobDev> jobsB(pool_total / 3 + 4), jobsC(pool_total + 4);
    int32_t nk = 0, nB = 0, nC = 0;

    Env E{};
    E.P = PhaseParams{P.max_gaps, P.max_intron, P.min_intron, P.max_mismatch, P.multi_hit, P.pair_end, P.all_sj};
    E.ref = RefView{nullptr, (const uint8_t *)pac.data(), G}; E.G = G;
    E.ends = ends.data(); E.end_chr = end_chr.data(); E.n_ends = (int)ends.size(); E.chr_fwd = fwd.data();
    E.codes = codes.data(); E.code_off = code_off.data(); E.rlen = rlen.data(); E.keys = keys.data(); E.seed_off = seed_off.data();
    E.cs = cs.data(); E.pool = pool.data();
    E.kjobs = kjobs.data(); E.kjob_count = &nk; E.khits = khits.data();

    auto genome = [&](int64_t p, int len) { std::vector<uint8_t> c(len > 0 ? len : 1); or_ref_codes(O, p, len, c.data()); std::string s(len, 'A'); for (int i = 0; i < len; i++) s[i] = "ACGT"[c[i]]; return s; };
    auto run_nw = [&](std::vector<NwJobDev> &jobs, int nj, std::vector<uint8_t> &ops, std::vector<int32_t> &nops, std::vector<int32_t> *aux) {
        int64_t o = 0, a = 0;
        for (int j = 0; j < nj; j++) { jobs[j].op_off = o; o += jobs[j].m + jobs[j].n; jobs[j].aux_off = a; if (aux && !(j & 1)) a += 2 * (jobs[j].m + 1); }
        ops.assign(o + 1, 0); nops.assign(nj + 1, 0);
        if (aux) aux->assign(a + 1, 0);
        std::vector<uint8_t> tmp;
        for (int j = 0; j < nj; j++) {
            std::string g = genome(jobs[j].gpos, jobs[j].n);
            tmp.assign(jobs[j].m + jobs[j].n + 2, 0);
            int k = or_nw(O, jobs[j].m, chars.data() + jobs[j].s1_off, jobs[j].n, g.data(), tmp.data());
            nops[j] = k;
            memcpy(ops.data() + jobs[j].op_off + jobs[j].m + jobs[j].n - k, tmp.data(), k);
        }
    };

    for (int c = 0; c < ncand; c++) phase_a(E, cs[c], pool.data() + cs[c].sv_off);
    for (int j = 0; j < nk; j++) {
        std::string g = genome(kjobs[j].gpos, kjobs[j].len2);
        int64_t o3[3];
        or_kmer_pair(O, kjobs[j].len1, chars.data() + kjobs[j].s1_off, kjobs[j].len2, g.data(), o3);
        khits[j].rpos = (int32_t)o3[0]; khits[j].gpos = (int32_t)o3[1]; khits[j].len = (int32_t)o3[2];
    }
    E.njobs = jobsB.data(); E.njob_count = &nB;
    for (int c = 0; c < ncand; c++) phase_b(E, cs[c], pool.data() + cs[c].sv_off);
    std::vector<uint8_t> opsB, opsC; std::vector<int32_t> nopsB, nopsC, aux;
    run_nw(jobsB, nB, opsB, nopsB, &aux);
    E.ops = opsB.data(); E.nops = nopsB.data(); E.done_jobs = jobsB.data(); E.xscratch = aux.data();
    E.njobs = jobsC.data(); E.njob_count = &nC;
    for (int c = 0; c < ncand; c++) phase_c(E, cs[c], pool.data() + cs[c].sv_off);
    run_nw(jobsC, nC, opsC, nopsC, nullptr);
    int64_t cig_total = 0;
    for (auto &c : cs) { c.cig_off = cig_total; cig_total += c.live && !c.skip ? c.cig_cap : 0; }
    std::vector<int32_t> cig(cig_total + 1);
    E.ops = opsC.data(); E.nops = nopsC.data(); E.done_jobs = jobsC.data(); E.cig = cig.data();
    for (int c = 0; c < ncand; c++) { phase_d(E, cs[c], pool.data() + cs[c].sv_off); if (cs[c].cig_n < 0) { fprintf(stderr, "CIGAR capacity overflow\n"); return 1; } }

    // ---- final pass ----
    std::vector<dartgpu_read_result> rr(n); std::vector<ReadOut> ro(n);
    int64_t nrep = 0;
    for (int i = 0; i < n; i++) { rr[i].report_off = nrep; int nc = (int)(cand_off[i + 1] - cand_off[i]); nrep += nc > 0 ? nc : 1; }
    std::vector<dartgpu_report> rep(nrep);
    for (int i = 0; i < n; i++) read_best(cs.data() + cand_off[i], (int)(cand_off[i + 1] - cand_off[i]), rep.data() + rr[i].report_off, ro[i]);
    for (int u = 0; u < (paired ? n / 2 : n); u++) {
        if (!paired) finish_single(ro[u], rep.data() + rr[u].report_off, cs.data() + cand_off[u], (int)(cand_off[u + 1] - cand_off[u]));
        else {
            int a = 2 * u, b = a + 1;
            finish_pair(ro[a], rep.data() + rr[a].report_off, cs.data() + cand_off[a], (int)(cand_off[a + 1] - cand_off[a]),
                        ro[b], rep.data() + rr[b].report_off, cs.data() + cand_off[b], (int)(cand_off[b + 1] - cand_off[b]), P.multi_hit != 0);
        }
    }
    std::vector<char> text; std::vector<dartgpu_junction> junc;
    for (int i = 0; i < n; i++) {
        rr[i].mapq = ro[i].mapq; rr[i].score = ro[i].score; rr[i].sub_score = ro[i].sub_score; rr[i].mis_num = ro[i].mis_num;
        rr[i].n_reports = ro[i].n_reports; rr[i].best = ro[i].best;
        int nc = (int)(cand_off[i + 1] - cand_off[i]);
        for (int k = 0; k < nc; k++) {
            const CandState &c = cs[cand_off[i] + k];
            dartgpu_report &p = rep[rr[i].report_off + k];
            p.cigar_off = (int32_t)text.size();
            if (c.live && !c.skip && c.AlnScore > 0) {
                size_t at = text.size(); text.resize(at + c.text_len);
                write_cigar_text(cig.data() + c.cig_off, c.cig_n, text.data() + at);
                p.cigar_len = c.text_len;
            } else p.cigar_len = 0;
        }
        int nj = emit_junctions(E, ro[i], cs.data() + cand_off[i], nc, i, nullptr);
        size_t at = junc.size(); junc.resize(at + nj);
        emit_junctions(E, ro[i], cs.data() + cand_off[i], nc, i, junc.data() + at);
    }
    dartgpu_map_result res{rr.data(), n, rep.data(), nrep, text.data(), (int64_t)text.size(), junc.data(), (int64_t)junc.size()};
    std::vector<Out> outs(1);
    format_batch(names, R, 0, res, P, outs[0]);
    int64_t unm = 0, uq = 0, prd = 0;
    int nj = write_outputs(out_fn, sj_fn, names, lens, G, outs, &unm, &uq, &prd);
    fprintf(stdout, "logic_harness: %d reads, %d candidates, %d k-mer jobs, %d+%d NW jobs, %d junction rows\n", n, ncand, nk, nB, nC, nj);
    or_free(O);
    return 0;
}
