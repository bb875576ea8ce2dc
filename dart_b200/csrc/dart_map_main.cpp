// dart_b200_map — a stand-in for the parts of the reference that "stay in place" around the GPU path, so that
// the C-ABI can be exercised end to end without the reference tree: the FASTQ reader
// (/root/reference/src/GetData.cpp:77-179), the SAM text writer (/root/reference/src/Mapping.cpp:208-369, :741-751)
// and the junction table (/root/reference/src/Mapping.cpp:567-577, :683-716).  It links libdartgpu.so through the
// public header only; INTEGRATION.md shows the equivalent patch to the reference's own ReadMapping().
//
//   dart_b200_map -i <index prefix> -f r1.fq [-f2 r2.fq] -o out.sam [-j junctions.tab] [-mis N] [-max_dup N]
//                 [-m] [-p] [-unique] [-all_sj] [-max_intron N] [-min_intron N] [-t host threads]
//                 [-batch reads per GPU call] [-devices 0,1,..] [-stats]
#include <chrono>
#include <thread>

#include "sam_io.h"

int main(int argc, char **argv)
{
    dartgpu_params P; dartgpu_default_params(&P);
    const char *index = nullptr, *f1 = nullptr, *f2 = nullptr, *out_fn = "output.sam", *sj_fn = "junctions.tab";
    bool interleaved = false, want_stats = false;
    int64_t batch = 1 << 20;
    std::vector<int> devices{0};
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(1); } return argv[++i]; };
        if (a == "-i") index = next();
        else if (a == "-f") f1 = next();
        else if (a == "-f2") f2 = next();
        else if (a == "-o") out_fn = next();
        else if (a == "-j") sj_fn = next();
        else if (a == "-mis") P.max_mismatch = atoi(next());
        else if (a == "-max_dup") P.max_dup = (uint32_t)atoi(next());
        else if (a == "-m") P.multi_hit = 1;
        else if (a == "-p") interleaved = true;
        else if (a == "-unique") P.unique = 1;
        else if (a == "-all_sj") P.all_sj = 1;
        else if (a == "-max_intron") P.max_intron = atoi(next());
        else if (a == "-min_intron") P.min_intron = atoi(next());
        else if (a == "-t") P.host_threads = atoi(next());
        else if (a == "-batch") batch = atoll(next());
        else if (a == "-stats") want_stats = true;
        else if (a == "-silent") {}
        else if (a == "-devices") { devices.clear(); char *tok = strtok(next(), ","); while (tok) { devices.push_back(atoi(tok)); tok = strtok(nullptr, ","); } }
        else { fprintf(stderr, "Error! Unknow parameter: %s\n", argv[i]); return 1; }
    }
    if (!index || !f1) { fprintf(stderr, "usage: %s -i prefix -f r1.fq [-f2 r2.fq] -o out.sam\n", argv[0]); return 1; }
    if (P.host_threads > 0) omp_set_num_threads(P.host_threads);

    Reads R; bool paired = false;
    if (!load_reads(f1, f2, interleaved, R, paired)) return 1;
    P.pair_end = paired ? 1 : 0;
    if (paired) batch &= ~(int64_t)1;
    const int64_t n = (int64_t)R.name.size();

    const int nd = (int)devices.size();
    std::vector<dartgpu_ctx *> ctx(nd, nullptr);
    for (int d = 0; d < nd; d++) {
        int rc = dartgpu_create_from_files(&ctx[d], devices[d], index, &P);
        if (rc != DARTGPU_OK) { fprintf(stderr, "dartgpu_create_from_files failed (%d): %s\n", rc, dartgpu_last_error(nullptr)); return 2; }
    }
    std::vector<std::string> names_for_format;
    for (int i = 0; i < dartgpu_num_sequences(ctx[0]); i++) names_for_format.push_back(dartgpu_sequence_name(ctx[0], i));
    // contiguous read range per GPU (pairs kept together); results merged in input order
    std::vector<Out> outs(nd);
    std::vector<int64_t> lo(nd + 1, 0);
    for (int d = 0; d <= nd; d++) { lo[d] = n * d / nd; if (paired) lo[d] &= ~(int64_t)1; }
    lo[nd] = n;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<int> rcs(nd, 0);
    auto worker = [&](int d) {
        for (int64_t b = lo[d]; b < lo[d + 1]; b += batch) {
            int64_t e = std::min(lo[d + 1], b + batch);
            dartgpu_reads rd; rd.n_reads = (int32_t)(e - b); rd.bases = R.seq.data(); rd.offsets = R.off.data() + b;
            dartgpu_map_result res;
            int rc = dartgpu_map_reads(ctx[d], &rd, &res);
            if (rc != DARTGPU_OK) { fprintf(stderr, "dartgpu_map_reads failed (%d): %s\n", rc, dartgpu_last_error(ctx[d])); rcs[d] = rc; return; }
            format_batch(names_for_format, R, b, res, P, outs[d]);
            if (want_stats) {
                dartgpu_stats s; dartgpu_get_stats(ctx[d], &s);
                fprintf(stderr, "[gpu %d] reads %lld  search %.2f ms  locate %.2f  sort+cluster %.2f  kmer %.2f (%llu jobs)  nw %.2f (%llu jobs, %llu cells)  host %.2f ms  launches %llu\n",
                        devices[d], (long long)(e - b), s.ms_search, s.ms_locate, s.ms_sort_cluster, s.ms_kmer, (unsigned long long)s.kmer_jobs,
                        s.ms_nw, (unsigned long long)s.nw_jobs, (unsigned long long)s.nw_cells, s.ms_host, (unsigned long long)s.kernel_launches);
            }
        }
    };
    if (nd == 1) worker(0);
    else { std::vector<std::thread> th; for (int d = 0; d < nd; d++) th.emplace_back(worker, d); for (auto &t : th) t.join(); }
    for (int d = 0; d < nd; d++) if (rcs[d]) return 3;
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    std::vector<std::string> names; std::vector<int64_t> lens;
    for (int i = 0; i < dartgpu_num_sequences(ctx[0]); i++) { names.push_back(dartgpu_sequence_name(ctx[0], i)); lens.push_back(dartgpu_sequence_length(ctx[0], i)); }
    int64_t unm = 0, uq = 0, prd = 0;
    int nj = write_outputs(out_fn, sj_fn, names, lens, dartgpu_genome_size(ctx[0]), outs, &unm, &uq, &prd);
    if (nj < 0) return 1;
    fprintf(stdout, "All the %lld %s reads have been processed in %.3f seconds (%.0f reads/s on %d GPU%s).\n", (long long)n,
            paired ? "paired-end" : "single-end", sec, n / std::max(sec, 1e-9), nd, nd > 1 ? "s" : "");
    fprintf(stdout, "\t# of total mapped reads = %lld\n\t# of unique mapped reads = %lld\n\t# of unmapped reads = %lld\n\t# of paired sequences = %lld\n\t# of splice junctions = %d (file: %s)\n",
            (long long)(n - unm), (long long)uq, (long long)unm, (long long)prd, nj, sj_fn);
    for (auto c : ctx) dartgpu_destroy(c);
    return 0;
}
