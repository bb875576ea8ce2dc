// dart_b200_map — FASTQ in, SAM + junctions.tab out, over libdartgpu.so through the public header only.
//
// FASTQ input takes the streaming path: a reader thread preads blocks of the file(s) into
// page-locked buffers and cuts them at record boundaries (newline counting — the only per-byte work left on the host); one
// worker thread per GPU keeps several blocks in flight on as many contexts (dartgpu_submit_fastq / dartgpu_wait_sam: parse,
// mate-2 flip, encoding, mapping and the SAM text all happen on the device) and writes the returned text in input order.
// What it replaces: the reader (/root/reference/src/GetData.cpp:77-179), the per-read loop (/root/reference/src/Mapping.cpp:598-640),
// the SAM text (/root/reference/src/Mapping.cpp:208-369, :741-751) and the junction table (/root/reference/src/Mapping.cpp:567-577,
// :683-716).  FASTA input (and -hostpath) uses the stand-in host reader / formatter of sam_io.h over dartgpu_map_reads.
// The reference's own binary over the same library is oracle/_ref/dart_gpu (integration/dart_gpu.patch).
//
//   dart_b200_map -i <index prefix> -f r1.fq [-f2 r2.fq] -o out.sam [-j junctions.tab] [-mis N] [-max_dup N]
//                 [-m] [-p] [-unique] [-all_sj] [-max_intron N] [-min_intron N] [-t host threads]
//                 [-batch reads per GPU call] [-devices 0,1,..] [-inflight contexts per GPU] [-writers threads] [-stats] [-hostpath]
#include <fcntl.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "sam_io.h"

namespace {

struct Block {                 // one batch of records: raw text of file 1 (and file 2), page-locked
    char *b1 = nullptr, *b2 = nullptr;
    int64_t cap = 0, len1 = 0, len2 = 0;
    int n_rec = 0;
    int64_t id = -1;
};

struct Stream {                // a file read in blocks that end at record boundaries
    int fd = -1;
    std::string carry;         // bytes behind the last cut
    bool eof = false;
    // fills buf with carry + file bytes; returns the bytes in buf
    int64_t fill(char *buf, int64_t cap)
    {
        int64_t n = (int64_t)carry.size();
        if (n > cap) { fprintf(stderr, "a record is larger than the block buffer\n"); exit(1); }
        memcpy(buf, carry.data(), (size_t)n);
        carry.clear();
        while (!eof && n < cap) {
            ssize_t got = read(fd, buf + n, (size_t)std::min<int64_t>(cap - n, 1 << 26));
            if (got <= 0) { eof = true; break; }
            n += got;
        }
        if (eof && n > 0 && n < cap && buf[n - 1] != '\n') buf[n++] = '\n';     // a last line without its newline
        return n;
    }
    void keep(const char *buf, int64_t used, int64_t n) { carry.assign(buf + used, (size_t)(n - used)); }
};

struct Pipeline {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Block *> free_blocks;
    std::vector<std::deque<Block *>> ready;      // per device, in id order
    bool reader_done = false;
    int64_t next_to_write = 0;
    bool failed = false;
};

} // namespace

int main(int argc, char **argv)
{
    dartgpu_params P; dartgpu_default_params(&P);
    const char *index = nullptr, *f1 = nullptr, *f2 = nullptr, *out_fn = "output.sam", *sj_fn = "junctions.tab";
    bool interleaved = false, want_stats = false, hostpath = false;
    int64_t batch = 1 << 19;
    int inflight = 4, writers = 0;
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
        else if (a == "-inflight") inflight = std::max(1, atoi(next()));
        else if (a == "-writers") writers = std::max(1, atoi(next()));
        else if (a == "-stats") want_stats = true;
        else if (a == "-hostpath") hostpath = true;
        else if (a == "-silent") {}
        else if (a == "-devices") { devices.clear(); char *tok = strtok(next(), ","); while (tok) { devices.push_back(atoi(tok)); tok = strtok(nullptr, ","); } }
        else { fprintf(stderr, "Error! Unknow parameter: %s\n", argv[i]); return 1; }
    }
    if (!index || !f1) { fprintf(stderr, "usage: %s -i prefix -f r1.fq [-f2 r2.fq] -o out.sam\n", argv[0]); return 1; }
    if (P.host_threads > 0) omp_set_num_threads(P.host_threads);
    const bool paired = f2 != nullptr || interleaved;
    P.pair_end = paired ? 1 : 0;
    if (paired) batch &= ~(int64_t)1;
    if (batch < 2) batch = 2;
    const int nd = (int)devices.size();

    // FastQFormat = first byte of the file is '@' (CheckReadFormat, Mapping.cpp:717-725)
    bool fastq = false;
    { FILE *fp = fopen(f1, "rb"); if (!fp) { fprintf(stderr, "Cannot access file:[%s]\n", f1); return 1; } fastq = fgetc(fp) == '@'; fclose(fp); }
    const bool streaming = fastq && !hostpath;

    // contexts: `inflight` per device on the streaming path (they share the device's index), one otherwise
    const int per_dev = streaming ? inflight : 1;
    std::vector<std::vector<dartgpu_ctx *>> ctx(nd);
    for (int d = 0; d < nd; d++)
        for (int k = 0; k < per_dev; k++) {
            dartgpu_ctx *c = nullptr;
            int rc = dartgpu_create_from_files(&c, devices[d], index, &P);
            if (rc != DARTGPU_OK) { fprintf(stderr, "dartgpu_create_from_files failed (%d): %s\n", rc, dartgpu_last_error(nullptr)); return 2; }
            ctx[d].push_back(c);
        }
    dartgpu_ctx *c0 = ctx[0][0];
    std::vector<std::string> names; std::vector<int64_t> lens;
    for (int i = 0; i < dartgpu_num_sequences(c0); i++) { names.push_back(dartgpu_sequence_name(c0, i)); lens.push_back(dartgpu_sequence_length(c0, i)); }
    const int64_t G = dartgpu_genome_size(c0);

    auto t0 = std::chrono::steady_clock::now();
    int64_t n_reads_total = 0, unm = 0, uq = 0, prd = 0;
    int nj = 0;

    if (!streaming) {
        // ---- host reader + host formatter over dartgpu_map_reads (FASTA input, or -hostpath) ----
        Reads R; bool pe = false;
        if (!load_reads(f1, f2, interleaved, R, pe)) return 1;
        const int64_t n = (int64_t)R.name.size();
        std::vector<Out> outs(nd);
        std::vector<int64_t> lo(nd + 1, 0);
        for (int d = 0; d <= nd; d++) { lo[d] = n * d / nd; if (paired) lo[d] &= ~(int64_t)1; }
        lo[nd] = n;
        std::vector<int> rcs(nd, 0);
        auto worker = [&](int d) {
            for (int64_t b = lo[d]; b < lo[d + 1]; b += batch) {
                int64_t e = std::min(lo[d + 1], b + batch);
                dartgpu_reads rd; rd.n_reads = (int32_t)(e - b); rd.bases = R.seq.data(); rd.offsets = R.off.data() + b;
                dartgpu_map_result res;
                int rc = dartgpu_map_reads(ctx[d][0], &rd, &res);
                if (rc != DARTGPU_OK) { fprintf(stderr, "dartgpu_map_reads failed (%d): %s\n", rc, dartgpu_last_error(ctx[d][0])); rcs[d] = rc; return; }
                format_batch(names, R, b, res, P, outs[d]);
            }
        };
        if (nd == 1) worker(0);
        else { std::vector<std::thread> th; for (int d = 0; d < nd; d++) th.emplace_back(worker, d); for (auto &t : th) t.join(); }
        for (int d = 0; d < nd; d++) if (rcs[d]) return 3;
        nj = write_outputs(out_fn, sj_fn, names, lens, G, outs, &unm, &uq, &prd);
        if (nj < 0) return 1;
        n_reads_total = n;
    } else {
        // ---- streaming path: raw FASTQ blocks -> GPU -> SAM text ----
        const int out_fd = open(out_fn, O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (out_fd < 0) { fprintf(stderr, "cannot write %s\n", out_fn); return 1; }
        int64_t file_off = 0;
        {
            std::string h = "@PG\tID:Dart\tPN:Dart\tVN:1.4.6\n";   // Mapping.cpp:741 (VersionStr, main.cpp:13)
            for (size_t i = 0; i < names.size(); i++) h += "@SQ\tSN:" + names[i] + "\tLN:" + std::to_string(lens[i]) + "\n";
            if (write(out_fd, h.data(), h.size()) != (ssize_t)h.size()) { fprintf(stderr, "cannot write %s\n", out_fn); return 1; }
            file_off = (int64_t)h.size();
        }
        const int rec_per_block = (int)(paired && !f2 ? batch : (f2 ? batch / 2 : batch));   // records of file 1 per block
        Stream s1, s2;
        s1.fd = open(f1, O_RDONLY);
        if (f2) { s2.fd = open(f2, O_RDONLY); if (s2.fd < 0) { fprintf(stderr, "Cannot access file:[%s]\n", f2); return 1; } }
        int64_t first_rec = 512;
        { char probe[8192]; ssize_t got = pread(s1.fd, probe, sizeof probe, 0); int32_t k = 0; int64_t used = dartgpu_fastq_cut(probe, got > 0 ? got : 0, 1, &k); if (k == 1) first_rec = used; }
        // a block buffer holds the wanted records with a little room; a block also ends when its buffer is full
        const int64_t cap = std::max<int64_t>(1 << 20, (int64_t)((double)rec_per_block * (double)first_rec * 1.06) + (1 << 16));
        const int n_blocks = nd * inflight + 1;
        std::vector<Block> blocks(n_blocks);
        Pipeline PL; PL.ready.resize(nd);
        for (auto &b : blocks) { b.cap = cap; PL.free_blocks.push_back(&b); }   // page-locked on first use, by the reader: pinning
                                                                                  // 100+ MB takes ~50 ms and now overlaps the GPU
        // ---- reader: the two files of a block are read and cut by two threads at once ----
        std::thread reader([&] {
            int64_t id = 0;
            for (;;) {
                Block *b;
                { std::unique_lock<std::mutex> lk(PL.mu); PL.cv.wait(lk, [&] { return !PL.free_blocks.empty() || PL.failed; }); if (PL.failed) break; b = PL.free_blocks.front(); PL.free_blocks.pop_front(); }
                if (!b->b1) {
                    b->b1 = (char *)dartgpu_alloc_pinned(cap); b->b2 = f2 ? (char *)dartgpu_alloc_pinned(cap) : nullptr;
                    if (!b->b1 || (f2 && !b->b2)) { fprintf(stderr, "cannot allocate page-locked block buffers\n"); std::lock_guard<std::mutex> lk(PL.mu); PL.failed = true; PL.reader_done = true; PL.cv.notify_all(); return; }
                }
                int64_t n1 = 0, n2 = 0, used1 = 0, used2 = 0;
                int32_t k1 = 0, k2 = 0;
                std::thread second;
                if (f2) second = std::thread([&] { n2 = s2.fill(b->b2, b->cap - 1); used2 = dartgpu_fastq_cut(b->b2, n2, rec_per_block, &k2); });
                n1 = s1.fill(b->b1, b->cap - 1);
                used1 = dartgpu_fastq_cut(b->b1, n1, rec_per_block, &k1);
                if (f2) {
                    second.join();
                    if (k2 < k1) used1 = dartgpu_fastq_cut(b->b1, n1, k2, &k1);      // the shorter side decides (GetData.cpp:152-155)
                    else if (k1 < k2) used2 = dartgpu_fastq_cut(b->b2, n2, k1, &k2);
                    s2.keep(b->b2, used2, n2);
                } else if (paired && (k1 & 1)) used1 = dartgpu_fastq_cut(b->b1, n1, k1 - 1, &k1);   // -p: mates stay together
                s1.keep(b->b1, used1, n1);
                if (k1 == 0) { std::lock_guard<std::mutex> lk(PL.mu); PL.free_blocks.push_back(b); break; }
                b->len1 = used1; b->len2 = used2; b->n_rec = k1; b->id = id;
                { std::lock_guard<std::mutex> lk(PL.mu); PL.ready[id % nd].push_back(b); }
                PL.cv.notify_all();
                id++;
            }
            { std::lock_guard<std::mutex> lk(PL.mu); PL.reader_done = true; }
            PL.cv.notify_all();
        });
        // ---- writers: a block's text goes to its place in the file (pwrite) while later blocks are already being mapped ----
        struct WriteJob { const char *p; int64_t n, off; int d; dartgpu_ctx *c; Block *b; };
        std::deque<WriteJob> wq;
        std::vector<std::deque<dartgpu_ctx *>> free_ctx(nd);
        for (int d = 0; d < nd; d++) for (auto c : ctx[d]) free_ctx[d].push_back(c);
        int writes_pending = 0;
        bool workers_done = false;
        auto writer = [&] {
            for (;;) {
                WriteJob j;
                {
                    std::unique_lock<std::mutex> lk(PL.mu);
                    PL.cv.wait(lk, [&] { return !wq.empty() || workers_done || PL.failed; });
                    if (wq.empty()) return;
                    j = wq.front(); wq.pop_front();
                }
                int64_t done = 0;
                bool ok = true;
                while (done < j.n) {
                    ssize_t w = pwrite(out_fd, j.p + done, (size_t)std::min<int64_t>(j.n - done, 1 << 27), j.off + done);
                    if (w <= 0) { ok = false; break; }
                    done += w;
                }
                {
                    std::lock_guard<std::mutex> lk(PL.mu);
                    if (!ok) { fprintf(stderr, "write to %s failed\n", out_fn); PL.failed = true; }
                    free_ctx[j.d].push_back(j.c); PL.free_blocks.push_back(j.b); writes_pending--;
                }
                PL.cv.notify_all();
            }
        };
        std::vector<std::map<std::pair<int64_t, int64_t>, std::pair<int, int>>> sj(nd);
        std::vector<int64_t> w_reads(nd, 0), w_unm(nd, 0), w_uq(nd, 0), w_prd(nd, 0);
        auto worker = [&](int d) {
            dartgpu_bind_host_thread(devices[d]);       // stay on the GPU's NUMA node (best effort)
            std::deque<std::pair<dartgpu_ctx *, Block *>> flying;
            auto fail_all = [&](const char *what, dartgpu_ctx *c, int rc) {
                fprintf(stderr, "%s failed (%d): %s\n", what, rc, dartgpu_last_error(c));
                { std::lock_guard<std::mutex> lk(PL.mu); PL.failed = true; }
                PL.cv.notify_all();
            };
            for (;;) {
                // submit while a context is free and a block is ready; sleep only when nothing is in flight
                for (;;) {
                    Block *b = nullptr; dartgpu_ctx *c = nullptr;
                    {
                        std::unique_lock<std::mutex> lk(PL.mu);
                        if (flying.empty())
                            PL.cv.wait(lk, [&] { return (!PL.ready[d].empty() && !free_ctx[d].empty()) || (PL.reader_done && PL.ready[d].empty()) || PL.failed; });
                        if (PL.failed) return;
                        if (!PL.ready[d].empty() && !free_ctx[d].empty()) {
                            b = PL.ready[d].front(); PL.ready[d].pop_front();
                            c = free_ctx[d].front(); free_ctx[d].pop_front();
                        }
                    }
                    if (!b) break;
                    dartgpu_fastq_block fb{b->b1, b->len1, f2 ? b->b2 : nullptr, b->len2, b->n_rec, 1, 0, 0};
                    int rc = dartgpu_submit_fastq(c, &fb);
                    if (rc != DARTGPU_OK) { fail_all("dartgpu_submit_fastq", c, rc); return; }
                    flying.push_back({c, b});
                }
                if (flying.empty()) return;          // reader done, nothing ready, nothing in flight
                dartgpu_ctx *c = flying.front().first; Block *b = flying.front().second;
                flying.pop_front();
                dartgpu_sam_result res;
                int rc = dartgpu_wait_sam(c, &res);
                if (rc != DARTGPU_OK) { fail_all("dartgpu_wait_sam", c, rc); return; }
                if (want_stats) {
                    dartgpu_stats s; dartgpu_get_stats(c, &s);
                    fprintf(stderr, "[gpu %d] block %lld reads %lld  h2d %.2f ms search %.2f locate %.2f sort+cluster %.2f kmer %.2f (%llu jobs) nw %.2f (%llu jobs, %llu cells) report+sam %.2f d2h %.2f  host %.2f ms  launches %llu  sam %.1f MB\n",
                            devices[d], (long long)b->id, (long long)res.n_reads, s.ms_h2d, s.ms_search, s.ms_locate, s.ms_sort_cluster, s.ms_kmer,
                            (unsigned long long)s.kmer_jobs, s.ms_nw, (unsigned long long)s.nw_jobs, (unsigned long long)s.nw_cells, s.ms_report, s.ms_d2h,
                            s.ms_host, (unsigned long long)s.kernel_launches, res.n_bytes / 1e6);
                }
                for (int64_t k = 0; k < res.n_junctions; k++) {
                    const dartgpu_junction &j = res.junctions[k];
                    auto it = sj[d].find({j.g1, j.g2});
                    if (it != sj[d].end()) it->second.second++; else sj[d][{j.g1, j.g2}] = {j.type, 1};
                }
                w_reads[d] += res.n_reads; w_unm[d] += res.n_unmapped; w_uq[d] += res.n_unique; w_prd[d] += res.n_paired;
                {   // blocks take their place in the file in input order, whichever GPU finishes first
                    std::unique_lock<std::mutex> lk(PL.mu);
                    PL.cv.wait(lk, [&] { return PL.next_to_write == b->id || PL.failed; });
                    if (PL.failed) return;
                    wq.push_back(WriteJob{res.sam, res.n_bytes, file_off, d, c, b});
                    file_off += res.n_bytes; writes_pending++;
                    PL.next_to_write = b->id + 1;
                }
                PL.cv.notify_all();
            }
        };
        std::vector<std::thread> th, wr;
        const int n_writers = writers > 0 ? writers : std::max(4, std::min(16, nd * 3));
        for (int i = 0; i < n_writers; i++) wr.emplace_back(writer);
        for (int d = 0; d < nd; d++) th.emplace_back(worker, d);
        for (auto &t : th) t.join();
        { std::unique_lock<std::mutex> lk(PL.mu); PL.cv.wait(lk, [&] { return writes_pending == 0 || PL.failed; }); workers_done = true; }
        PL.cv.notify_all();
        for (auto &t : wr) t.join();
        PL.cv.notify_all();
        reader.join();
        close(out_fd);
        if (PL.failed) return 3;
        // junction counts summed by key over the GPUs (UpdateGlobalSJMap, Mapping.cpp:567-577); the type of a (g1,g2) pair is a
        // function of the genome (the splice motif), identical wherever it is seen
        std::map<std::pair<int64_t, int64_t>, std::pair<int, int>> all;
        for (int d = 0; d < nd; d++)
            for (auto &kv : sj[d]) { auto it = all.find(kv.first); if (it != all.end()) it->second.second += kv.second.second; else all[kv.first] = kv.second; }
        nj = write_junctions(sj_fn, names, lens, G, all);
        for (int d = 0; d < nd; d++) { n_reads_total += w_reads[d]; unm += w_unm[d]; uq += w_uq[d]; prd += w_prd[d]; }
        for (auto &b : blocks) { dartgpu_free_pinned(b.b1); dartgpu_free_pinned(b.b2); }
        close(s1.fd); if (f2) close(s2.fd);
    }
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stdout, "All the %lld %s reads have been processed in %.3f seconds (%.0f reads/s on %d GPU%s, %s).\n", (long long)n_reads_total,
            paired ? "paired-end" : "single-end", sec, n_reads_total / std::max(sec, 1e-9), nd, nd > 1 ? "s" : "",
            streaming ? "FASTQ parse, mapping and SAM text on the device" : "host reader and formatter");
    fprintf(stdout, "\t# of total mapped reads = %lld\n\t# of unique mapped reads = %lld\n\t# of unmapped reads = %lld\n\t# of paired sequences = %lld\n\t# of splice junctions = %d (file: %s)\n",
            (long long)(n_reads_total - unm), (long long)uq, (long long)unm, (long long)prd, nj, sj_fn);
    for (auto &v : ctx) for (auto c : v) dartgpu_destroy(c);
    return 0;
}
