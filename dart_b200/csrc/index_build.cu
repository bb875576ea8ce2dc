// FM-index construction on the GPU: <prefix>.bwt and <prefix>.sa, byte-identical to what the reference's builder writes.
//
// Replaces steps 2-5 of bwa_idx_build (/root/reference/src/BWT_Index/bwtindex.c:77-148): bwt_bwtgen2 (BWT-SW, single
// threaded: ~100 s for 186 Mbp on the GPU box's host, hours for 3.1 Gbp), bwt_bwtupdate_core (Occ interleave, :53-75),
// bwt_cal_sa(bwt, 32) (bwt.c:101-123) and the two dumps (bwt.c:174-196).  The BWT of a text is unique, so a different
// construction must produce the same bytes (tests/test_index_build.py compares against the reference's own files).
//
// Method (sized for 180 GB of HBM; a 3.1 Gbp genome = 6.2 G suffixes builds in one pass over a handful of key ranges):
//   text T = forward strand + reverse complement, 2 bits per base (the same ref2 array the mapping kernels read);
//   1. histogram of the leading 6-mers -> split [0, 4^6) into consecutive ranges of at most `limit` suffixes;
//   2. per range: collect (key = the first 32 bases as a u64, position) of every suffix that starts in the range,
//      cub radix sort; suffixes that still tie after 32 bases (repeats, the end of the text) are ordered exactly by a
//      word-wise comparison of the text (on the host: they are rare outside repeat-rich genomes);
//   3. per range: emit BWT symbols T[SA-1] (one byte per row), every 32nd SA entry and the primary row;
//   4. drop the primary row, pack 16 symbols per word, per-128-symbol histograms -> exclusive scan -> the interleaved
//      image the reference's loader expects.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include <omp.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dartgpu_internal.h"

namespace dartgpu {

namespace {

constexpr int PFX = 6;                       // bases of the range-splitting prefix
constexpr int NPFX = 1 << (2 * PFX);

// 32 bases starting at p, first base in the top bits; beyond the text the guard words read as A (the smallest symbol)
__host__ __device__ inline uint64_t bases64(const uint32_t *T, uint64_t p)
{
    const uint64_t w = p >> 4;
    const int sh = 2 * (int)(p & 15);
    const uint64_t hi = (uint64_t)T[w] << 32 | T[w + 1];
    return sh ? (hi << sh) | ((uint64_t)T[w + 2] >> (32 - sh)) : hi;
}

__device__ __forceinline__ uint32_t base_at(const uint32_t *T, uint64_t p) { return (T[p >> 4] >> (30 - 2 * (int)(p & 15))) & 3u; }

__global__ void k_prefix_hist(const uint32_t *__restrict__ T, uint64_t n, unsigned long long *hist)
{
    __shared__ unsigned int sh[NPFX];
    for (int i = threadIdx.x; i < NPFX; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += stride)
        atomicAdd(&sh[bases64(T, p) >> (64 - 2 * PFX)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < NPFX; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}

// every suffix whose leading PFX-mer lies in [lo, hi): append (key, position); one atomic per warp
__global__ void k_collect(const uint32_t *__restrict__ T, uint64_t n, uint32_t lo, uint32_t hi, uint64_t *keys, uint64_t *vals,
                          unsigned long long *count)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    const uint64_t n_round = (n + 31) & ~31ull;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += stride) {
        uint64_t key = 0;
        bool take = false;
        if (p < n) {
            key = bases64(T, p);
            const uint32_t pf = (uint32_t)(key >> (64 - 2 * PFX));
            take = pf >= lo && pf < hi;
        }
        const unsigned m = __ballot_sync(0xffffffffu, take);
        if (m) {
            unsigned long long base = 0;
            if (lane == __ffs(m) - 1) base = atomicAdd(count, (unsigned long long)__popc(m));
            base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
            if (take) {
                const uint64_t o = base + __popc(m & ((1u << lane) - 1u));
                keys[o] = key; vals[o] = p;
            }
        }
    }
}

__global__ void k_tie_flags(const uint64_t *__restrict__ keys, uint64_t cnt, uint8_t *flags)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) {
        const uint64_t k = keys[i];
        flags[i] = (i > 0 && keys[i - 1] == k) || (i + 1 < cnt && keys[i + 1] == k);
    }
}

__global__ void k_gather_ties(const uint32_t *__restrict__ idx, uint64_t n_tied, const uint64_t *__restrict__ keys,
                              const uint64_t *__restrict__ vals, uint64_t *tk, uint64_t *tv)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tied; i += stride) { tk[i] = keys[idx[i]]; tv[i] = vals[idx[i]]; }
}

__global__ void k_scatter_ties(const uint32_t *__restrict__ idx, uint64_t n_tied, const uint64_t *__restrict__ tv, uint64_t *vals)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_tied; i += stride) vals[idx[i]] = tv[i];
}

// rows [row0, row0+cnt) of the suffix array: BWT symbol per row (4 marks the primary row), every sa_intv-th SA entry
__global__ void k_emit_rows(const uint32_t *__restrict__ T, const uint64_t *__restrict__ vals, uint64_t cnt, uint64_t row0,
                            uint8_t *sym, uint64_t *sa_samp, uint64_t sa_intv, unsigned long long *primary)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += stride) {
        const uint64_t sa = vals[i], row = row0 + i;
        if (sa == 0) { sym[row] = 4; *primary = row; }
        else sym[row] = (uint8_t)base_at(T, sa - 1);
        if (row % sa_intv == 0) sa_samp[row / sa_intv] = sa;
    }
}

struct Cnt4 { uint64_t c[4]; };
struct Cnt4Add { __host__ __device__ Cnt4 operator()(const Cnt4 &a, const Cnt4 &b) const { Cnt4 r; for (int i = 0; i < 4; i++) r.c[i] = a.c[i] + b.c[i]; return r; } };

// one thread per 128-symbol block of the BWT string (primary row dropped): 8 packed words + the block's histogram
__global__ void k_pack_blocks(const uint8_t *__restrict__ sym, uint64_t n, uint64_t primary, uint32_t *words, Cnt4 *hist, uint64_t n_blocks)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += stride) {
        Cnt4 h{};
        for (int w = 0; w < 8; w++) {
            uint32_t x = 0;
            for (int i = 0; i < 16; i++) {
                const uint64_t j = b * 128 + w * 16 + i;
                if (j < n) {
                    const uint32_t s = sym[j + (j >= primary)];
                    x |= s << (30 - 2 * i);
                    h.c[s]++;
                }
            }
            if (b * 128 + (uint64_t)w * 16 < n) words[b * 8 + w] = x;
        }
        hist[b] = h;
    }
}

// the file image: per block 4 x u64 cumulative counts (8 words) then its symbol words; one trailing count record
__global__ void k_interleave(const uint32_t *__restrict__ words, const Cnt4 *__restrict__ cum, uint64_t n, uint64_t n_blocks, uint32_t *image)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n_words = (n + 15) / 16;
    for (uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; b <= n_blocks; b += stride) {
        uint32_t *o = image + b * 16;
        if (b == n_blocks) o = image + n_words + n_blocks * 8;      // after the last (possibly partial) block
        const Cnt4 c = cum[b];
        for (int i = 0; i < 4; i++) { o[2 * i] = (uint32_t)c.c[i]; o[2 * i + 1] = (uint32_t)(c.c[i] >> 32); }
        if (b < n_blocks)
            for (int w = 0; w < 8; w++)
                if (b * 8 + w < n_words) o[8 + w] = words[b * 8 + w];
    }
}

// exact suffix order for ties: the first 32 bases are equal; compare 32 bases at a time, a suffix that runs out of text
// first is the smaller one ('$' sorts before every base)
struct SuffixLess {
    const uint32_t *T; uint64_t n;
    bool operator()(uint64_t a, uint64_t b) const
    {
        if (a == b) return false;
        for (uint64_t off = 32;; off += 32) {
            const uint64_t pa = a + off, pb = b + off;
            if (pa >= n || pb >= n) return a > b;              // equal up to the shorter one's end: shorter (= later start) first
            const uint64_t ka = bases64(T, pa), kb = bases64(T, pb);
            if (ka != kb) {
                // bases beyond the end read as A: a genuine difference inside both suffixes decides; otherwise the
                // one that ends inside this window is the smaller
                const uint64_t la = n - pa, lb = n - pb;       // bases left
                const int first_diff = __builtin_clzll(ka ^ kb) >> 1;
                if ((uint64_t)first_diff < std::min<uint64_t>({la, lb, 32})) return ka < kb;
                return a > b;
            }
            if (n - pa <= 32 || n - pb <= 32) return a > b;
        }
    }
};

static int grid_for(uint64_t n, int threads = 256) { uint64_t g = (n + threads - 1) / threads; return (int)std::max<uint64_t>(1, std::min<uint64_t>(g, (uint64_t)sm_count() * 16)); }

static void write_file(const std::string &fn, const std::vector<std::pair<const void *, size_t>> &parts)
{
    FILE *fp = fopen(fn.c_str(), "wb");
    if (!fp) throw std::make_pair(DARTGPU_ERR_INDEX, "cannot write " + fn);
    for (auto &p : parts)
        if (p.second && fwrite(p.first, 1, p.second, fp) != p.second) { fclose(fp); throw std::make_pair(DARTGPU_ERR_INDEX, "short write to " + fn); }
    fclose(fp);
}

} // namespace

void build_index_files(int device, const uint8_t *pac, int64_t l_pac, const char *prefix, uint64_t limit)
{
    DG_CUDA(cudaSetDevice(device));
    cudaStream_t st = nullptr;
    const uint64_t G = (uint64_t)l_pac, n = 2 * G, sa_intv = 32;
    if (l_pac <= 0) throw std::make_pair(DARTGPU_ERR_ARG, std::string("empty genome"));

    // the text, both strands, 2 bits per base (+ guard words that read as A)
    DevBuf<uint32_t> T;
    const size_t t_words = (size_t)((n + 15) / 16 + 8);
    T.reserve(t_words);
    DG_CUDA(cudaMemsetAsync(T.p, 0, T.cap * 4, st));
    {
        DevBuf<uint8_t> dpac;
        dpac.reserve((size_t)G / 4 + 1);
        DG_CUDA(cudaMemcpyAsync(dpac.p, pac, (size_t)G / 4 + 1, cudaMemcpyHostToDevice, st));
        launch_build_ref2(dpac.p, T.p, (int64_t)G, st);
        DG_CUDA(cudaGetLastError());
        DG_CUDA(cudaStreamSynchronize(st));
    }
    std::vector<uint32_t> hT(t_words);      // the host keeps a copy for the exact tie-break
    DG_CUDA(cudaMemcpy(hT.data(), T.p, t_words * 4, cudaMemcpyDeviceToHost));

    // ranges of leading 6-mers with at most `limit` suffixes each
    DevBuf<unsigned long long> d_hist;
    d_hist.reserve(NPFX + 2);
    DG_CUDA(cudaMemsetAsync(d_hist.p, 0, (NPFX + 2) * 8, st));
    k_prefix_hist<<<grid_for(n), 256, 0, st>>>(T.p, n, d_hist.p);
    std::vector<unsigned long long> hist(NPFX);
    DG_CUDA(cudaMemcpy(hist.data(), d_hist.p, NPFX * 8, cudaMemcpyDeviceToHost));
    if (limit == 0) {
        size_t free_b = 0, total_b = 0;
        DG_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t fixed = (n + 1) + (n / sa_intv + 2) * 8 + (n / 16 + 16) * 8 + (n / 128 + 2) * 64 + (256ull << 20);
        const uint64_t room = free_b > fixed ? free_b - fixed : 0;
        limit = std::min<uint64_t>(room / 48, 1500000000ull);    // keys + values, double-buffered, + sort scratch
    }
    uint64_t biggest = *std::max_element(hist.begin(), hist.end());
    if (limit < biggest) limit = biggest;
    if (limit >= (1ull << 31)) throw std::make_pair(DARTGPU_ERR_NOMEM, std::string("a 6-mer bucket exceeds 2^31 suffixes"));
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    for (uint32_t lo = 0; lo < NPFX;) {
        uint64_t acc = 0;
        uint32_t hi = lo;
        while (hi < NPFX && acc + hist[hi] <= limit) acc += hist[hi++];
        ranges.push_back({lo, hi});
        lo = hi;
    }
    uint64_t max_range = 0;
    for (auto &r : ranges) { uint64_t a = 0; for (uint32_t i = r.first; i < r.second; i++) a += hist[i]; max_range = std::max(max_range, a); }

    DevBuf<uint8_t> sym;          sym.reserve(n + 2);
    DevBuf<uint64_t> sa_samp;     sa_samp.reserve(n / sa_intv + 2);
    DevBuf<unsigned long long> d_primary; d_primary.reserve(2);
    DevBuf<uint64_t> keys, vals, keys2, vals2;
    keys.reserve(max_range + 1); vals.reserve(max_range + 1); keys2.reserve(max_range + 1); vals2.reserve(max_range + 1);
    DevBuf<uint8_t> flags;        flags.reserve(max_range + 1);
    DevBuf<uint32_t> tied_idx;    tied_idx.reserve(max_range + 1);
    DevBuf<unsigned long long> d_count; d_count.reserve(2);
    size_t sort_tmp = 0, sel_tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_tmp, keys.p, keys2.p, vals.p, vals2.p, (int)max_range, 0, 64, st);
    cub::DeviceSelect::Flagged(nullptr, sel_tmp, cub::CountingInputIterator<uint32_t>(0), flags.p, tied_idx.p, (int *)d_count.p, (int)max_range, st);
    DevBuf<uint8_t> tmp;
    tmp.reserve(std::max(sort_tmp, sel_tmp) + 256);

    // row 0 is the empty suffix: SA[0] = n, its BWT symbol is the last base of the text
    {
        const uint8_t last = (uint8_t)((hT[(n - 1) >> 4] >> (30 - 2 * (int)((n - 1) & 15))) & 3u);
        DG_CUDA(cudaMemcpyAsync(sym.p, &last, 1, cudaMemcpyHostToDevice, st));
        DG_CUDA(cudaStreamSynchronize(st));
    }
    uint64_t row0 = 1;
    std::vector<uint32_t> h_idx;
    std::vector<uint64_t> h_tk, h_tv;
    DevBuf<uint64_t> d_tk, d_tv;
    for (auto &r : ranges) {
        DG_CUDA(cudaMemsetAsync(d_count.p, 0, 16, st));
        k_collect<<<grid_for(n), 256, 0, st>>>(T.p, n, r.first, r.second, keys.p, vals.p, d_count.p);
        unsigned long long cnt = 0;
        DG_CUDA(cudaMemcpyAsync(&cnt, d_count.p, 8, cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaStreamSynchronize(st));
        if (cnt == 0) continue;
        size_t tb = tmp.cap;
        // the leading PFX bases only matter across buckets of the range; sort all 64 bits (ranges span several buckets)
        cub::DeviceRadixSort::SortPairs(tmp.p, tb, keys.p, keys2.p, vals.p, vals2.p, (int)cnt, 0, 64, st);
        k_tie_flags<<<grid_for(cnt), 256, 0, st>>>(keys2.p, cnt, flags.p);
        tb = tmp.cap;
        cub::DeviceSelect::Flagged(tmp.p, tb, cub::CountingInputIterator<uint32_t>(0), flags.p, tied_idx.p, (int *)d_count.p, (int)cnt, st);
        int n_tied = 0;
        DG_CUDA(cudaMemcpyAsync(&n_tied, d_count.p, 4, cudaMemcpyDeviceToHost, st));
        DG_CUDA(cudaStreamSynchronize(st));
        if (n_tied > 0) {
            d_tk.reserve(n_tied); d_tv.reserve(n_tied);
            k_gather_ties<<<grid_for(n_tied), 256, 0, st>>>(tied_idx.p, n_tied, keys2.p, vals2.p, d_tk.p, d_tv.p);
            h_tk.resize(n_tied); h_tv.resize(n_tied);
            DG_CUDA(cudaMemcpyAsync(h_tk.data(), d_tk.p, (size_t)n_tied * 8, cudaMemcpyDeviceToHost, st));
            DG_CUDA(cudaMemcpyAsync(h_tv.data(), d_tv.p, (size_t)n_tied * 8, cudaMemcpyDeviceToHost, st));
            DG_CUDA(cudaStreamSynchronize(st));
            // runs of equal keys are contiguous in the compacted list (the flags mark whole runs)
            std::vector<int> starts;
            for (int i = 0; i < n_tied; i++) if (i == 0 || h_tk[i] != h_tk[i - 1]) starts.push_back(i);
            starts.push_back(n_tied);
            SuffixLess less{hT.data(), n};
            const int64_t n_runs = (int64_t)starts.size() - 1;
#pragma omp parallel for schedule(dynamic, 64)
            for (int64_t g = 0; g < n_runs; g++) std::sort(h_tv.begin() + starts[g], h_tv.begin() + starts[g + 1], less);
            DG_CUDA(cudaMemcpyAsync(d_tv.p, h_tv.data(), (size_t)n_tied * 8, cudaMemcpyHostToDevice, st));
            k_scatter_ties<<<grid_for(n_tied), 256, 0, st>>>(tied_idx.p, n_tied, d_tv.p, vals2.p);
        }
        k_emit_rows<<<grid_for(cnt), 256, 0, st>>>(T.p, vals2.p, cnt, row0, sym.p, sa_samp.p, sa_intv, d_primary.p);
        DG_CUDA(cudaGetLastError());
        DG_CUDA(cudaStreamSynchronize(st));
        row0 += cnt;
    }
    if (row0 != n + 1) throw std::make_pair(DARTGPU_ERR_INDEX, std::string("index build: suffix count mismatch"));
    unsigned long long primary = 0;
    DG_CUDA(cudaMemcpy(&primary, d_primary.p, 8, cudaMemcpyDeviceToHost));
    keys.release(); vals.release(); keys2.release(); vals2.release(); flags.release(); tied_idx.release();

    // BWT string -> packed words + per-block histograms -> cumulative counts -> interleaved image
    const uint64_t n_blocks = (n + 127) / 128, n_words = (n + 15) / 16, image_words = n_words + (n_blocks + 1) * 8;
    DevBuf<uint32_t> words;  words.reserve(n_blocks * 8 + 8);
    DevBuf<Cnt4> hist4, cum4; hist4.reserve(n_blocks + 2); cum4.reserve(n_blocks + 2);
    DG_CUDA(cudaMemsetAsync(hist4.p, 0, (n_blocks + 2) * sizeof(Cnt4), st));
    k_pack_blocks<<<grid_for(n_blocks, 128), 128, 0, st>>>(sym.p, n, primary, words.p, hist4.p, n_blocks);
    size_t scan_tmp = 0;
    cub::DeviceScan::ExclusiveScan(nullptr, scan_tmp, hist4.p, cum4.p, Cnt4Add(), Cnt4{}, (int)(n_blocks + 1), st);
    tmp.reserve(scan_tmp + 256);
    cub::DeviceScan::ExclusiveScan(tmp.p, scan_tmp, hist4.p, cum4.p, Cnt4Add(), Cnt4{}, (int)(n_blocks + 1), st);
    sym.release();
    DevBuf<uint32_t> image;  image.reserve(image_words + 16);
    k_interleave<<<grid_for(n_blocks + 1), 256, 0, st>>>(words.p, cum4.p, n, n_blocks, image.p);
    DG_CUDA(cudaGetLastError());
    std::vector<uint32_t> h_image(image_words);
    DG_CUDA(cudaMemcpy(h_image.data(), image.p, image_words * 4, cudaMemcpyDeviceToHost));
    Cnt4 total{};
    DG_CUDA(cudaMemcpy(&total, cum4.p + n_blocks, sizeof(Cnt4), cudaMemcpyDeviceToHost));
    const uint64_t n_sa = (n + sa_intv) / sa_intv;
    std::vector<uint64_t> h_sa(n_sa);
    DG_CUDA(cudaMemcpy(h_sa.data(), sa_samp.p, n_sa * 8, cudaMemcpyDeviceToHost));

    uint64_t head[5] = {primary, total.c[0], total.c[0] + total.c[1], total.c[0] + total.c[1] + total.c[2],
                        total.c[0] + total.c[1] + total.c[2] + total.c[3]};
    const std::string pre(prefix);
    write_file(pre + ".bwt", {{head, sizeof head}, {h_image.data(), image_words * 4}});
    const uint64_t sa_head[2] = {sa_intv, n};
    write_file(pre + ".sa", {{head, sizeof head}, {sa_head, sizeof sa_head}, {h_sa.data() + 1, (n_sa - 1) * 8}});
}

} // namespace dartgpu
