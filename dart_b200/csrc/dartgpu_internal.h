// Internal declarations shared by the CUDA kernels, the host orchestration and the C-ABI glue.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/dartgpu.h"

namespace dartgpu {

// ---------------------------------------------------------------------------------------------------
// Device-resident index (replicated into every GPU's HBM once, at context creation)
// ---------------------------------------------------------------------------------------------------
// Occ table: one 64-byte block per 128 BWT symbols, re-laid-out from the BWA block (4 x u64 counts followed
// by 8 x u32 of symbols, /root/reference/src/BWT_Index/bwtindex.c:53-75) into four 16-byte quarters
//     quarter q = { u64 count of symbol q before the block, u64 symbols [32q, 32q+32) of the block }
// so that the 4 lanes of a search group each issue ONE 128-bit load (together one 64-byte segment) and
// each lane owns the count it needs plus a quarter of the popcount work.  Symbol j of a quarter sits at
// bits 62-2j (first symbol in the top bits, as in the BWA words).
struct DevIndex {
    const ulonglong2 *occ;   // n_blocks * 4 quarters
    uint64_t n_blocks;
    const uint64_t *sa;      // sampled suffix array, sa[0] = (uint64_t)-1 (/root/reference/src/bwt_index.cpp:31)
    uint64_t sa_mask;        // sa_intv - 1
    int sa_shift;            // log2(sa_intv)
    uint64_t primary, seq_len;
    uint64_t L2[5];
    const uint32_t *ref2;    // reference over [0,2G), 2 bits/base, 16 bases per word, base i at bits 30-2(i&15)
    int64_t G;
    const int64_t *chr_ends; // sorted last coordinates of every sequence on both strands (ChrLocMap keys)
    int n_ends;
    int force64;             // tests only (DARTGPU_FORCE_IDX64=1): run the 64-bit interval kernels on a small index
};

struct SearchRec {           // one qualifying BWT_Search result: SA interval (of the reverse-complemented match) still to be located
    uint64_t sa_begin;
    uint32_t freq;
    uint16_t start, len;
};

// seed key: gPos << 31 | rPos << 15 | len  — sorting the keys sorts by (gPos, rPos) as CompByGenomePos does
__host__ __device__ inline uint64_t seed_key(uint64_t g, uint32_t r, uint32_t len) { return g << 31 | (uint64_t)r << 15 | len; }
__host__ __device__ inline int64_t key_gpos(uint64_t k) { return (int64_t)(k >> 31); }
__host__ __device__ inline int key_rpos(uint64_t k) { return (int)((k >> 15) & 0xFFFF); }
__host__ __device__ inline int key_len(uint64_t k) { return (int)(k & 0x7FFF); }

// device read encoding (one byte per base): 0..3 = A,C,G,T, 8..11 = a,c,g,t (bit 3 = lower case: the reference compares
// raw characters in places), 4 = any other symbol, 5 = a literal 'N' (the only symbol that breaks an 8-mer,
// /root/reference/src/KmerAnalysis.cpp:44).  Bit 2 set <=> not ACGT; (code & 3) is the base otherwise.
enum { CODE_OTHER = 4, CODE_N = 5 };

struct DevStats {            // device-side work counters (see dartgpu_stats)
    unsigned long long ext_steps, ext_blocks, lf_steps, hits, seeds;
};

// ---------------------------------------------------------------------------------------------------
// small RAII buffers
// ---------------------------------------------------------------------------------------------------
struct CudaError { cudaError_t e; const char *what; const char *file; int line; };
#define DG_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw ::dartgpu::CudaError{e_, #x, __FILE__, __LINE__}; } while (0)

template <class T> struct DevBuf {
    T *p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        DG_CUDA(cudaMalloc((void **)&p, want * sizeof(T)));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    ~DevBuf() { release(); }
};
template <class T> struct PinBuf {
    T *p = nullptr; size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        DG_CUDA(cudaMallocHost((void **)&p, want * sizeof(T)));
        cap = want;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    ~PinBuf() { release(); }
};

// ---------------------------------------------------------------------------------------------------
// kernel launchers (each file owns its kernels; all work is enqueued on `st`)
// ---------------------------------------------------------------------------------------------------
// index_device.cu
void launch_relayout_occ(const uint32_t *bwt_words, ulonglong2 *occ, uint64_t n_blocks, cudaStream_t st);
void launch_build_ref2(const uint8_t *pac, uint32_t *ref2, int64_t G, cudaStream_t st);
void launch_read_layout(const int64_t *off, int n, int32_t *rlen, uint32_t *padded, cudaStream_t st);
void launch_encode_reads(const uint8_t *raw, const int64_t *off, const int64_t *dev_off, int n, uint8_t *codes, cudaStream_t st);

// seed_kernels.cu
struct SeedLaunch {
    const uint8_t *codes; const int64_t *dev_off; const int32_t *rlen; int n_reads;
    int cap_rec; uint32_t max_dup; int max_gaps, max_intron;
    SearchRec *recs; uint32_t *nrec; uint32_t *nhits;          // search output
    int64_t *seed_off;                                          // n_reads+1, exclusive scan of nhits
    uint64_t *keys; uint32_t *meta;                             // per seed slot
    int32_t *cand_begin, *cand_count, *cand_score; uint32_t *ncand;
    uint32_t *mid_list; uint32_t *mid_count;                    // reads with 9..32 seeds (one warp each)
    uint32_t *big_list; uint32_t *big_count;                    // reads with more than 32 seeds (one block each)
    uint64_t *big_scratch; size_t big_scratch_per_cta;          // global sort scratch for reads that exceed smem
    DevStats *stats;
};
void launch_search(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st);
void launch_scan_hits(const SeedLaunch &a, void *tmp, size_t tmp_bytes, cudaStream_t st);
size_t scan_tmp_bytes(int n);
void launch_expand_locate(const DevIndex &ix, const SeedLaunch &a, int64_t total_seeds, cudaStream_t st);
void launch_sort_cluster(const DevIndex &ix, const SeedLaunch &a, cudaStream_t st);
void launch_scan_u32_to_i64(const uint32_t *in, int64_t *out, int n, void *tmp, size_t tmp_bytes, cudaStream_t st);

// nw_kernel.cu
struct NwJobDev { int64_t s1_off; int64_t gpos; int64_t op_off; int64_t flag_off; int64_t aux_off; int32_t m, n; };
void launch_nw(const DevIndex &ix, const uint8_t *codes, const NwJobDev *jobs, int n_jobs,
               uint32_t *flags, int32_t *rowbuf, size_t rowbuf_per_warp, uint8_t *ops, int32_t *nops, cudaStream_t st);
int nw_grid_warps();

// kmer_kernel.cu
struct KmerJobDev { int64_t s1_off; int64_t gpos; int32_t len1, len2; };
void launch_kmer(const DevIndex &ix, const uint8_t *codes, const KmerJobDev *jobs, int n_jobs, int max_len1,
                 dartgpu_kmer_hit *out, cudaStream_t st);

} // namespace dartgpu
