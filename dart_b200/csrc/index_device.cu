// One-time GPU-side re-layout of the BWA-format index into the HBM-resident tables the kernels read.
// File formats: /root/reference/src/bwt_index.cpp:15-35 (.sa), :102-121 (.bwt), :193-212 (.pac decode);
// block interleave written by /root/reference/src/BWT_Index/bwtindex.c:53-75.
#include "dartgpu_internal.h"

namespace dartgpu {

// BWA block = 8 words of counts (4 x u64, little endian) + up to 8 words of symbols (16 per word, first
// symbol in the top bits).  Output quarter q = { count[q], symbols[32q..32q+32) as one u64 }.
__global__ void k_relayout_occ(const uint32_t *__restrict__ w, ulonglong2 *__restrict__ occ, uint64_t n_quarters)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; t < n_quarters; t += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t b = t >> 2;
        int q = (int)(t & 3);
        const uint32_t *p = w + b * 16;
        ulonglong2 o;
        o.x = (uint64_t)p[2 * q] | (uint64_t)p[2 * q + 1] << 32;
        o.y = (uint64_t)p[8 + 2 * q] << 32 | (uint64_t)p[8 + 2 * q + 1];
        occ[t] = o;
    }
}

void launch_relayout_occ(const uint32_t *bwt_words, ulonglong2 *occ, uint64_t n_blocks, cudaStream_t st)
{
    uint64_t nq = n_blocks * 4;
    int grid = (int)((nq + 255) / 256 < 148 * 16 ? (nq + 255) / 256 : 148 * 16);
    if (grid < 1) grid = 1;
    k_relayout_occ<<<grid, 256, 0, st>>>(bwt_words, occ, nq);
}

// RefSequence over both strands, 2 bits per base: position p < G is the .pac base, position p >= G is the
// complement of position 2G-1-p (/root/reference/src/bwt_index.cpp:199-209).
__device__ __forceinline__ uint32_t pac_base(const uint8_t *pac, int64_t p)
{
    return (pac[p >> 2] >> ((~p & 3) << 1)) & 3;
}

__global__ void k_build_ref2(const uint8_t *__restrict__ pac, uint32_t *__restrict__ ref2, int64_t G, int64_t n_words)
{
    int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        uint32_t v = 0;
        for (int i = 0; i < 16; i++) {
            int64_t p = w * 16 + i;
            uint32_t c = 0;
            if (p < G) c = pac_base(pac, p);
            else if (p < 2 * G) c = 3 - pac_base(pac, 2 * G - 1 - p);
            v |= c << (30 - 2 * i);
        }
        ref2[w] = v;
    }
}

void launch_build_ref2(const uint8_t *pac, uint32_t *ref2, int64_t G, cudaStream_t st)
{
    int64_t n_words = (2 * G + 15) / 16 + 2; // two guard words so 64-bit window reads never run off the end
    int64_t blocks = (n_words + 255) / 256;
    int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
    if (grid < 1) grid = 1;
    k_build_ref2<<<grid, 256, 0, st>>>(pac, ref2, G, n_words);
}

} // namespace dartgpu
