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


// ---------------------------------------------------------------------------------------------------
// read batch: raw ASCII -> device codes, on the device (the host only memcpy's the caller's bases into pinned memory)
// ---------------------------------------------------------------------------------------------------
// padded length of every read (each read starts on a 16-byte boundary so the search kernel stages 16 bases per load)
__global__ void k_read_layout(const int64_t *__restrict__ off, int n, int32_t *rlen, uint32_t *padded)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += gridDim.x * blockDim.x) {
        if (i == n) { padded[i] = 0; continue; }
        int rl = (int)(off[i + 1] - off[i]);
        rlen[i] = rl;
        padded[i] = (uint32_t)((rl + 15) & ~15);
    }
}

// codes: 0..3 = ACGT, 8..11 = acgt, 5 = 'N', 4 = anything else (dartgpu_internal.h)
__global__ void k_encode_reads(const uint8_t *__restrict__ raw, const int64_t *__restrict__ off, const int64_t *__restrict__ dev_off,
                               int n, uint8_t *codes)
{
    __shared__ uint8_t tab[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        uint8_t v = 4;
        switch (i) {
        case 'A': v = 0; break; case 'C': v = 1; break; case 'G': v = 2; break; case 'T': v = 3; break;
        case 'a': v = 8; break; case 'c': v = 9; break; case 'g': v = 10; break; case 't': v = 11; break;
        case 'N': v = 5; break;
        }
        tab[i] = v;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int64_t base0 = off[0];
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nwarps) {
        const int64_t src = off[r] - base0;
        const int rl = (int)(off[r + 1] - off[r]);
        const int chunks = (rl + 15) >> 4;
        uint4 *dst = reinterpret_cast<uint4 *>(codes + dev_off[r]);
        for (int ch = lane; ch < chunks; ch += 32) {
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; k++) {
                int p = ch * 16 + k;
                uint32_t c = p < rl ? tab[raw[src + p]] : 4u;
                w[k >> 2] |= c << (8 * (k & 3));
            }
            dst[ch] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

void launch_read_layout(const int64_t *off, int n, int32_t *rlen, uint32_t *padded, cudaStream_t st)
{
    int grid = (n + 1 + 255) / 256; if (grid > 148 * 8) grid = 148 * 8;
    k_read_layout<<<grid, 256, 0, st>>>(off, n, rlen, padded);
}
void launch_encode_reads(const uint8_t *raw, const int64_t *off, const int64_t *dev_off, int n, uint8_t *codes, cudaStream_t st)
{
    int64_t want = ((int64_t)n * 32 + 255) / 256;
    int grid = (int)(want < 148 * 16 ? want : 148 * 16); if (grid < 1) grid = 1;
    k_encode_reads<<<grid, 256, 0, st>>>(raw, off, dev_off, n, codes);
}

} // namespace dartgpu
