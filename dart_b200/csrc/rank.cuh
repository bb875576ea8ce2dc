// Rank arithmetic on one quarter of an Occ block (32 BWT symbols in a u64, first symbol in the top bits), shared by the
// search and locate kernels and unit-tested on the host (tests/host/rank_test.cpp).
//
// The forward-extension step of BWT_Search (/root/reference/src/bwt_search.cpp:152-170) needs, for the read's next base c,
// only  Occ(c,·)  (new interval on the reverse strand) and  sum over symbols > c of Occ(·)  (shift of the forward-strand
// interval), not the four counts bwt_occ4 produces: two indicator words, two POPCs per block.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RK_HD __host__ __device__ __forceinline__
#else
#define RK_HD inline
#endif

namespace dartgpu {

RK_HD int popc32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
    return __popc(x);
#else
    return __builtin_popcount(x);
#endif
}

// indicators live on even bit positions only: fold the two halves of a u64 into one u32 -> a single POPC
RK_HD uint32_t fold_even(uint64_t x) { return (uint32_t)x | ((uint32_t)(x >> 32) << 1); }

// among the first n (0..32) symbols of b: eq = number equal to c, gt = number greater than c
RK_HD void count_eq_gt(uint64_t b, int n, int c, int &eq, int &gt)
{
    const uint64_t M5 = 0x5555555555555555ull;
    uint64_t v = n > 0 ? b >> (64 - 2 * n) : 0ull;   // the n symbols right-aligned; everything above reads as 0 = 'A'
    uint64_t lo = v & M5, hi = (v >> 1) & M5;
    uint64_t CH = (c & 2) ? M5 : 0ull, CL = (c & 1) ? M5 : 0ull;
    uint64_t eqm = ~(hi ^ CH) & ~(lo ^ CL) & M5;
    uint64_t gtm = (hi & ~CH) | (~(hi ^ CH) & lo & ~CL);
    eq = popc32(fold_even(eqm)) - (c == 0 ? 32 - n : 0);
    gt = popc32(fold_even(gtm));
}

RK_HD int count_eq(uint64_t b, int n, int c)
{
    const uint64_t M5 = 0x5555555555555555ull;
    uint64_t v = n > 0 ? b >> (64 - 2 * n) : 0ull;
    uint64_t lo = v & M5, hi = (v >> 1) & M5;
    uint64_t CH = (c & 2) ? M5 : 0ull, CL = (c & 1) ? M5 : 0ull;
    return popc32(fold_even(~(hi ^ CH) & ~(lo ^ CL) & M5)) - (c == 0 ? 32 - n : 0);
}

} // namespace dartgpu
