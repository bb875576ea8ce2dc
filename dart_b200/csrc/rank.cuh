// Rank arithmetic of the search / locate / index-load kernels, shared with the host unit test (tests/host/rank_test.cpp).
//
// The forward-extension step of BWT_Search (/root/reference/src/bwt_search.cpp:152-170) needs, for the read's next base c,
// only Occ(c,.) at the two ends of the interval once the search is restated as a backward search of the reverse complement
// (seed_kernels.cu) — not the four counts bwt_occ4 produces.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RK_HD __host__ __device__ __forceinline__
#else
#define RK_HD inline
#endif

namespace dartgpu {

// ---------------------------------------------------------------------------------------------------
// Occ32: the one-sector checkpoint block the thread-per-chain kernels read (round-1 ncu of the 4-lane kernel: ~100
// warp-instructions per lane per step, integer-pipe bound; a block that one thread can rank alone removes the
// shuffles and three quarters of the thread-instructions, and halves the bytes a rank query touches).
//     32 bytes per 64 BWT symbols = ONE 32-byte sector:
//         u32 cnt[4]   occurrences of A,C,G,T before the block (fits: every symbol occurs < 2^32 times, checked at load)
//         u64 lo, hi   bit-planes of the 64 symbols: symbol i = (hi >> i & 1) << 1 | (lo >> i & 1)
// A rank query = one 256-bit load of the block (LDG.E.256 on sm_100).
// ---------------------------------------------------------------------------------------------------
struct Occ32 { uint32_t cnt[4]; uint64_t lo, hi; };

RK_HD int popc64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}

// number of symbols == c among symbols 0..t (inclusive, t = 0..63) of a block
RK_HD uint32_t occ32_eq_upto(uint64_t lo, uint64_t hi, int c, uint32_t t)
{
    const uint64_t nCL = (c & 1) ? 0ull : ~0ull, nCH = (c & 2) ? 0ull : ~0ull;
    const uint64_t m = (2ull << t) - 1ull;          // t = 63: 2<<63 wraps to 0, minus 1 = all ones
    return (uint32_t)popc64((hi ^ nCH) & (lo ^ nCL) & m);
}

RK_HD int occ32_symbol(uint64_t lo, uint64_t hi, uint32_t t) { return (int)(((lo >> t) & 1ull) | (((hi >> t) & 1ull) << 1)); }

} // namespace dartgpu
