// Rank arithmetic on one quarter of an Occ block (32 BWT symbols in a u64, first symbol in the top bits), shared by the
// search and locate kernels and unit-tested on the host (tests/host/rank_test.cpp).
//
// The forward-extension step of BWT_Search (/root/reference/src/bwt_search.cpp:152-170) needs, for the read's next base c,
// only  Occ(c,·)  (new interval on the reverse strand) and  sum over symbols > c of Occ(·)  (shift of the forward-strand
// interval), not the four counts bwt_occ4 produces.  Everything is done on 32-bit words (ncu, round 1: the first version
// spent ~100 of its 305 warp-instructions per step emulating 64-bit shifts and logic):
//   * the 2-bit symbols are split into two 32-bit bit-planes (low bits, high bits), both halves of the u64 folded into
//     one word: symbol i < 16 sits at bit 31-2i, symbol 16+i at bit 30-2i;
//   * "equal to c" and "greater than c" are 3-input boolean functions of the planes (LOP3);
//   * "among the first n symbols" is an AND with a 33-entry mask table; 2 POPC per block.
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

// bit-planes of a 32-symbol word
RK_HD void planes32(uint64_t b, uint32_t &lo, uint32_t &hi)
{
    const uint32_t bl = (uint32_t)b, bh = (uint32_t)(b >> 32);
    lo = (bl & 0x55555555u) | ((bh << 1) & 0xAAAAAAAAu);
    hi = ((bl >> 1) & 0x55555555u) | (bh & 0xAAAAAAAAu);
}

// plane bits of the first n (0..32) symbols
RK_HD uint32_t prefix_mask32(int n)
{
    uint32_t m = 0;
    for (int j = 0; j < n; j++) m |= j < 16 ? 1u << (31 - 2 * j) : 1u << (30 - 2 * (j - 16));
    return m;
}

// CH / CL: all-ones when bit 1 / bit 0 of c is set.  eq = #symbols == c, gt = #symbols > c under mask m.
RK_HD void count_eq_gt32(uint32_t lo, uint32_t hi, uint32_t m, uint32_t CH, uint32_t CL, int &eq, int &gt)
{
    const uint32_t t1 = hi ^ CH, t2 = lo ^ CL;
    eq = popc32(~t1 & ~t2 & m);
    gt = popc32(((hi & ~CH) | (~t1 & lo & ~CL)) & m);
}

RK_HD int count_eq32(uint32_t lo, uint32_t hi, uint32_t m, uint32_t CH, uint32_t CL)
{
    return popc32(~(hi ^ CH) & ~(lo ^ CL) & m);
}

// the symbol at index j (0..31) of a quarter word
RK_HD int symbol_at(uint64_t b, int j) { return (int)((b >> (62 - 2 * j)) & 3); }


// ---------------------------------------------------------------------------------------------------
// Occ32: the one-sector checkpoint block the thread-per-chain kernels read (round-1 ncu of the 4-lane kernel: ~100
// warp-instructions per lane per step, integer-pipe bound; a block that one thread can rank alone removes the
// shuffles and three quarters of the thread-instructions, and halves the bytes a rank query touches).
//     32 bytes per 64 BWT symbols = ONE 32-byte sector:
//         u32 cnt[4]   occurrences of A,C,G,T before the block (fits: every symbol occurs < 2^32 times, checked at load)
//         u64 lo, hi   bit-planes of the 64 symbols: symbol i = (hi >> i & 1) << 1 | (lo >> i & 1)
// A rank query = one 128-bit load of the planes + one 32-bit load of cnt[c] from the same sector.
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
