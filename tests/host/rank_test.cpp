// TEST INFRASTRUCTURE: exhaustive-in-(n,c) check of dart_b200/csrc/rank.cuh against a symbol-by-symbol count.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../dart_b200/csrc/rank.cuh"
int main()
{
    std::mt19937_64 rng(1234);
    long checked = 0;
    // Occ32 (one-sector block, one thread per query): planes of 64 symbols, inclusive prefix counts
    for (int it = 0; it < 4000; it++) {
        uint64_t lo = rng(), hi = rng();
        if (it % 7 == 0) lo = 0; if (it % 11 == 0) hi = ~0ull; if (it % 13 == 0) { lo = ~0ull; hi = 0; }
        for (uint32_t t = 0; t < 64; t++)
            for (int c = 0; c < 4; c++) {
                uint32_t eq = 0;
                for (uint32_t j = 0; j <= t; j++) eq += dartgpu::occ32_symbol(lo, hi, j) == c;
                if (dartgpu::occ32_eq_upto(lo, hi, c, t) != eq) { printf("OCC32 MISMATCH t=%u c=%d\n", t, c); return 1; }
                checked++;
            }
        for (uint32_t j = 0; j < 64; j++)
            if (dartgpu::occ32_symbol(lo, hi, j) != (int)(((lo >> j) & 1) | (((hi >> j) & 1) << 1))) return 1;
    }
    printf("rank ok: %ld cases\n", checked);
    return 0;
}
