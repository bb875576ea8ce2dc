// TEST INFRASTRUCTURE: exhaustive-in-(n,c) check of dart_b200/csrc/rank.cuh against a symbol-by-symbol count.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../dart_b200/csrc/rank.cuh"
int main()
{
    std::mt19937_64 rng(1234);
    long checked = 0;
    uint32_t tab[33];
    for (int n = 0; n <= 32; n++) tab[n] = dartgpu::prefix_mask32(n);
    for (int it = 0; it < 20000; it++) {
        uint64_t b = rng();
        if (it % 7 == 0) b = 0; if (it % 11 == 0) b = ~0ull; if (it % 13 == 0) b = 0x5555555555555555ull << (it & 1);
        uint32_t lo, hi;
        dartgpu::planes32(b, lo, hi);
        for (int n = 0; n <= 32; n++)
            for (int c = 0; c < 4; c++) {
                int eq = 0, gt = 0;
                for (int j = 0; j < n; j++) { int s = dartgpu::symbol_at(b, j); eq += s == c; gt += s > c; }
                int e2, g2;
                uint32_t CH = (c & 2) ? ~0u : 0u, CL = (c & 1) ? ~0u : 0u;
                dartgpu::count_eq_gt32(lo, hi, tab[n], CH, CL, e2, g2);
                if (e2 != eq || g2 != gt || dartgpu::count_eq32(lo, hi, tab[n], CH, CL) != eq) {
                    printf("MISMATCH b=%llx n=%d c=%d: %d/%d vs %d/%d\n", (unsigned long long)b, n, c, e2, g2, eq, gt); return 1; }
                checked++;
            }
    }
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
