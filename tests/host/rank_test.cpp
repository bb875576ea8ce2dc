// TEST INFRASTRUCTURE: exhaustive-in-(n,c) check of dart_b200/csrc/rank.cuh against a symbol-by-symbol count.
#include <cstdio>
#include <cstdlib>
#include <random>
#include "../../dart_b200/csrc/rank.cuh"
int main()
{
    std::mt19937_64 rng(1234);
    long checked = 0;
    for (int it = 0; it < 20000; it++) {
        uint64_t b = rng();
        if (it % 7 == 0) b = 0; if (it % 11 == 0) b = ~0ull; if (it % 13 == 0) b = 0x5555555555555555ull << (it & 1);
        for (int n = 0; n <= 32; n++)
            for (int c = 0; c < 4; c++) {
                int eq = 0, gt = 0;
                for (int j = 0; j < n; j++) { int s = (int)((b >> (62 - 2 * j)) & 3); eq += s == c; gt += s > c; }
                int e2, g2;
                dartgpu::count_eq_gt(b, n, c, e2, g2);
                if (e2 != eq || g2 != gt || dartgpu::count_eq(b, n, c) != eq) { printf("MISMATCH b=%llx n=%d c=%d: %d/%d vs %d/%d\n", (unsigned long long)b, n, c, e2, g2, eq, gt); return 1; }
                checked++;
            }
    }
    printf("rank ok: %ld cases\n", checked);
    return 0;
}
