// TEST TOOLING: order-independent fingerprint of the records of a SAM file (header lines skipped), for comparing the
// reference's multithreaded output (records in chunk-completion order) with ours (input order) at sizes
// where sorting 10 GB of text is not an option.  Prints: records, sum and xor of the 64-bit FNV-1a hashes of the lines.
// With a second argument, also writes every line hash (u64, file order) there, so differing records can be located.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
int main(int argc, char **argv)
{
    if (argc < 2) { fprintf(stderr, "usage: samhash file.sam [hashes.bin]\n"); return 1; }
    FILE *fp = fopen(argv[1], "rb");
    if (!fp) { perror(argv[1]); return 1; }
    FILE *out = argc > 2 ? fopen(argv[2], "wb") : nullptr;
    std::vector<char> buf(1 << 24);
    std::vector<uint64_t> hs;
    uint64_t n = 0, sum = 0, x = 0, h = 1469598103934665603ull;
    bool at_start = true, header = false;
    size_t got;
    while ((got = fread(buf.data(), 1, buf.size(), fp)) > 0)
        for (size_t i = 0; i < got; i++) {
            const char c = buf[i];
            if (at_start) { header = c == '@'; at_start = false; h = 1469598103934665603ull; }
            if (c == '\n') {
                if (!header) { n++; sum += h; x ^= h; if (out) { hs.push_back(h); if (hs.size() == (1u << 20)) { fwrite(hs.data(), 8, hs.size(), out); hs.clear(); } } }
                at_start = true;
            } else if (!header) h = (h ^ (unsigned char)c) * 1099511628211ull;
        }
    if (out) { fwrite(hs.data(), 8, hs.size(), out); fclose(out); }
    fclose(fp);
    printf("%llu %016llx %016llx\n", (unsigned long long)n, (unsigned long long)sum, (unsigned long long)x);
    return 0;
}
