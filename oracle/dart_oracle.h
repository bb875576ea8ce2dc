/* TEST INFRASTRUCTURE — the CPU oracle. Never linked, imported or executed by the product path.
 *
 * A from-scratch CPU restatement of the arithmetic on DART's per-read mapping hot path
 * (SURVEY.md §8a), written in the same reformulated shape the CUDA kernels use so that each
 * reformulation is proven equal to the reference on the CPU before it is trusted on the GPU:
 *
 *   or_rank4 / or_lf / or_locate     bwt_occ4, bwt_2occ4, bwt_occ, bwt_invPsi, bwt_sa
 *                                    (/root/reference/src/bwt_search.cpp:43-137)
 *   or_search                        BWT_Search            (bwt_search.cpp:139-182)
 *   or_seed_read                     IdentifySeedPairs     (AlignmentCandidates.cpp:181-215)
 *   or_cluster_read                  GenerateAlignmentCandidate (AlignmentCandidates.cpp:241-288)
 *   or_kmer_pair                     GenerateLongestSimplePairsFromFragmentPair and its helpers
 *                                    (KmerAnalysis.cpp:25-166)
 *   or_nw                            nw_alignment          (nw_alignment.cpp:8-82; integer half-units)
 *   or_gapped_partition              IdentifyBestGappedPartition (AlignmentCandidates.cpp:385-467)
 *
 * PINNING: the reference ships no tests or golden vectors for this path, so the
 * oracle is pinned against the reference itself: oracle/_ref/libdartref.so (the reference's own
 * objects behind oracle/ref_taps.cpp) in tests/test_oracle_vs_reference.py, and against vectors
 * generated from it and committed under tests/golden/ (tools/make_golden.py).
 */
#ifndef DART_ORACLE_H
#define DART_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct or_index or_index;

/* work counters: the "algorithmic bytes / cells" of SURVEY.md §8d are computed from these */
typedef struct {
    uint64_t searches;      /* BWT_Search calls                                  */
    uint64_t ext_steps;     /* forward-extension steps (bwt_2occ4 calls)          */
    uint64_t ext_blocks;    /* 64-byte Occ blocks those steps touch (1 or 2 each) */
    uint64_t lf_steps;      /* bwt_invPsi calls during locate (1 block each)      */
    uint64_t hits;          /* SA locates                                         */
    uint64_t seeds;         /* seeds written                                      */
    uint64_t read_bases;    /* bases of reads seeded                              */
    uint64_t nw_calls, nw_cells;
    uint64_t kmer_calls, kmer_window_bases, kmer_read_bases;
    uint64_t lf_steps_rc;   /* LF steps when the same hits are located from the reverse-complement interval (what the CUDA path walks) */
} or_counters;

or_index *or_load(const char *prefix);           /* <prefix>.bwt .sa .pac .ann */
void      or_free(or_index *);
int64_t   or_genome_size(const or_index *);      /* G; coordinates live in [0,2G) */
int       or_num_chromosomes(const or_index *);
void      or_counters_get(const or_index *, or_counters *out);
void      or_counters_reset(or_index *);
/* reference bases as codes 0..3 over the doubled coordinate space [0,2G) */
void      or_ref_codes(const or_index *, int64_t pos, int len, uint8_t *out);

void      or_rank4(const or_index *, uint64_t k, uint64_t cnt[4]);   /* bwt_occ4 */
uint64_t  or_locate(or_index *, uint64_t k);                         /* bwt_sa   */

/* BWT_Search: returns freq, *len = match length (0 when too repetitive); locs in SA order */
int or_search(or_index *, const uint8_t *codes, int start, int stop, int max_dup,
              int *len, uint64_t *locs, int cap);
/* The same search located from the other half of the bi-interval: positions of revcomp(match), mirrored back with
 * p -> 2G - p - len.  Must be the same set as or_search's locs (the text is its own reverse complement). */
int or_search_mirrored(or_index *, const uint8_t *codes, int start, int stop, int max_dup,
                       int *len, uint64_t *locs, int cap);
/* IdentifySeedPairs: seeds sorted by (gPos,rPos). Returns count (may exceed cap; only cap are written) */
int or_seed_read(or_index *, const uint8_t *codes, int rlen, int max_dup,
                 int32_t *rpos, int64_t *gpos, int32_t *len, int cap);
/* GenerateAlignmentCandidate over sorted seeds: candidates are runs [c_begin, c_begin+c_count) of the seed array */
int or_cluster_read(const or_index *, int rlen, int nseeds, const int32_t *rpos, const int64_t *gpos,
                    const int32_t *len, int max_gaps, int max_intron,
                    int32_t *c_begin, int32_t *c_count, int32_t *c_score, int cap);
/* GenerateLongestSimplePairsFromFragmentPair: frag1 = read chars, frag2 = genome chars. out = {rPos,gPos,len} */
void or_kmer_pair(or_index *, int len1, const char *frag1, int len2, const char *frag2, int64_t out3[3]);
/* nw_alignment: ops[] in left-to-right order, one per column: 0 = both advance, 1 = gap in s1 ('-' in the
 * read string, consumes s2), 2 = gap in s2. Returns the number of columns. */
int or_nw(or_index *, int m, const char *s1, int n, const char *s2, uint8_t *ops);
/* IdentifyBestGappedPartition. seq = read chars. out = {p, left_ext, right_ext} */
void or_gapped_partition(or_index *, const char *seq, int rGaps, int l_rpos, int l_rlen, int64_t l_gpos,
                         int l_glen, int r_rpos, int64_t r_gpos, int max_mismatch, int out3[3]);

#ifdef __cplusplus
}
#endif
#endif
