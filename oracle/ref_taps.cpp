// TEST INFRASTRUCTURE — not product code.
//
// Stage taps over the UNMODIFIED reference: thin extern "C" wrappers that call
// the reference's own free functions (declared in /root/reference/src/structure.h:187-233)
// so tests can pin the oracle restatement (oracle/dart_oracle.cpp) and the CUDA
// path against what the reference really computes, stage by stage.
//
// Built by oracle/Makefile into oracle/_ref/libdartref.so by linking the reference's
// objects compiled in place from /root/reference/src (nothing is copied into the repo).
// `main.cpp` is compiled with -Dmain=dart_reference_main so its globals
// (Refbwt, MaxGaps, MaxDupNum, ... src/main.cpp:9-18) exist in the library.
//
// The per-read driver below (ref_map_*) replays the body of ReadMapping()
// (src/Mapping.cpp:600-643) call for call; the loop body itself cannot be linked
// because it is welded to the FASTQ reader and the output lock.
#include "structure.h"
#include <stdint.h>

// non-static reference functions that structure.h does not declare
extern void SetSingleAlignmentFlag(ReadItem_t& read);
extern void SetPairedAlignmentFlag(ReadItem_t& read1, ReadItem_t& read2);
extern void EvaluateMAPQ(ReadItem_t& read);
extern void RemoveRedundantCandidates(vector<AlignmentCandidate_t>& AlignmentVec);
extern bool CheckPairedAlignmentCandidates(vector<AlignmentCandidate_t>& v1, vector<AlignmentCandidate_t>& v2);
extern void RemoveUnMatedAlignmentCandidates(vector<AlignmentCandidate_t>& v1, vector<AlignmentCandidate_t>& v2);
extern void CheckPairedFinalAlignments(ReadItem_t& read1, ReadItem_t& read2);
extern void UpdateLocalSJMap(AlignmentCandidate_t& Aln, map<pair<int64_t, int64_t>, SpliceJunction_t>& LocalSJMap);
extern void OutputPairedAlignments(ReadItem_t& read1, ReadItem_t& read2, int& u, int& un, int& p, vector<string>& out);
extern void OutputSingledAlignments(ReadItem_t& read, int& u, int& un, vector<string>& out);
extern void GetComplementarySeq(int len, char* seq, char* rseq);
extern bwtint_t bwt_sa(bwtint_t k);
extern void bwt_2occ4(const bwt_t *bwt, bwtint_t k, bwtint_t l, bwtint_t cntk[4], bwtint_t cntl[4]);
extern int iChromsomeNum;

typedef struct { int p; int left_ext; int right_ext; } GappedExtension_t; // src/AlignmentCandidates.cpp:8-13
extern GappedExtension_t IdentifyBestGappedPartition(char* seq, int rGaps, SeedPair_t& LeftSeed, SeedPair_t& RightSeed);

extern "C" {

// Same defaults main() assigns before parsing flags (src/main.cpp:101-117).
int ref_load(const char* prefix, int threads)
{
	MaxGaps = 5; MaxDupNum = 100; iThreadNum = threads > 0 ? threads : 4;
	bPairEnd = false; bDebugMode = false; bMultiHit = false; bUnique = false; bSilent = true;
	bFindAllJunction = false; MaxIntronSize = 500000; MinIntronSize = 5; MaxMismatch = 0;
	OutputFileFormat = 0; FastQFormat = true;
	IndexFileName = (char*)prefix;
	if (!CheckBWAIndexFiles(prefix)) return -1;
	RefIdx = bwa_idx_load(prefix);
	if (RefIdx == 0) return -2;
	Refbwt = RefIdx->bwt;
	RestoreReferenceInfo();
	return 0;
}

void ref_set_params(int max_mismatch, int max_dup, int max_intron, int min_intron,
                    int multi_hit, int pair_end, int all_sj, int unique)
{
	MaxMismatch = max_mismatch;
	// clamp exactly as the flag parser does (src/main.cpp:173-178, :187)
	MaxDupNum = (unsigned int)max_dup; if (MaxDupNum < 100) MaxDupNum = 100; else if (MaxDupNum >= 10000) MaxDupNum = 10000;
	if ((MaxIntronSize = max_intron) < 100000) MaxIntronSize = 100000;
	MinIntronSize = min_intron;
	bMultiHit = multi_hit != 0; bPairEnd = pair_end != 0; bFindAllJunction = all_sj != 0; bUnique = unique != 0;
}

int64_t ref_genome_size() { return GenomeSize; }
int ref_num_chromosomes() { return iChromsomeNum; }
int64_t ref_chromosome(int i, char* name, int cap)
{
	snprintf(name, cap, "%s", ChromosomeVec[i].name);
	return ChromosomeVec[i].len;
}
const char* ref_sequence() { return RefSequence; }
uint64_t ref_primary() { return Refbwt->primary; }
void ref_L2(uint64_t* out5) { for (int i = 0; i < 5; i++) out5[i] = Refbwt->L2[i]; }

void ref_2occ4(uint64_t k, uint64_t l, uint64_t* cntk, uint64_t* cntl) { bwt_2occ4(Refbwt, k, l, cntk, cntl); }
uint64_t ref_sa(uint64_t k) { return bwt_sa(k); }

// BWT_Search (src/bwt_search.cpp:139-182). Returns freq; locs holds min(freq,cap) hits in SA order.
int ref_bwt_search(const uint8_t* enc, int start, int stop, int* len, uint64_t* locs, int cap)
{
	bwtSearchResult_t r = BWT_Search((uint8_t*)enc, start, stop);
	*len = r.len;
	for (int j = 0; j < r.freq && j < cap; j++) locs[j] = r.LocArr[j];
	if (r.LocArr) delete[] r.LocArr;
	return r.freq;
}

// IdentifySeedPairs (src/AlignmentCandidates.cpp:181-215). Returns number of seeds.
int ref_identify_seed_pairs(int rlen, const uint8_t* enc, int32_t* rpos, int64_t* gpos, int32_t* len, int cap)
{
	vector<SeedPair_t> v = IdentifySeedPairs(rlen, (uint8_t*)enc);
	for (int i = 0; i < (int)v.size() && i < cap; i++) { rpos[i] = v[i].rPos; gpos[i] = v[i].gPos; len[i] = v[i].rLen; }
	return (int)v.size();
}

// IdentifySeedPairs + GenerateAlignmentCandidate (src/AlignmentCandidates.cpp:241-288), flattened.
// Returns number of candidates; *nseeds_total gets the number of seeds written.
int ref_candidates(int rlen, const uint8_t* enc, int32_t* c_score, int64_t* c_posdiff, int32_t* c_nseeds, int cap_c,
                   int32_t* s_rpos, int64_t* s_gpos, int32_t* s_len, int cap_s, int* nseeds_total)
{
	vector<SeedPair_t> v = IdentifySeedPairs(rlen, (uint8_t*)enc);
	vector<AlignmentCandidate_t> a = GenerateAlignmentCandidate(rlen, v);
	int ns = 0;
	for (int i = 0; i < (int)a.size(); i++) {
		if (i < cap_c) { c_score[i] = a[i].Score; c_posdiff[i] = a[i].PosDiff; c_nseeds[i] = (int)a[i].SeedVec.size(); }
		for (size_t j = 0; j < a[i].SeedVec.size(); j++, ns++)
			if (ns < cap_s) { s_rpos[ns] = a[i].SeedVec[j].rPos; s_gpos[ns] = a[i].SeedVec[j].gPos; s_len[ns] = a[i].SeedVec[j].rLen; }
	}
	*nseeds_total = ns;
	return (int)a.size();
}

// GenerateLongestSimplePairsFromFragmentPair (src/KmerAnalysis.cpp:134-166). out = {rPos, gPos, rLen}.
void ref_kmer_pair(int len1, const char* f1, int len2, const char* f2, int64_t* out3)
{
	SeedPair_t s = GenerateLongestSimplePairsFromFragmentPair(len1, (char*)f1, len2, (char*)f2);
	out3[2] = s.rLen;
	out3[0] = s.rLen > 0 ? s.rPos : 0; out3[1] = s.rLen > 0 ? s.gPos : 0; // rPos/gPos are unset when nothing was found
}

// nw_alignment (src/nw_alignment.cpp:18-82). Returns aligned length; o1/o2 are NUL-terminated.
int ref_nw(int m, const char* s1, int n, const char* s2, char* o1, char* o2, int cap)
{
	string a(s1, m), b(s2, n);
	nw_alignment(m, a, n, b);
	snprintf(o1, cap, "%s", a.c_str()); snprintf(o2, cap, "%s", b.c_str());
	return (int)a.length();
}

// IdentifyBestGappedPartition (src/AlignmentCandidates.cpp:385-467). out = {p, left_ext, right_ext}.
void ref_gapped_partition(const char* seq, int rGaps, int l_rpos, int l_rlen, int64_t l_gpos, int l_glen,
                          int r_rpos, int64_t r_gpos, int* out3)
{
	SeedPair_t L, R; memset(&L, 0, sizeof L); memset(&R, 0, sizeof R);
	L.rPos = l_rpos; L.rLen = l_rlen; L.gPos = l_gpos; L.gLen = l_glen;
	R.rPos = r_rpos; R.gPos = r_gpos;
	GappedExtension_t g = IdentifyBestGappedPartition((char*)seq, rGaps, L, R);
	out3[0] = g.p; out3[1] = g.left_ext; out3[2] = g.right_ext;
}

// ---- per-read driver: the body of ReadMapping() (src/Mapping.cpp:600-643) -------------------------

static void make_read(ReadItem_t& r, const char* name, const char* seq, int rlen, const char* qual)
{
	memset(&r, 0, sizeof r); // the F1 canonicalisation: zero sub_score/mis_num/mapq
	r.rlen = rlen;
	r.header = new char[strlen(name) + 1]; strcpy(r.header, name);
	r.seq = new char[rlen + 1]; memcpy(r.seq, seq, rlen); r.seq[rlen] = '\0';
	r.qual = new char[rlen + 1]; memcpy(r.qual, qual, rlen); r.qual[rlen] = '\0';
}
static void encode_read(ReadItem_t& r)
{
	r.EncodeSeq = new uint8_t[r.rlen];
	for (int i = 0; i < r.rlen; i++) r.EncodeSeq[i] = nst_nt4_table[(int)r.seq[i]];
}
static void free_read(ReadItem_t& r)
{
	delete[] r.header; delete[] r.seq; delete[] r.qual; delete[] r.EncodeSeq; delete[] r.AlnReportArr;
}

struct RefSJ { map<pair<int64_t, int64_t>, SpliceJunction_t> m; };
void* ref_sj_new() { return new RefSJ; }
void ref_sj_free(void* p) { delete (RefSJ*)p; }
int ref_sj_size(void* p) { return (int)((RefSJ*)p)->m.size(); }
// dump in map order: g1, g2 (absolute), type, count
void ref_sj_dump(void* p, int64_t* g1, int64_t* g2, int32_t* type, int32_t* count)
{
	int i = 0;
	for (auto& kv : ((RefSJ*)p)->m) { g1[i] = kv.first.first; g2[i] = kv.first.second; type[i] = kv.second.type; count[i] = kv.second.iCount; i++; }
}

static int emit(vector<string>& v, char* out, int cap)
{
	int n = 0;
	for (auto& s : v) { if (n + (int)s.size() + 1 >= cap) return -1; memcpy(out + n, s.c_str(), s.size()); n += (int)s.size(); out[n++] = '\n'; }
	out[n] = '\0';
	return n;
}

// Single-end read: src/Mapping.cpp:627-639, :643. seq/qual as in the FASTQ record.
int ref_map_single(const char* name, const char* seq, int rlen, const char* qual, void* sj, char* out, int cap)
{
	ReadItem_t r; make_read(r, name, seq, rlen, qual); encode_read(r);
	vector<SeedPair_t> sv = IdentifySeedPairs(r.rlen, r.EncodeSeq);
	vector<AlignmentCandidate_t> av = GenerateAlignmentCandidate(r.rlen, sv);
	RemoveRedundantCandidates(av);
	GenMappingReport(true, r, av);
	SetSingleAlignmentFlag(r); EvaluateMAPQ(r);
	if (sj && (r.mapq == 50 || (bFindAllJunction && r.score > 0))) UpdateLocalSJMap(av[r.iBestAlnCanIdx], ((RefSJ*)sj)->m);
	int u = 0, un = 0; vector<string> lines;
	OutputSingledAlignments(r, u, un, lines);
	int n = emit(lines, out, cap);
	free_read(r);
	return n;
}

// Pair: src/Mapping.cpp:600-623, :642. Mate 2 is given as in the FASTQ record; it is reverse-complemented
// (qualities reversed) here exactly as GetNextChunk does at load (src/GetData.cpp:157-168).
int ref_map_pair(const char* name1, const char* seq1, int rlen1, const char* qual1,
                 const char* name2, const char* seq2, int rlen2, const char* qual2, void* sj, char* out, int cap)
{
	ReadItem_t r1, r2; make_read(r1, name1, seq1, rlen1, qual1); make_read(r2, name2, seq2, rlen2, qual2);
	{
		char* rseq = new char[rlen2]; GetComplementarySeq(rlen2, r2.seq, rseq);
		copy(rseq, rseq + rlen2, r2.seq); delete[] rseq;
		string rq = r2.qual; reverse(rq.begin(), rq.end()); copy(rq.c_str(), rq.c_str() + rlen2, r2.qual);
	}
	encode_read(r1); encode_read(r2);
	vector<SeedPair_t> s1 = IdentifySeedPairs(r1.rlen, r1.EncodeSeq);
	vector<AlignmentCandidate_t> a1 = GenerateAlignmentCandidate(r1.rlen, s1);
	vector<SeedPair_t> s2 = IdentifySeedPairs(r2.rlen, r2.EncodeSeq);
	vector<AlignmentCandidate_t> a2 = GenerateAlignmentCandidate(r2.rlen, s2);
	if (CheckPairedAlignmentCandidates(a1, a2)) RemoveUnMatedAlignmentCandidates(a1, a2);
	RemoveRedundantCandidates(a1); RemoveRedundantCandidates(a2);
	GenMappingReport(true, r1, a1);
	GenMappingReport(false, r2, a2);
	CheckPairedFinalAlignments(r1, r2);
	SetPairedAlignmentFlag(r1, r2);
	EvaluateMAPQ(r1); EvaluateMAPQ(r2);
	if (sj) {
		if (r1.mapq == 50 || (bFindAllJunction && r1.score > 0)) UpdateLocalSJMap(a1[r1.iBestAlnCanIdx], ((RefSJ*)sj)->m);
		if (r2.mapq == 50 || (bFindAllJunction && r2.score > 0)) UpdateLocalSJMap(a2[r2.iBestAlnCanIdx], ((RefSJ*)sj)->m);
	}
	int u = 0, un = 0, p = 0; vector<string> lines;
	OutputPairedAlignments(r1, r2, u, un, p, lines);
	int n = emit(lines, out, cap);
	free_read(r1); free_read(r2);
	return n;
}

// ---- hot-path-only timing (bench.py cpu_baseline_hotpath) -----------------------------------------------
// The per-read loop body of ReadMapping() (src/Mapping.cpp:600-639) over reads that are ALREADY parsed and
// encoded in memory, on `threads` pthreads, without OutputPaired/SingledAlignments and without any file IO:
// the like-for-like CPU denominator of the GPU path's reads/s (the stock binary's wall clock also pays for
// FASTQ parsing, SAM text and fwrite).  bases = ASCII as ReadItem_t.seq holds them (mate 2 already flipped,
// src/GetData.cpp:157-168), offsets[n+1].  Returns seconds spent in the parallel region.
} // extern "C"
#include <pthread.h>
#include <time.h>
namespace {
struct HotArgs { const char* bases; const int64_t* off; int n; int paired; int tid, nth; long mapped; };
static void hot_one(ReadItem_t& r, const char* bases, const int64_t* off, int i)
{
	memset(&r, 0, sizeof r);
	r.rlen = (int)(off[i + 1] - off[i]);
	r.seq = (char*)bases + off[i];
	r.EncodeSeq = new uint8_t[r.rlen];
	for (int k = 0; k < r.rlen; k++) r.EncodeSeq[k] = nst_nt4_table[(int)(unsigned char)r.seq[k]];
}
static void* hot_worker(void* p)
{
	HotArgs* a = (HotArgs*)p;
	map<pair<int64_t, int64_t>, SpliceJunction_t> sj;
	const int unit = a->paired ? 2 : 1, units = a->n / unit;
	const int lo = (int)((int64_t)units * a->tid / a->nth), hi = (int)((int64_t)units * (a->tid + 1) / a->nth);
	for (int u = lo; u < hi; u++) {
		if (a->paired) {
			ReadItem_t r1, r2; hot_one(r1, a->bases, a->off, 2 * u); hot_one(r2, a->bases, a->off, 2 * u + 1);
			vector<SeedPair_t> s1 = IdentifySeedPairs(r1.rlen, r1.EncodeSeq);
			vector<AlignmentCandidate_t> a1 = GenerateAlignmentCandidate(r1.rlen, s1);
			vector<SeedPair_t> s2 = IdentifySeedPairs(r2.rlen, r2.EncodeSeq);
			vector<AlignmentCandidate_t> a2 = GenerateAlignmentCandidate(r2.rlen, s2);
			if (CheckPairedAlignmentCandidates(a1, a2)) RemoveUnMatedAlignmentCandidates(a1, a2);
			RemoveRedundantCandidates(a1); RemoveRedundantCandidates(a2);
			GenMappingReport(true, r1, a1); GenMappingReport(false, r2, a2);
			CheckPairedFinalAlignments(r1, r2);
			SetPairedAlignmentFlag(r1, r2);
			EvaluateMAPQ(r1); EvaluateMAPQ(r2);
			if (r1.mapq == 50 || (bFindAllJunction && r1.score > 0)) UpdateLocalSJMap(a1[r1.iBestAlnCanIdx], sj);
			if (r2.mapq == 50 || (bFindAllJunction && r2.score > 0)) UpdateLocalSJMap(a2[r2.iBestAlnCanIdx], sj);
			a->mapped += (r1.score > 0) + (r2.score > 0);
			delete[] r1.EncodeSeq; delete[] r2.EncodeSeq; delete[] r1.AlnReportArr; delete[] r2.AlnReportArr;
		} else {
			ReadItem_t r; hot_one(r, a->bases, a->off, u);
			vector<SeedPair_t> sv = IdentifySeedPairs(r.rlen, r.EncodeSeq);
			vector<AlignmentCandidate_t> av = GenerateAlignmentCandidate(r.rlen, sv);
			RemoveRedundantCandidates(av);
			GenMappingReport(true, r, av);
			SetSingleAlignmentFlag(r); EvaluateMAPQ(r);
			if (r.mapq == 50 || (bFindAllJunction && r.score > 0)) UpdateLocalSJMap(av[r.iBestAlnCanIdx], sj);
			a->mapped += r.score > 0;
			delete[] r.EncodeSeq; delete[] r.AlnReportArr;
		}
	}
	return 0;
}
}
extern "C" double ref_hotpath(const char* bases, const int64_t* offsets, int n_reads, int paired, int threads, long* mapped)
{
	if (threads < 1) threads = 1;
	vector<HotArgs> args(threads);
	vector<pthread_t> th(threads);
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < threads; t++) {
		args[t] = HotArgs{bases, offsets, n_reads, paired, t, threads, 0};
		pthread_create(&th[t], NULL, hot_worker, &args[t]);
	}
	long m = 0;
	for (int t = 0; t < threads; t++) { pthread_join(th[t], NULL); m += args[t].mapped; }
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (mapped) *mapped = m;
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
