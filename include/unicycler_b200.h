/* unicycler_b200 — C ABI of the B200-native replacement for the long-read alignment hot path of
 * Unicycler v0.5.1 (libunicycler_b200.so).
 *
 * Part 1 is byte-for-byte the extern "C" surface that unicycler/cpp_wrappers.py binds through ctypes
 * for this path; a maintainer drops the library in as unicycler/cpp_functions.so.
 * Every returned char* is malloc()ed by the library and released by the caller with freeCString().
 * Field 9 (index 8) of a result string is wall-clock milliseconds and therefore not reproducible.
 *
 * Part 2 is additive: batch entry points (so that one call fills the GPU) and introspection used by
 * the benchmark.  Plain pointers and sizes only — no C++ or torch types cross this boundary.
 *
 * All citations are relative to /root/reference/unicycler/.
 */
#ifndef UNICYCLER_B200_H
#define UNICYCLER_B200_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Part 1 — the reference's own C ABI (hot-path symbols, implemented natively on the GPU)
 * ---------------------------------------------------------------------------------------------- */

/* replaces include/semi_global_align.h:54-61 (src/semi_global_align.cpp:24-153); cpp_wrappers.py:33-55.
 * refSeqs is the handle returned by newRefSeqs().  lowScoreThreshold / returnBad are unused there too. */
char* semiGlobalAlignment(char* readName, char* readSeq, int verbosity, char* minimapAlignmentsStr,
                          void* refSeqs, int matchScore, int mismatchScore, int gapOpenScore,
                          int gapExtensionScore, double lowScoreThreshold, bool returnBad,
                          int sensitivityLevel);

/* replaces include/global_align.h:23-27 (src/global_align.cpp:19-90); cpp_wrappers.py:80-95 */
char* fullyGlobalAlignment(char* s1, char* s2, int matchScore, int mismatchScore, int gapOpenScore,
                           int gapExtensionScore, bool useBanding, int bandSize);

/* replaces include/path_align.h:23-27 (src/path_align.cpp:18-92); cpp_wrappers.py:102-117 */
char* pathAlignment(char* s1, char* s2, int matchScore, int mismatchScore, int gapOpenScore,
                    int gapExtensionScore, bool useBanding, int bandSize);

/* replaces include/random_alignments.h:26-27 (src/random_alignments.cpp:30-52); cpp_wrappers.py:161-175.
 * Returns "mean,stddev" formatted with %f.  Set UNICYCLER_B200_SEED to make the RNG reproducible
 * (the reference seeds std::mt19937 from std::random_device). */
char* getRandomSequenceAlignmentScores(int seqLength, int n, int matchScore, int mismatchScore,
                                       int gapOpenScore, int gapExtensionScore);

/* replace include/ref_seqs.h:21-25 (src/ref_seqs.cpp:14-24); cpp_wrappers.py:138-156 */
void* newRefSeqs(void);
void addRefSeq(void* refSeqs, char* name, char* sequence);
void deleteRefSeqs(void* refSeqs);

/* replaces include/string_functions.h:28 (src/string_functions.cpp:27-29); cpp_wrappers.py:123-133 */
void freeCString(char* p);

/* replaces include/semi_global_align_exhaustive.h (src/semi_global_align_exhaustive.cpp:18-67); cpp_wrappers.py:61-74.
 * Unbanded alignment with all four end gaps free; result string like fullyGlobalAlignment, "" on failure. */
char* semiGlobalAlignmentExhaustive(char* s1, char* s2, int matchScore, int mismatchScore, int gapOpenScore,
                                    int gapExtensionScore);

/* replace include/start_end_align.h:22-24 (src/start_end_align.cpp:19-101); cpp_wrappers.py:325-357.
 * Position in s2 where the alignment of s1 to the start / end of s2 ends / starts; -1 on failure. */
int startAlignment(char* s1, char* s2, int matchScore, int mismatchScore, int gapOpenScore, int gapExtensionScore);
int endAlignment(char* s1, char* s2, int matchScore, int mismatchScore, int gapOpenScore, int gapExtensionScore);

/* replaces include/overlap_align.h (src/overlap_align.cpp:17-81); cpp_wrappers.py:180-200.
 * Overlap of the end of s1 with the start of s2: "overlap1,overlap2", "-1,-1" on failure. */
char* overlapAlignment(char* s1, char* s2, int matchScore, int mismatchScore, int gapOpenScore,
                       int gapExtensionScore, int guessOverlap);

/* The remaining symbols cpp_wrappers.py dereferences at import (multipleSequenceAlignment, minimapAlignReads,
 * minimapAlignReadsWithSettings, miniasmAssembly, simulateDepths, getRandomSequenceAlignmentErrorRates)
 * are exported as forwarders: they dlopen() the library named by UNICYCLER_B200_FORWARD_LIB (the stock
 * cpp_functions.so) and abort with a clear message if it is not set.  They are outside the hot path. */

/* ------------------------------------------------------------------------------------------------
 * Part 2 — additive batch / introspection API
 * ---------------------------------------------------------------------------------------------- */

/* n pairwise alignments in one launch.  mode 0 = fullyGlobalAlignment, 1 = pathAlignment.
 * results[i] receives a malloc()ed string (release each with freeCString).  Returns 0 on success. */
int ub200_globalAlignmentBatch(int n, const char* const* s1, const char* const* s2, int mode,
                               int matchScore, int mismatchScore, int gapOpenScore, int gapExtensionScore,
                               bool useBanding, int bandSize, char** results);

/* The bandedChainAlignment + ScoredAlignment step of alignReadToReferenceRange
 * (src/semi_global_align.cpp:294-311) for a given seed chain: seeds = nSeeds x 6 int64
 * (beginH, beginV, endH, endV, lowerDiag, upperDiag).  readName carries the strand suffix. */
char* ub200_chainAlignment(const char* readSeq, const char* trimmedRefSeq, const int64_t* seeds, int nSeeds,
                           int matchScore, int mismatchScore, int gapOpenScore, int gapExtensionScore,
                           int bandSize, const char* readName, const char* refName, int refOffset);

/* n chain alignments in one launch (same arguments as ub200_chainAlignment, arrays of length n;
 * seedOffsets has n+1 entries into seeds). */
int ub200_chainAlignmentBatch(int n, const char* const* readSeqs, const char* const* refSeqs,
                              const int64_t* seeds, const int64_t* seedOffsets, int matchScore, int mismatchScore,
                              int gapOpenScore, int gapExtensionScore, int bandSize, const char* const* readNames,
                              const char* const* refNames, const int* refOffsets, char** results);

/* semiGlobalAlignment for n reads in one launch (all reads against the same refSeqs handle). */
int ub200_semiGlobalAlignmentBatch(int n, const char* const* readNames, const char* const* readSeqs,
                                   const char* const* minimapAlignmentsStrs, void* refSeqs, int matchScore,
                                   int mismatchScore, int gapOpenScore, int gapExtensionScore,
                                   int sensitivityLevel, char** results);

/* Host seeding stage only (src/semi_global_align.cpp:197-291): returns, for the given read strand and
 * trimmed reference window, "nChains;" followed per chain by "nSeeds:bH,bV,eH,eV,lo,up|...;" — used by the
 * tests to pin the seeding against golden seed chains. */
char* ub200_seedChains(const char* readSeq, const char* trimmedRefSeq, int sensitivityLevel);

/* Common k-mer points of one read strand and one window [refStart, refStart + refLen) of `ref`
 * (src/semi_global_align.cpp:197-207 over KmerPositions, src/kmers.cpp:51-65): (x = read position, y = window
 * position) pairs in the reference's order.  where = 0: the host join used by semiGlobalAlignment; where = 1: the
 * device join used by ub200_semiGlobalAlignmentBatch.  Writes at most cap pairs into xy and
 * returns the number of points (-1: no usable device). */
int64_t ub200_commonKmers(const char* readSeq, const char* ref, int refStart, int refLen, int k, int where,
                          int32_t* xy, int64_t cap);
/* Counters of the device k-mer join inside the last ub200_semiGlobalAlignmentBatch call. */
void ub200_lastJoinStats(double* kernelMs, int64_t* launches, int64_t* points, int64_t* h2dBytes, int64_t* d2hBytes);

/* Counters of the last engine run on this process: DP cells (reference definition), kernel
 * milliseconds (CUDA events), launches, H2D / D2H milliseconds. */
void ub200_lastStats(int64_t* cells, double* kernelMs, int64_t* launches, double* h2dMs, double* d2hMs);

/* Device-resident benchmark hooks for the chain path: prepare uploads and plans n chain jobs once,
 * run launches the kernel on the resident inputs (returns kernel ms), finish fetches and formats. */
int ub200_chainBenchPrepare(int n, const char* const* readSeqs, const char* const* refSeqs,
                            const int64_t* seeds, const int64_t* seedOffsets, int matchScore, int mismatchScore,
                            int gapOpenScore, int gapExtensionScore, int bandSize);
double ub200_chainBenchRun(void);
int ub200_chainBenchFinish(char** results);
/* `steps` back-to-back launches on the resident inputs; returns their total CUDA-event time in ms. */
double ub200_chainBenchRunSteps(int steps);
/* Bytes moved by the last run (host->device, device->host), trace bytes written per launch, resident CTAs. */
void ub200_lastTransferBytes(int64_t* h2d, int64_t* d2h, int64_t* traceBytes, int* ctas);
/* Reference DP-cell count and sub-DP count of one banded-chain alignment; planner only. */
int64_t ub200_chainCells(int readLen, int refLen, const int64_t* seeds, int nSeeds, int bandSize, int* nGrids);

/* Planner introspection: 10 int32 per sub-DP (kind, nH, nV, banded, lo, up, h0, v0, hNext, vNext) into out (cap grids);
 * returns the number of sub-DPs of the banded-chain alignment, -1 if the chain is unsupported. */
int ub200_chainPlan(int readLen, int refLen, const int64_t* seeds, int nSeeds, int bandSize, int32_t* out, int cap);

/* The i.i.d. ACGT pairs that getRandomSequenceAlignmentScores(seqLength, n, ...) aligns when its generator is seeded
 * with `seed` (std::mt19937 + std::uniform_int_distribution<int>(0, 3), s1 then s2 of each pair,
 * src/random_alignments.cpp:33-40, 167-185).  s1[i] / s2[i] receive malloc()ed strings.  Lets a test feed the very
 * same pairs to the reference's fullyGlobalAlignment. */
int ub200_calibrationPairs(int seqLength, int n, unsigned seed, char** s1, char** s2);

/* the tallies Alignment.tally_up_score_and_errors (unicycler/alignment.py:142-216) derives from a
 * CIGAR by walking it base by base in Python.  readSeq = the read as aligned (reverse-complemented for '-'), refSeq =
 * the whole reference.  Returns "matches,mismatches,insertions,deletions,rawScore,alignmentLength,percentIdentity,
 * scaledScore" (malloc()ed; doubles printed with 17 significant digits), "" if only soft clips remain. */
char* ub200_alignmentTallies(const char* readSeq, const char* refSeq, int readStartPos, int refStartPos,
                             const char* cigar, int matchScore, int mismatchScore, int gapOpenScore,
                             int gapExtensionScore);

/* Coalescer counters: device batches run so far, and ABI requests they served (requests / batches > 1 means that
 * concurrent per-read calls were merged into shared launches). */
void ub200_coalescerStats(int64_t* batches, int64_t* requests);

/* Selects the CUDA device for this process' engine (before first use, never while calls are in flight).
 * Returns 0 on success, -1 when a batch is running. */
int ub200_setDevice(int device);
/* Integer-pipe microbenchmark (dependent-free IADD3/VIMNMX mix on all SMs): returns int32 ops/s. */
double ub200_intPeakOpsPerSec(void);
const char* ub200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* UNICYCLER_B200_H */
