// Host seeding stage of semiGlobalAlignment (unicycler/src/semi_global_align.cpp:156-291 and helpers):
// common k-mers -> kd-tree line tracing -> seed merge -> sparse global chaining.  0.6 % of the
// reference's runtime and floating-point / container-order sensitive, so it stays on the host and keeps
// the reference's container types and iteration orders (SURVEY.md §7 hard part 4).
#pragma once
#include <atomic>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "host_align.hpp"

namespace ub200 {

// All k-mer start positions of a read (KmerPositions, include/kmers.h:26 / src/kmers.cpp:51-65).  The
// reference keys an unordered_map by the k-mer STRING; only the per-k-mer position lists (ascending) are ever
// observed, so k-mers made of upper-case ACGT only are keyed by their 2-bit code in an open-addressing table
// and every other k-mer (N, lower case, ...) keeps the literal string key.
struct KmerPosMap {
    int k = 0;
    uint32_t mask = 0;
    std::vector<int32_t> head;   // slot -> first position, -1 = empty
    std::vector<uint32_t> key;   // slot -> 2-bit code
    std::vector<int32_t> next;   // position -> next position with the same k-mer
    std::unordered_map<std::string, std::vector<int> > other;
};

struct SensitivityParams {  // include/settings.h:17-42
    int kSize, bandSize, minLineTraceCount, maxLineTraceCount;
};
SensitivityParams sensitivityParams(int level);

// KmerPositions::addPositions (src/kmers.cpp:51-65)
void buildKmerPositions(const std::string& sequence, int kSize, KmerPosMap& out);

struct RangeSeeds {
    std::vector<std::vector<ChainSeed> > chains;  // one per accepted point set, in the reference's order
    std::string console;                          // verbosity > 2 text
};

// Everything alignReadToReferenceRange does before bandedChainAlignment (semi_global_align.cpp:197-291).
// The returned chains stop at the first point set whose chain is empty or whose gap area exceeds
// MAX_BANDED_ALIGNMENT_GAP_AREA, like the early returns at :286-291.
// joinedXY != nullptr: the range's common k-mer points (x = read position, y = window position, the reference's
// order) were already found by the device join (kmerjoin.hpp); readKmers is not used then.
void seedRange(const std::string& readSeq, const KmerPosMap& readKmers, const std::string& trimmedRefSeq,
               const SensitivityParams& sp, int verbosity, const std::string& refName, int refStart, int refEnd,
               RangeSeeds& out, const int32_t* joinedXY = nullptr, size_t nJoined = 0);

// The host join alone (x, y pairs in the reference's order): the per-read entry point's join, and the checker of
// the device join in tests/.
void commonKmerPoints(const std::string& readSeq, const std::string& trimmedRefSeq, int kSize, std::vector<int32_t>& xy);

std::string reverseComplement(const std::string& s);  // src/string_functions.cpp:52-79

}  // namespace ub200
