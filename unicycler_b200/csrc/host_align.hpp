// Host side of the alignment path: chain planning (grid geometry), traceback gluing and
// the ScoredAlignment result formatter.  Pure C++ (no CUDA, no PyTorch).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "engine.hpp"

namespace ub200 {

struct ChainSeed {  // Seed<Simple>: begin/end positions and diagonal bounds (seqan/seeds/seeds_seed_simple.h)
    long beginH, beginV, endH, endV, lowerDiag, upperDiag;
};

struct Scoring {
    int match, mismatch, gapOpen, gapExtend;
};

// basic/alphabet_residue_tabs.h:113-140 (char -> Dna5 code)
void toDna5(const char* s, size_t n, std::vector<uint8_t>& out);

// Plans the sub-DP sequence of bandedChainAlignment(align, chain, score, AlignConfig<true,true,true,true>, b)
// (seqan/seeds/banded_chain_alignment_impl.h:1212-1296 with :737-1177).  Returns false when the chain is
// empty (the reference returns MinValue without touching the alignment).
bool planChain(const std::vector<ChainSeed>& chain, long lenH, long lenV, long bandExtension,
               std::vector<GridDesc>& grids, long long* colTabCount = nullptr);

// One-grid plans for globalAlignment(align, score, AlignConfig<...>[, lo, up]).
// Returns false if _isValidDPSettings would fail (seqan/align/dp_algorithm_impl.h:117-157).
bool planGlobal(long lenH, long lenV, bool banded, long lo, long up, bool freeFirstRow, bool freeFirstCol,
                bool freeLastRow, bool freeLastCol, std::vector<GridDesc>& grids);

// Replays _glueTracebacks over the per-grid local trace sets (seqan/seeds/banded_chain_alignment_traceback.h:96-203)
// and returns traceSet[0] (seeds/banded_chain_alignment.h:207).  empty=true mirrors "empty(traceSet)".
void glueChain(const std::vector<GridDesc>& grids, const JobResult& res, std::vector<Seg>& trace, bool& empty);

struct AlignmentRecord {  // the fields of ScoredAlignment (unicycler/src/scoredalignment.cpp)
    int readStart = -1, readEnd = 0, refStart = -1, refEnd = 0, rawScore = 0;
    double scaledScore = 0.0;
    std::string cigar;
    bool emptyAlignment = true;
};

// ScoredAlignment constructor on a trace (scoredalignment.cpp:16-136).  H is the read row, V the reference row.
void scoreAlignment(const std::vector<Seg>& trace, bool traceEmpty, const uint8_t* H, long lenH, const uint8_t* V,
                    long lenV, int refOffset, bool startImmediately, bool goToEndSeq1, bool goToEndSeq2,
                    const Scoring& sc, AlignmentRecord& rec);

// getFullString (scoredalignment.cpp:139-156)
std::string fullString(const AlignmentRecord& rec, const std::string& readName, const std::string& refName,
                       long long milliseconds);

}  // namespace ub200
