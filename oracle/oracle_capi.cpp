// ORACLE — TEST INFRASTRUCTURE ONLY (see dp_oracle.hpp).
//
// C ABI around the CPU restatement so that tests/ and bench.py can drive it through
// ctypes.  Mirrors the reference entry points it restates:
//   oracle_fullyGlobalAlignment  <- unicycler/src/global_align.cpp:19-90
//   oracle_pathAlignment         <- unicycler/src/path_align.cpp:18-92
//   oracle_chainAlignment        <- unicycler/src/semi_global_align.cpp:294-311 (the
//                                   bandedChainAlignment + ScoredAlignment step, given a seed chain)
#include <cstdlib>
#include <cstring>
#include <string>

#include "dp_oracle.hpp"

using namespace orc;

static char* dupString(const std::string& s) {
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.c_str(), s.size() + 1);
    return p;
}

static thread_local long long g_lastCells = 0;
static thread_local long long g_lastGrids = 0;

extern "C" {

void oracle_free(char* p) { free(p); }

// DP cells (reference definition) and sub-DP count of the last call on this thread.
long long oracle_lastCells() { return g_lastCells; }
long long oracle_lastGrids() { return g_lastGrids; }

static char* globalLike(const char* s1, const char* s2, int match, int mismatch, int gapOpen, int gapExtend,
                        bool useBanding, int bandSize, bool path) {
    std::string a(s1), b(s2);
    std::vector<uint8_t> H = toDna5(a), V = toDna5(b);
    Score sc{match, mismatch, gapExtend, gapOpen};
    FreeEnds fe{false, false, false, path};  // AlignConfig<false,false,true,false> -> free last column
    long lower = -bandSize, upper = bandSize;
    long diff = (long)b.size() - (long)a.size();
    if (!path) {
        if (diff > 0) lower -= diff;
        else if (diff < 0) upper -= diff;
    } else {
        if (diff < 0) upper -= diff;
    }
    Trace tr;
    int score = 0;
    CellCounter cc;
    bool ok = globalAlignmentTrace(H, V, sc, fe, useBanding, lower, upper, tr, score, &cc);
    g_lastCells = cc.cells;
    g_lastGrids = cc.grids;
    if (!ok) return dupString("");
    if (path && score < -1000000) return dupString("");
    std::string rowH, rowV;
    traceToRows(tr, H, V, rowH, rowV);
    return dupString(scoredAlignmentString(rowH, rowV, "s1", "s2", 0, true, true, !path, sc));
}

char* oracle_fullyGlobalAlignment(const char* s1, const char* s2, int match, int mismatch, int gapOpen,
                                  int gapExtend, bool useBanding, int bandSize) {
    return globalLike(s1, s2, match, mismatch, gapOpen, gapExtend, useBanding, bandSize, false);
}

char* oracle_pathAlignment(const char* s1, const char* s2, int match, int mismatch, int gapOpen, int gapExtend,
                           bool useBanding, int bandSize) {
    return globalLike(s1, s2, match, mismatch, gapOpen, gapExtend, useBanding, bandSize, true);
}

// seeds: nSeeds x 6 longs (beginH, beginV, endH, endV, lowerDiag, upperDiag), chain order.
// readName carries the strand suffix ('+' / '-') like signedReadName in the reference.
char* oracle_chainAlignment(const char* readSeq, const char* trimmedRefSeq, const long* seeds, int nSeeds,
                            int match, int mismatch, int gapOpen, int gapExtend, int bandSize,
                            const char* readName, const char* refName, int refOffset) {
    std::string a(readSeq), b(trimmedRefSeq);
    std::vector<uint8_t> H = toDna5(a), V = toDna5(b);
    Score sc{match, mismatch, gapExtend, gapOpen};
    FreeEnds fe{true, true, true, true};
    std::vector<Seed> chain((size_t)nSeeds);
    for (int i = 0; i < nSeeds; ++i)
        chain[(size_t)i] = Seed{seeds[6 * i], seeds[6 * i + 1], seeds[6 * i + 2],
                                seeds[6 * i + 3], seeds[6 * i + 4], seeds[6 * i + 5]};
    Trace tr;
    bool empty = true;
    int score = 0;
    CellCounter cc;
    bool ok;
    try {
        ok = bandedChainAlignmentTrace(H, V, chain, sc, fe, (unsigned)bandSize, tr, empty, score, &cc);
    } catch (std::exception& e) {
        g_lastCells = cc.cells;
        g_lastGrids = cc.grids;
        return dupString(std::string("!ERROR:") + e.what());
    }
    g_lastCells = cc.cells;
    g_lastGrids = cc.grids;
    if (!ok) return dupString("");
    std::string rowH, rowV;
    if (empty) {  // Align rows left untouched: ungapped full sources
        static const char* A = "ACGTN";
        for (uint8_t c : H) rowH.push_back(A[c]);
        for (uint8_t c : V) rowV.push_back(A[c]);
    } else
        traceToRows(tr, H, V, rowH, rowV);
    return dupString(scoredAlignmentString(rowH, rowV, readName, refName, refOffset, false, false, false, sc));
}

}  // extern "C"
