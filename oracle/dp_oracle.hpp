// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// build, link or call anything in this directory.
//
// CPU restatement (scalar C++, no SIMD, no threads) of the pairwise-DP engine that
// Unicycler v0.5.1 uses through its vendored, patched SeqAn 2.3.1.  It is written as a
// LITERAL emulation of the reference control flow (column descriptors, matrix
// navigators, scouts, traceback coordinator) so that tie-breaking, band geometry and
// the banded-chain hand-off are reproduced bit for bit.  All file:line citations are
// relative to /root/reference/unicycler/include/seqan unless stated otherwise.
//
// Parity status: PINNED.  tests/test_oracle_vs_golden.py checks this restatement
// against golden vectors generated from the unmodified reference library
// (oracle/_ref/libunicycler_ref.so, recipe oracle/Makefile.ref) by
// tests/golden/make_golden.py, including the reference's own known-answer tests
// (test/test_cpp_wrappers.py, test/test_semi_global_alignment.py).
#pragma once
#include <climits>
#include <cstdint>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

static const int NEG = INT_MIN / 2;  // align/dp_cell.h:122-124 (DPCellDefaultInfinity)

// align/dp_cell_affine.h:60-66; linear cells only use .s (align/dp_cell_linear.h)
struct Cell {
    int s = NEG, h = NEG, v = NEG;
};

// Score<int, Simple>(match, mismatch, gapExtend, gapOpen) — score/score_simple.h
struct Score {
    int match, mismatch, gapExtend, gapOpen;
    bool affine() const { return gapOpen != gapExtend; }  // align/dp_setup.h:246-262
};

// align/dp_profile.h:146-153 (TraceBitMap_)
enum : uint8_t { T_NONE = 0, T_D = 1, T_H = 2, T_V = 4, T_HO = 8, T_VO = 16, T_MH = 32, T_MV = 64 };

// align/dp_trace_segment.h (TraceSegment_)
struct Seg {
    long hBeg, vBeg, len;
    int dir;  // T_D, T_H or T_V
    long hEnd() const { return dir == T_V ? hBeg : hBeg + len; }  // _getEndHorizontal
    long vEnd() const { return dir == T_H ? vBeg : vBeg + len; }  // _getEndVertical
};
typedef std::vector<Seg> Trace;  // stored end-of-alignment first, like the reference

enum Algo { ALGO_GLOBAL, ALGO_CHAIN };
enum MatLoc { LOC_INITIAL, LOC_INNER, LOC_FINAL };  // seeds/banded_chain_alignment_profile.h:55-60
struct FreeEnds {
    bool firstRow, firstCol, lastRow, lastCol;  // FreeEndGaps_<FirstRow, FirstColumn, LastRow, LastColumn>
};

struct InitCell {  // Triple<unsigned, unsigned, TDPCell>, basic/triple_base.h:142-152
    unsigned i1, i2;
    Cell c;
    bool affine;
    bool cellLess(const Cell& a, const Cell& b) const {
        if (affine) return a.s < b.s && a.h < b.h && a.v < b.v;  // align/dp_cell_affine.h:113-118
        return a.s < b.s;                                         // align/dp_cell_linear.h:109-113
    }
    bool operator<(const InitCell& o) const {
        if (i1 < o.i1) return true;
        if (i1 == o.i1 && i2 < o.i2) return true;
        if (i1 == o.i1 && i2 == o.i2 && cellLess(c, o.c)) return true;
        return false;
    }
};

// seeds/banded_chain_alignment_scout.h:62-79 (DPScoutState_<BandedChainAlignmentScoutState>)
struct ChainState {
    unsigned hNext = 0, vNext = 0;
    std::vector<Cell> hInitCur, vInitCur, hInitNext, vInitNext;
    std::set<InitCell> nextInitCells;
};

struct Seed {  // Seed<Simple>: seeds/seeds_seed_simple.h
    long beginH, beginV, endH, endV, lowerDiag, upperDiag;
};

struct BadScore : std::runtime_error {  // the "RRW" throws, align/dp_algorithm_impl.h:1591-1593
    BadScore() : std::runtime_error("Bad Seqan alignment score") {}
};

// Statistics shared with the benchmark: DP cells as the reference allocates them
// (align/dp_algorithm_impl.h:1547-1560: dimH * dimV of the score/trace matrix).
struct CellCounter {
    long long cells = 0;
    long long grids = 0;
};

// Debug hook (unit tests of the product's geometry helpers): called for every computed cell
// with the navigator state of the literal emulation.
typedef void (*CellHook)(int col, int row, long tpos, long tLeap, int cp, int cl, int ct, int dimV);
extern CellHook g_cellHook;
// Called once per sub-DP: kind (0 initial, 1 inner, 2 final, 3 default-scout global), dims, band, next-grid origin.
typedef void (*GridHook)(int kind, long nH, long nV, int banded, long lo, long up, long hNext, long vNext);
extern GridHook g_gridHook;

// One DP problem (one call of _computeAlignment, align/dp_algorithm_impl.h:1513-1604).
struct DPProblem {
    const uint8_t* H;
    long nH;
    const uint8_t* V;
    long nV;
    Score sc;
    bool banded;
    long lower, upper;
    bool complete;  // CompleteTrace vs SingleTrace
    Algo algo;
    FreeEnds fe;
    MatLoc loc;
    ChainState* st;
};

// Global / path alignment front-ends (align/global_alignment_unbanded.h:242-262,
// align/global_alignment_banded.h:93).  Returns false if SeqAn would throw.
bool globalAlignmentTrace(const std::vector<uint8_t>& H, const std::vector<uint8_t>& V, const Score& sc,
                          const FreeEnds& fe, bool banded, long lower, long upper, Trace& out, int& score,
                          CellCounter* cc = nullptr);

// bandedChainAlignment (seeds/banded_chain_alignment.h:188-210) with AlignConfig<true,true,true,true>.
// Returns false if SeqAn would throw; traceEmpty reports the "empty(traceSet)" early return.
bool bandedChainAlignmentTrace(const std::vector<uint8_t>& H, const std::vector<uint8_t>& V,
                               const std::vector<Seed>& chain, const Score& sc, const FreeEnds& fe,
                               unsigned bandExtension, Trace& out, bool& traceEmpty, int& score,
                               CellCounter* cc = nullptr);

// _adaptTraceSegmentsTo (align/dp_traceback_adaptor.h:60-118) + row streaming.
void traceToRows(const Trace& tr, const std::vector<uint8_t>& H, const std::vector<uint8_t>& V,
                 std::string& rowH, std::string& rowV);

// ScoredAlignment (unicycler/src/scoredalignment.cpp:16-156).  Field 8 (milliseconds) is
// emitted as "0".
std::string scoredAlignmentString(const std::string& rowRead, const std::string& rowRef,
                                  const std::string& readName, const std::string& refName, int refOffset,
                                  bool startImmediately, bool goToEndSeq1, bool goToEndSeq2, const Score& sc,
                                  double* scaledOut = nullptr);

std::vector<uint8_t> toDna5(const std::string& s);  // basic/alphabet_residue_tabs.h:113-140

}  // namespace orc
