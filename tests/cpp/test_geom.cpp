// CPU unit test: the product's closed-form storage geometry (dpgeom.hpp) against the
// oracle's literal navigator emulation (oracle/dp_oracle.cpp debug hook).
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>

#include "../../oracle/dp_oracle.hpp"
#include "../../unicycler_b200/csrc/dpgeom.hpp"

struct HookCell { int col, row; long tpos, tLeap; int cp, cl, ct, dimV; };
static std::vector<HookCell> g_cells;
static void hook(int col, int row, long tpos, long tLeap, int cp, int cl, int ct, int dimV) {
    g_cells.push_back(HookCell{col, row, tpos, tLeap, cp, cl, ct, dimV});
}

int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937 rng(12345);
    orc::g_cellHook = hook;
    long checkedCells = 0, problems = 0, bad = 0;
    for (int it = 0; it < iters; ++it) {
        int nH = 2 + rng() % 40, nV = 2 + rng() % 40;
        int lo = -1 - (int)(rng() % 45), up = 1 + (int)(rng() % 45);
        if (rng() % 4 == 0) { nH = 2 + rng() % 12; nV = 2 + rng() % 12; }
        std::vector<uint8_t> H(nH), V(nV);
        for (auto& c : H) c = rng() % 4;
        for (auto& c : V) c = rng() % 4;
        orc::Score sc{3, -6, -2, -5};
        orc::FreeEnds fe{true, true, true, true};
        orc::Trace tr; int score;
        g_cells.clear();
        bool ok = orc::globalAlignmentTrace(H, V, sc, fe, true, lo, up, tr, score);
        (void)ok;
        if (g_cells.empty()) continue;
        ++problems;
        ub200::GridGeom g = ub200::makeGeom(nH, nV, 1, lo, up);
        // expected column order from the walker
        ub200::BandWalker w; w.init(g);
        ub200::ColInfo ci;
        size_t k = 0;
        bool fail = false;
        while (w.next(ci)) {
            for (int c = 0; c < ci.nCells; ++c, ++k) {
                if (k >= g_cells.size()) { fail = true; break; }
                const HookCell& hc = g_cells[k];
                int ct = (c == 0) ? ub200::CT_FIRST : (c == ci.nCells - 1 ? ub200::CT_LAST : ub200::CT_INNER);
                int row = ci.rowTop + c;
                int cv = ci.cvFirst + c;
                int leap = (ct == ub200::CT_LAST) ? ci.tLeapLast : ci.tLeap;
                // oracle enum order: ColLoc {FULL, TOP, MIDDLE, BOTTOM}, ColProp {INITIAL, INNER, FINAL}
                if (hc.col != ci.j || hc.row != row || hc.cp != ci.cp || hc.cl != ci.cl || hc.ct != ct ||
                    hc.tpos != (long)ci.j * g.dimV + cv || hc.tLeap != leap || hc.dimV != g.dimV)
                    fail = true;
                // closed form
                if (cv != row + ub200::storageOffset(g, ci.j)) fail = true;
                if (row < ub200::colTop(g, ci.j) || row > ub200::colBottom(g, ci.j)) fail = true;
                ++checkedCells;
                if (fail) break;
            }
            if (fail) break;
        }
        if (k != g_cells.size()) fail = true;
        if (fail) {
            ++bad;
            if (bad <= 5) fprintf(stderr, "MISMATCH nH=%d nV=%d lo=%d up=%d (cell %zu of %zu)\n", nH, nV, lo, up, k, g_cells.size());
        }
    }
    // bandLastColumn / bandColumnsFrom (the number of column descriptors a grid needs, planned without walking)
    // against the walker itself, exhaustively over small geometries
    long geoms = 0;
    for (int nH = 1; nH <= 36; ++nH)
        for (int nV = 1; nV <= 36; ++nV)
            for (int lo = -40; lo < 0; ++lo)
                for (int up = 1; up <= 40; ++up) {
                    if (lo <= -nV && up >= nH) continue;   // planned as an unbanded grid
                    const ub200::GridGeom g = ub200::makeGeom(nH, nV, 1, lo, up);
                    ub200::BandWalker w; w.init(g);
                    ub200::ColInfo ci;
                    int prev = -1;
                    bool consecutive = true;
                    while (w.next(ci)) { if (ci.j != prev + 1) consecutive = false; prev = ci.j; }
                    ++geoms;
                    bool ok = consecutive && prev == ub200::bandLastColumn(g);
                    for (int hNext = 0; hNext <= nH && ok; ++hNext)
                        ok = ub200::bandColumnsFrom(g, hNext) == (prev >= hNext ? prev - hNext + 1 : 0);
                    if (!ok) {
                        ++bad;
                        if (bad <= 5) fprintf(stderr, "COLUMN COUNT nH=%d nV=%d lo=%d up=%d: walker ends at %d, closed form %d\n", nH, nV, lo, up, prev, ub200::bandLastColumn(g));
                    }
                }
    printf("problems %ld cells %ld column-count geometries %ld bad %ld\n", problems, checkedCells, geoms, bad);
    return bad ? 1 : 0;
}
