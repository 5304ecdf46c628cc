// Geometry of one DP grid, shared by host planner, device kernels and CPU unit tests.
//
// The device works in plain matrix coordinates (row i = position in the vertical
// sequence, column j = position in the horizontal sequence).  The reference's traceback
// hand-off between banded-chain grids, however, is defined in terms of SeqAn's matrix
// navigators ("storage row" cv of a banded column, the navigator's lane leap), so the
// few places where those leak into results (tracking options, traceback coordinator)
// need the mapping (i, j) <-> (cv, tLeap).  This header provides it:
//   * storageOffset(): closed form cv - i per column
//     (seqan/align/dp_matrix_navigator_trace_matrix.h:72-196),
//   * BandWalker: the column sequence of _computeBandedAlignment
//     (seqan/align/dp_algorithm_impl.h:515-860) with column descriptors and lane leaps.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define UB_HD __host__ __device__ __forceinline__
#else
#define UB_HD inline
#endif

namespace ub200 {

enum ColProp : int32_t { CP_INITIAL = 0, CP_INNER = 1, CP_FINAL = 2 };
enum ColLoc : int32_t { CL_FULL = 0, CL_TOP = 1, CL_MIDDLE = 2, CL_BOTTOM = 3 };
enum CellType : int32_t { CT_FIRST = 0, CT_INNER = 1, CT_LAST = 2 };

UB_HD int32_t imin(int32_t a, int32_t b) { return a < b ? a : b; }
UB_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }

struct GridGeom {
    int32_t nH, nV, banded, lo, up;
    int32_t dimV;    // vertical length of SeqAn's trace matrix (dp_algorithm_impl.h:1547-1560)
    int32_t U, L;    // clipped diagonals
    int32_t narrow;  // banded and the band is narrower than the matrix height
};

UB_HD GridGeom makeGeom(int32_t nH, int32_t nV, int32_t banded, int32_t lo, int32_t up) {
    GridGeom g;
    g.nH = nH; g.nV = nV; g.banded = banded; g.lo = lo; g.up = up;
    if (!banded) {
        g.dimV = nV + 1; g.U = nH; g.L = -nV; g.narrow = 0;
    } else {
        g.U = imin(nH, up);
        g.L = imax(lo, -nV);
        int32_t bandSize = g.U - g.L + 1;
        g.dimV = imin(nV + 1, bandSize);
        g.narrow = (bandSize <= nV + 1) ? 1 : 0;
    }
    return g;
}

// cv(i, j) - i for the trace-matrix navigator.  Unbanded: 0.
UB_HD int32_t storageOffset(const GridGeom& g, int32_t j) {
    if (!g.banded) return 0;
    if (g.narrow) return g.U - j;
    return imax(0, g.nV + g.L - j) - imax(0, j - g.up);
}

// First / last row of column j that lies inside the band (matrix coordinates).
UB_HD int32_t colTop(const GridGeom& g, int32_t j) { return g.banded ? imax(0, j - g.up) : 0; }
UB_HD int32_t colBottom(const GridGeom& g, int32_t j) { return g.banded ? imin(g.nV, j - g.lo) : g.nV; }

struct ColInfo {
    int32_t j;         // matrix column
    int32_t cp, cl;    // column property / location
    int32_t rowTop;    // matrix row of the FirstCell
    int32_t nCells;    // 1 => only a FirstCell is computed
    int32_t tLeap;     // navigator lane leap seen by First/Inner cells
    int32_t tLeapLast; // ... and by the LastCell
    int32_t cvFirst;   // storage row of the FirstCell
};

// Literal walk over the columns of a banded DP.  next() yields columns in computation
// order.  Only upper >= 0 > lower bands are supported (every caller on the hot path).
struct BandWalker {
    GridGeom g;
    int32_t vBegin, vEnd, h, hEndTop, hEndMid, hEndBottom;
    int32_t tpos_col, tpos_cv, tLeap;
    int32_t phase;  // 0 initial, 1 top, 2 mid, 3 bottom, 4 tail, 5 done
    int32_t midIsFull;

    UB_HD void init(const GridGeom& gg) {
        g = gg;
        const int32_t nH = g.nH, nV = g.nV, lo = g.lo, up = g.up;
        vBegin = 0 - imin(0, 1 + up);
        vEnd = 0 - imin(0, imax(-nV, lo));
        h = imax(0, imin(nH - 1, lo));
        hEndTop = imin(nH - 1, imax(0, up));
        hEndMid = imin(nH - 1, imax(0, nV + lo));
        midIsFull = up > nV + lo;
        if (midIsFull) { int32_t t = hEndTop; hEndTop = hEndMid; hEndMid = t; }
        hEndBottom = imax(0, imin(nH, up + nV) - 1);
        int32_t lastPos = imax(-(g.dimV - 1), lo);
        tLeap = g.dimV + lastPos;
        tpos_col = 0;
        tpos_cv = tLeap - 1;
        phase = 0;
    }

    UB_HD void advanceTo(int32_t delta) {  // tpos += delta, keeping (col, cv) normalised
        int32_t cv = tpos_cv + delta;
        while (cv >= g.dimV) { cv -= g.dimV; ++tpos_col; }
        tpos_cv = cv;
    }

    UB_HD ColInfo emit(int32_t cp, int32_t cl, int32_t rowTop, int32_t nCells) {
        ColInfo c;
        c.cp = cp; c.cl = cl; c.rowTop = rowTop; c.nCells = nCells;
        // FirstCell navigation (dp_matrix_navigator_trace_matrix.h:118-154)
        if (cp != CP_INITIAL) {
            if (cl == CL_TOP) --tLeap;
            advanceTo(tLeap);
        }
        c.j = tpos_col;
        c.cvFirst = tpos_cv;
        c.tLeap = tLeap;
        if (nCells > 1) {
            advanceTo(nCells - 1);
            if (cp != CP_INITIAL && cl == CL_BOTTOM) ++tLeap;
        }
        c.tLeapLast = tLeap;
        return c;
    }

    // returns false when the walk is finished
    UB_HD bool next(ColInfo& out) {
        const int32_t nH = g.nH, nV = g.nV, lo = g.lo, up = g.up;
        if (phase == 0) {
            phase = 1;
            if (h == nH - 1) { phase = 5; out = emit(CP_INITIAL, CL_TOP, 0, 1); return true; }
            if (hEndBottom == 0) { phase = 5; out = emit(CP_INITIAL, CL_BOTTOM, vBegin, 1); return true; }
            if (lo <= -nV) out = emit(CP_INITIAL, CL_FULL, 0, vEnd - vBegin + 1);
            else out = emit(CP_INITIAL, CL_TOP, 0, vEnd - vBegin + 1);
            return true;
        }
        if (phase == 1) {
            if (h != hEndTop) {
                ++h; ++vEnd;
                out = emit(CP_INNER, CL_TOP, 0, vEnd - vBegin + 1);
                return true;
            }
            phase = 2;
        }
        if (phase == 2) {
            if (h != hEndMid) {
                ++h;
                if (midIsFull) out = emit(CP_INNER, CL_FULL, 0, vEnd - vBegin + 1);
                else { ++vBegin; ++vEnd; out = emit(CP_INNER, CL_MIDDLE, vBegin, vEnd - vBegin + 1); }
                return true;
            }
            phase = 3;
        }
        if (phase == 3) {
            if (h != hEndBottom) {
                ++h; ++vBegin;
                out = emit(CP_INNER, CL_BOTTOM, vBegin, vEnd - vBegin + 1);
                return true;
            }
            phase = 4;
        }
        if (phase == 4) {
            phase = 5;
            if (h < nH - 1) { out = emit(CP_INNER, CL_BOTTOM, vBegin + 1, 1); return true; }
            if (h == nH - 1) {
                if (up == nH - nV) { out = emit(CP_FINAL, CL_BOTTOM, vBegin + 1, 1); return true; }
                if (up >= nH) {
                    if (lo + nV > nH) { ++vEnd; out = emit(CP_FINAL, CL_TOP, 0, vEnd - vBegin + 1); }
                    else if (lo + nV + 1 > nH) { ++vEnd; out = emit(CP_FINAL, CL_TOP, 0, vEnd - vBegin + 1); }
                    else out = emit(CP_FINAL, CL_FULL, 0, vEnd - vBegin + 1);
                } else {
                    ++vBegin;
                    if (lo + nV <= nH) {
                        if (lo + nV == nH) { ++vEnd; out = emit(CP_FINAL, CL_MIDDLE, vBegin, vEnd - vBegin + 1); }
                        else out = emit(CP_FINAL, CL_BOTTOM, vBegin, vEnd - vBegin + 1);
                    } else { ++vEnd; out = emit(CP_FINAL, CL_MIDDLE, vBegin, vEnd - vBegin + 1); }
                }
                return true;
            }
        }
        return false;
    }
};

// The walker yields the columns 0 .. bandLastColumn(g), one after the other (checked against the walker for every
// geometry up to 45 x 45 with bands up to +-50, tests/cpp/test_geom.cpp): the number of column descriptors a grid
// needs right of the next grid's origin is known without walking.
UB_HD int32_t bandLastColumn(const GridGeom& g) {
    const int32_t hEndBottom = imax(0, imin(g.nH, g.up + g.nV) - 1);
    return (g.nH == 1 || hEndBottom == 0) ? 0 : hEndBottom + 1;
}
UB_HD int32_t bandColumnsFrom(const GridGeom& g, int32_t hNext) { return imax(0, bandLastColumn(g) - imax(0, hNext) + 1); }

// seeds/banded_chain_alignment_impl.h:282-377 (_determineTrackingOptions), literal.
// chainFinal: BandedChainFinalDPMatrix; feLastRow/feLastCol: free end gaps.
struct TrackOpts {
    bool lastCol, lastRow, storeCol, storeRow;
};
UB_HD TrackOpts chainTrackingOptions(int32_t ch, int32_t cv, int32_t tLeap, int32_t cp, int32_t cl, int32_t ct,
                                     int32_t hNext, int32_t vNext, bool chainFinal, bool feLastRow,
                                     bool feLastCol) {
    TrackOpts o;
    o.lastCol = o.lastRow = o.storeCol = o.storeRow = false;
    if (ch >= hNext) {
        if (cl == CL_BOTTOM) {
            if (cv + tLeap == vNext) o.storeRow = true;
        } else {
            if (cv == vNext) o.storeRow = true;
        }
        if (ch == hNext && cv >= vNext) o.storeCol = true;
        if (ct == CT_LAST) {
            if (chainFinal) { if (feLastRow) o.lastRow = true; }
            else o.lastRow = true;
        }
        if (cp == CP_FINAL) {
            if (ct == CT_LAST) o.lastCol = o.lastRow = true;
            else if (cl != CL_FULL || cv >= vNext) {
                if (chainFinal) { if (feLastCol) o.lastCol = true; }
                else o.lastCol = true;
            }
        }
    }
    return o;
}


// ---------------------------------------------------------------------------------------
// Strip / checkpoint geometry of the device engine (shared with the host memory planner).
// ---------------------------------------------------------------------------------------
constexpr int32_t SH = 256;              // strip height of task grids (32 lanes x 8 rows)
constexpr int32_t CKW = 64;              // column-checkpoint spacing == recompute tile width
constexpr int32_t CKR = 64;              // row-checkpoint spacing == recompute tile height (32 lanes x 2 rows)
constexpr int32_t SEG = 1024;            // columns per work item
constexpr int32_t WINBYTES = 46 * 1024;  // shared-memory window per control warp (trace bytes, or the pass-1 box)

// first / last column of strip s (rows s*SHR+1 .. (s+1)*SHR) that holds band cells
UB_HD int32_t stripJlo(const GridGeom& g, int32_t s, int32_t SHR) { return g.banded ? imax(1, s * SHR + 1 + g.lo) : 1; }
UB_HD int32_t stripJhi(const GridGeom& g, int32_t s, int32_t SHR) {
    if (!g.banded) return g.nH;
    int32_t r1 = imin(g.nV, (s + 1) * SHR);
    return imin(g.nH, r1 + g.up);
}
// number of strips that hold band cells
UB_HD int32_t stripCount(const GridGeom& g, int32_t SHR) {
    int32_t rowsReach = g.banded ? imin(g.nV, g.nH - g.lo) : g.nV;
    return (rowsReach + SHR - 1) / SHR;
}
// first checkpoint column (in units of CKW) inside strip s, and how many the strip holds
UB_HD int32_t ckFirst(const GridGeom& g, int32_t s) { return (stripJlo(g, s, SH) + CKW - 1) / CKW; }
UB_HD int32_t ckCount(const GridGeom& g, int32_t s) {
    int32_t n = stripJhi(g, s, SH) / CKW - ckFirst(g, s) + 1;
    return n > 0 ? n : 0;
}

// A grid is "local" when one warp can fill it with the full trace kept in its shared-memory window.
struct LocalPlan {
    int32_t local, RR, RRS, pitch, jhi;  // rows per lane, bytes per (column, lane) slot, lanes per column, last column
};
UB_HD LocalPlan localPlan(const GridGeom& g) {
    LocalPlan p;
    p.RR = (g.nV <= 64) ? 2 : (g.nV <= 96) ? 3 : (g.nV <= 128) ? 4 : 8;
    p.RRS = (p.RR == 3) ? 4 : p.RR;
    int32_t lanes = (g.nV + p.RR - 1) / p.RR;
    p.pitch = (lanes + 1) & ~1;
    p.jhi = stripJhi(g, 0, 32 * p.RR);
    p.local = (g.nV <= 32 * p.RR && (int64_t)p.jhi * p.pitch * p.RRS <= (int64_t)WINBYTES) ? 1 : 0;
    return p;
}

// Reference cell count of a grid: dimH * dimV of the allocated score matrix.
UB_HD int64_t gridCells(int32_t nH, int32_t nV, int32_t banded, int32_t lo, int32_t up) {
    GridGeom g = makeGeom(nH, nV, banded, lo, up);
    int64_t dimH = (int64_t)nH + 1 - (banded ? imax(0, lo) : 0);
    return dimH * (int64_t)g.dimV;
}

}  // namespace ub200
