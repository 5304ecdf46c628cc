// Device side of the B200 DP engine (sm_100a).  See engine.hpp for the job model.
//
// Execution model
//   * persistent CTAs (NWARPS warps) pull jobs from an atomic queue; a job's grids run
//     back to back inside the CTA because grid k+1 is initialised from the traceback of
//     grid k (seeds/banded_chain_alignment_traceback.h:296-330).
//   * fill: the matrix is cut into horizontal strips of SH = 32*R rows.  Inside a strip
//     lane t owns R consecutive rows and walks the columns one step behind lane t-1
//     (anti-diagonal wavefront, cell hand-off with __shfl_up); strips are pipelined over
//     the CTA's warps, `lag` 32-column chunks apart, with one __syncthreads per chunk
//     phase; the row between two strips travels through an L2-resident buffer.
//   * trace: one byte per cell (the reference's TraceBitMap_ value), written as one 16-byte
//     store per lane and step into a skewed, fully coalesced layout:
//        addr(i,j) = stripBase[s] + ((j - jlo(s) + t)*32 + t)*R + r,
//        s=(i-1)/SH, t=((i-1)%SH)/R, r=(i-1)%R.
//   * tracking + traceback: warp 0 finds the tied maxima over the tracked cells and walks
//     the trace exactly like SeqAn's TracebackCoordinator_, in SeqAn's storage coordinates
//     (dpgeom.hpp maps them to matrix coordinates).
#pragma once
#include <cuda_runtime.h>

#include "dpgeom.hpp"
#include "engine.hpp"

namespace ub200 {

constexpr int R = 8;
constexpr int SH = 32 * R;
constexpr int NWARPS = 16;
constexpr int NTHREADS = NWARPS * 32;
constexpr unsigned FULLMASK = 0xffffffffu;

struct DCell {
    int s, h, v;
};

struct JobDev {
    long long hOff, vOff;
    int lenH, lenV;
    int match, mismatch, gapOpen, gapExtend;
    int fe;  // bit0 firstRow, bit1 firstCol, bit2 lastRow, bit3 lastCol
    int complete;
    int gridBegin, gridCount;
    long long outOff;
    int outCap;
    int pad;
};

struct JobOut {
    int status, score, outLen, pad;
    long long prof[6];  // cycles: setup, init, fill, track, traceback, total (thread 0)
};

struct ScratchLayout {  // byte offsets inside one CTA's scratch block
    long long trace, stripBase, bnd, initRow, initCol, hInitNext, vInitNext, box, lastRow, lastCol, cand, planted,
        colTab, total;
    int bndStride, maxCand, maxPlanted, maxColTab;
    long long maxBox;
    int maxCapH, maxCapV, maxNH, maxNV;
    long long maxTrace;
    int maxStrips, pad;
};

struct KParams {
    const JobDev* jobs;
    const GridDesc* grids;
    const uint8_t* seq;
    int* out;
    JobOut* jobOut;
    const int* order;  // job processing order (largest first)
    int nJobs;
    int* queue;
    uint8_t* scratch;
    long long scratchStride;
    ScratchLayout lay;
};

struct PlantedCell {
    int i1, i2;
    DCell c;
};

// Everything a grid needs, built once per grid by thread 0 (shared memory).
struct GridCtx {
    GridGeom g;
    int kind, h0, v0, hNext, vNext, capNextH, capNextV;
    int match, mismatch, go, ge, fe, complete, affine;
    const uint8_t* seqH;
    const uint8_t* seqV;
    // scratch pointers
    uint8_t* trace;
    long long* stripBase;
    int2* bnd;
    DCell *initRow, *initCol, *hInitNext, *vInitNext, *box, *lastRow, *lastCol;
    int* cand;
    PlantedCell* planted;
    ColInfo* colTab;
    // strips / schedule
    int NS, P, lag, totalPhases, bndStride;
    int colZeroMax;
    // capture
    int capEdges;  // 1: last row + last column; 0: box
    int boxRow0, boxH, boxW;
    // limits
    int maxCand, maxPlanted, maxColTab;
    long long maxBox, maxTrace;
};

__device__ __forceinline__ int stripJlo(const GridGeom& g, int s) {
    return g.banded ? imax(1, s * SH + 1 + g.lo) : 1;
}
__device__ __forceinline__ int stripJhi(const GridGeom& g, int s) {
    if (!g.banded) return g.nH;
    int r1 = imin(g.nV, (s + 1) * SH);
    return imin(g.nH, r1 + g.up);
}
__device__ __forceinline__ int stripChunks(const GridGeom& g, int s) {
    int ncols = stripJhi(g, s) - stripJlo(g, s) + 1;
    return ncols > 0 ? (ncols + 62) / 32 : 0;
}

__device__ __forceinline__ size_t traceAddr(const GridCtx& G, int i, int j) {
    int s = (i - 1) / SH;
    int rem = (i - 1) - s * SH;
    int t = rem / R, r = rem - t * R;
    int k = j - stripJlo(G.g, s) + t;
    return (size_t)G.stripBase[s] + ((size_t)k * 32 + t) * R + r;
}

// ---------------------------------------------------------------------------------------
// cell recurrences (seqan/align/dp_formula_affine.h:459-636, dp_formula_linear.h:150-291)
// mode: 0 = RecursionDirectionAll, 1 = UpperDiagonal, 2 = LowerDiagonal, 3 = outside the band
// ---------------------------------------------------------------------------------------
template <bool AFF, bool CT, bool BANDED>
__device__ __forceinline__ uint32_t cellUpdate(int& s, int& h, int& v, int sl, int hl, int su, int vu, int sd,
                                               int sub, int go, int ge, int mode) {
    // Branch-free formulation.  The only loop-carried chain inside a lane is su -> e -> v -> s (3 ops);
    // everything else depends on the previous column only.
    uint32_t tv;
    if (AFF && !BANDED) {
        // Fast path.  Loop-carried chain per cell: su -> VIADDMNMX -> VIMNMX3 -> s (2 ops).
        const int a = hl + ge, b = sl + go;
        const int hh = max(a, b);
        const int c = vu + ge, e = su + go;
        const int vv = __viaddmax_s32(su, go, c);
        const int d = sd + sub;
        const int ss = __vimax3_s32(vv, hh, d);
        uint32_t tvH, tvV, tvM;
        if (CT) {
            tvH = ((a >= b) ? (uint32_t)T_H : 0u) | ((a <= b) ? (uint32_t)T_HO : 0u);
            tvV = ((c >= e) ? (uint32_t)T_V : 0u) | ((c <= e) ? (uint32_t)T_VO : 0u);
            tvM = ((vv >= hh) ? (uint32_t)T_MV : 0u) | ((vv <= hh) ? (uint32_t)T_MH : 0u);
            // m = max(vv,hh):  m <= d  <=>  d == s ;  m >= d  <=>  vv == s || hh == s
            tv = tvH | tvV | ((d == ss) ? (uint32_t)T_D : 0u) | ((vv == ss || hh == ss) ? tvM : 0u);
        } else {
            tvH = (a < b) ? (uint32_t)T_HO : (uint32_t)T_H;
            tvV = (c < e) ? (uint32_t)T_VO : (uint32_t)T_V;
            tvM = (vv < hh) ? (uint32_t)T_MH : (uint32_t)T_MV;
            tv = tvH | tvV | ((d == ss) ? (uint32_t)T_D : tvM);
        }
        s = ss; h = hh; v = vv;
    } else if (AFF) {
        const int a = hl + ge, b = sl + go;
        const int c = vu + ge, e = su + go;
        const int d = sd + sub;
        int hh = max(a, b);
        int vv = max(c, e);
        uint32_t tvH, tvV;
        if (CT) {
            tvH = ((a >= b) ? (uint32_t)T_H : 0u) | ((a <= b) ? (uint32_t)T_HO : 0u);
            tvV = ((c >= e) ? (uint32_t)T_V : 0u) | ((c <= e) ? (uint32_t)T_VO : 0u);
        } else {
            tvH = (a < b) ? (uint32_t)T_HO : (uint32_t)T_H;
            tvV = (c < e) ? (uint32_t)T_VO : (uint32_t)T_V;
        }
        if (BANDED) {
            const bool top = (mode == 1), bot = (mode == 2);
            vv = top ? NEG_INF : vv; tvV = top ? 0u : tvV;
            hh = bot ? NEG_INF : hh; tvH = bot ? 0u : tvH;
        }
        const int m = max(vv, hh);
        uint32_t tvM;
        if (CT) tvM = ((vv >= hh) ? (uint32_t)T_MV : 0u) | ((vv <= hh) ? (uint32_t)T_MH : 0u);
        else tvM = (vv < hh) ? (uint32_t)T_MH : (uint32_t)T_MV;
        if (BANDED) {
            tvM = (mode == 1) ? (uint32_t)T_MH : tvM;
            tvM = (mode == 2) ? (uint32_t)T_MV : tvM;
        }
        const uint32_t gap = tvH | tvV;
        s = max(m, d);
        if (CT) tv = gap | ((m <= d) ? (uint32_t)T_D : 0u) | ((m >= d) ? tvM : 0u);
        else tv = gap | ((m <= d) ? (uint32_t)T_D : tvM);
        h = hh; v = vv;
    } else {
        const int x0 = sd + sub;
        int tV = su + ge, tH = sl + ge;
        if (BANDED) {
            tV = (mode == 1) ? INT32_MIN : tV;   // UpperDiagonal: no vertical candidate
            tH = (mode == 2) ? INT32_MIN : tH;   // LowerDiagonal: no horizontal candidate
        }
        const int x1 = max(x0, tV);
        const int x2 = max(x1, tH);
        uint32_t t1;
        if (CT) t1 = ((x0 >= tV) ? (uint32_t)T_D : 0u) | ((x0 <= tV) ? (uint32_t)(T_V | T_MV) : 0u);
        else t1 = (x0 < tV) ? (uint32_t)(T_V | T_MV) : (uint32_t)T_D;
        if (CT) tv = ((x1 >= tH) ? t1 : 0u) | ((x1 <= tH) ? (uint32_t)(T_H | T_MH) : 0u);
        else tv = (x1 < tH) ? (uint32_t)(T_H | T_MH) : t1;
        s = x2; h = NEG_INF; v = NEG_INF;
    }
    if (BANDED) {
        const bool outside = (mode == 3);
        s = outside ? NEG_INF : s; h = outside ? NEG_INF : h; v = outside ? NEG_INF : v; tv = outside ? 0u : tv;
    }
    return tv;
}

// S and V-matrix value of the cell just above strip s in column j (j >= 1)
template <bool BANDED>
__device__ __forceinline__ void upBoundary(const GridCtx& G, int s, int j, int& bS, int& bV) {
    int rowAbove = s * SH;
    if (BANDED) {
        int d = j - rowAbove;
        if (d < G.g.lo || d > G.g.up) { bS = NEG_INF; bV = NEG_INF; return; }
    }
    if (s == 0) {
        DCell c = G.initRow[j];
        bS = c.s; bV = c.v;
    } else {
        int2 b = __ldcg(&G.bnd[(size_t)((s - 1) & 1) * G.bndStride + j]);
        bS = b.x; bV = b.y;
    }
}

// Loop-invariant scalars of one strip chunk, kept in registers (GridCtx lives in shared memory).
struct StepConsts {
    int match, mismatch, go, ge, nV, nH, lo, up;
    uint8_t* tbase;   // trace base of this strip
    int2* bndOut;     // boundary row written by lane 31 (nullptr: no strip below)
};

template <bool AFF, bool CT, bool BANDED, bool CAP>
__device__ __forceinline__ void stripSteps(const GridCtx& G, const StepConsts& K, int c, int lane, int jlo, int jhi,
                                           int i0, int (&Sl)[R], int (&Hl)[R], const uint32_t (&vcw)[R / 4],
                                           int& prevUpS, int& pubS, int& pubV, int& curHc, int bS, int bV, int hcN) {
    const int match = K.match, mismatch = K.mismatch, go = K.go, ge = K.ge;
    const int lo = K.lo, up = K.up;
    // capture parameters (slow variant only)
    int capEdges = 0, hNext = 0, boxRow0 = 0, boxH = 0;
    DCell *box = nullptr, *lastRow = nullptr, *lastCol = nullptr;
    if (CAP) {
        capEdges = G.capEdges; hNext = G.hNext; boxRow0 = G.boxRow0; boxH = G.boxH;
        box = G.box; lastRow = G.lastRow; lastCol = G.lastCol;
    }
#pragma unroll 1
    for (int kk = 0; kk < 32; ++kk) {
        const int k = 32 * c + kk;
        int inS = __shfl_up_sync(FULLMASK, pubS, 1);
        int inV = __shfl_up_sync(FULLMASK, pubV, 1);
        int inHc = __shfl_up_sync(FULLMASK, curHc, 1);
        const int l0S = __shfl_sync(FULLMASK, bS, kk);
        const int l0V = __shfl_sync(FULLMASK, bV, kk);
        const int l0Hc = __shfl_sync(FULLMASK, hcN, kk);
        if (lane == 0) { inS = l0S; inV = l0V; inHc = l0Hc; }
        curHc = inHc;
        const int j = jlo + k - lane;
        const bool act = (k >= lane) && (j <= jhi);
        if (act) {
            int Sd = prevUpS, Su = inS, Vu = inV;
            uint32_t tw[R / 4];
#pragma unroll
            for (int w4 = 0; w4 < R / 4; ++w4) tw[w4] = 0u;
            // per-byte equality of this lane's vertical codes with the column's horizontal code
            const uint32_t hc4 = (uint32_t)curHc * 0x01010101u;
            uint32_t eq[R / 4];
#pragma unroll
            for (int w4 = 0; w4 < R / 4; ++w4) eq[w4] = __vcmpeq4(vcw[w4], hc4);
            int vArr[CAP ? R : 1];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int mode = 0;
                if (BANDED) {
                    const int d = j - (i0 + r);
                    mode = (d < lo || d > up) ? 3 : (d == up ? 1 : (d == lo ? 2 : 0));
                }
                const int sub = (eq[r >> 2] & (1u << (8 * (r & 3)))) ? match : mismatch;
                int ns, nh, nv;
                const uint32_t tv = cellUpdate<AFF, CT, BANDED>(ns, nh, nv, Sl[r], Hl[r], Su, Vu, Sd, sub, go, ge, mode);
                Sd = Sl[r];
                Sl[r] = ns; Hl[r] = nh; Su = ns; Vu = nv;
                if (CAP) vArr[r] = nv;
                tw[r >> 2] |= tv << (8 * (r & 3));
            }
            prevUpS = inS;
            pubS = Su; pubV = Vu;
            static_assert(R == 8, "trace store assumes 8 rows per lane");
            *reinterpret_cast<uint2*>(K.tbase + ((size_t)k * 32 + lane) * R) = make_uint2(tw[0], tw[1]);
            if (K.bndOut != nullptr && lane == 31) __stcg(&K.bndOut[j], make_int2(Su, Vu));
            if (CAP) {
                if (capEdges) {
                    const int rl = K.nV - i0;
                    if (rl >= 0 && rl < R) {
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (r == rl) lastRow[j] = DCell{Sl[r], Hl[r], vArr[r]};
                    }
                    if (j == K.nH) {
#pragma unroll
                        for (int r = 0; r < R; ++r)
                            if (i0 + r <= K.nV) lastCol[i0 + r] = DCell{Sl[r], Hl[r], vArr[r]};
                    }
                } else if (j >= hNext) {
                    DCell* col = box + (size_t)(j - hNext) * boxH - boxRow0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int i = i0 + r;
                        bool inb = true;
                        if (BANDED) { const int d = j - i; inb = (d >= lo && d <= up); }
                        if (i >= boxRow0 && i <= K.nV && inb) col[i] = DCell{Sl[r], Hl[r], vArr[r]};
                    }
                }
            }
        }
    }
}

template <bool AFF, bool CT, bool BANDED>
__device__ __noinline__ void fillGrid(const GridCtx& G) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int Sl[R], Hl[R];
    uint32_t vcw[R / 4];
    int prevUpS = NEG_INF, pubS = NEG_INF, pubV = NEG_INF, curHc = 0;
    int jlo = 1, jhi = 0, i0 = 1, nch = 0;
    const GridGeom& g = G.g;
#pragma unroll 1
    for (int p = 0; p < G.totalPhases; ++p) {
        int t = p - warp * G.lag;
        if (t >= 0) {
            int q = t / G.P;
            int c = t - q * G.P;
            int s = warp + NWARPS * q;
            if (s < G.NS) {
                if (c == 0) {  // strip start: load this lane's rows
                    jlo = stripJlo(g, s);
                    jhi = stripJhi(g, s);
                    nch = stripChunks(g, s);
                    i0 = s * SH + lane * R + 1;
#pragma unroll
                    for (int w4 = 0; w4 < R / 4; ++w4) vcw[w4] = 0;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int i = i0 + r;
                        uint32_t code = (i <= g.nV) ? (uint32_t)G.seqV[i - 1] : 255u;
                        vcw[r >> 2] |= code << (8 * (r & 3));
                        if (jlo == 1 && i <= G.colZeroMax) { DCell ic = G.initCol[i]; Sl[r] = ic.s; Hl[r] = ic.h; }
                        else { Sl[r] = NEG_INF; Hl[r] = NEG_INF; }
                    }
                    if (jlo == 1) prevUpS = (i0 - 1 <= G.colZeroMax) ? G.initCol[i0 - 1].s : NEG_INF;
                    else {
                        prevUpS = NEG_INF;
                        if (lane == 0) { int bS, bV; upBoundary<BANDED>(G, s, jlo - 1, bS, bV); prevUpS = bS; }
                    }
                    pubS = NEG_INF; pubV = NEG_INF; curHc = 0;
                }
                if (c < nch) {
                    int jj = jlo + 32 * c + lane;
                    int bS = NEG_INF, bV = NEG_INF, hcN = 0;
                    if (jj <= jhi) { hcN = G.seqH[jj - 1]; upBoundary<BANDED>(G, s, jj, bS, bV); }
                    StepConsts K;
                    K.match = G.match; K.mismatch = G.mismatch; K.go = G.go; K.ge = G.ge;
                    K.nV = g.nV; K.nH = g.nH; K.lo = g.lo; K.up = g.up;
                    K.tbase = G.trace + (size_t)G.stripBase[s];
                    K.bndOut = (s + 1 < G.NS) ? (G.bnd + (size_t)(s & 1) * G.bndStride) : nullptr;
                    bool cap;
                    int jmaxChunk = jlo + 32 * c + 31;  // largest column any lane touches in this chunk
                    if (G.capEdges) cap = ((s + 1) * SH >= g.nV) || (jmaxChunk >= g.nH);
                    else cap = (jmaxChunk >= G.hNext) && ((s + 1) * SH >= G.boxRow0);
                    if (cap)
                        stripSteps<AFF, CT, BANDED, true>(G, K, c, lane, jlo, jhi, i0, Sl, Hl, vcw, prevUpS, pubS, pubV,
                                                          curHc, bS, bV, hcN);
                    else
                        stripSteps<AFF, CT, BANDED, false>(G, K, c, lane, jlo, jhi, i0, Sl, Hl, vcw, prevUpS, pubS, pubV,
                                                           curHc, bS, bV, hcN);
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// traceback (seqan/align/dp_traceback_impl.h, seeds/banded_chain_alignment_traceback.h)
// in SeqAn storage coordinates (col, cv).  Executed by ALL lanes of warp 0 with identical
// (warp-uniform) control flow: the walk itself is serial, but the trace bytes come from a
// 64x64 window staged in shared memory that the 32 lanes load cooperatively (one window
// serves >= 64 steps), and only lane 0 writes results.
// ---------------------------------------------------------------------------------------
constexpr int WIN = 64;

struct Coord {
    int currCol, currRow, endCol, endRow, bp1, bp2;
    bool inBandFlag;
    __device__ __forceinline__ bool reachedEnd() const { return currCol <= endCol || currRow <= endRow; }
    __device__ __forceinline__ bool isInBand() const {
        if (!inBandFlag) return false;
        return currCol > bp1 || currCol <= bp2;
    }
};

struct OutStream {
    int* buf;
    int cap, len;
    bool overflow;
    int h0, v0;
    int lane;
    __device__ __forceinline__ void put(int x) {
        if (len < cap) { if (lane == 0) buf[len] = x; }
        else overflow = true;
        ++len;
    }
    __device__ __forceinline__ void patch(int pos, int x) {
        if (pos < cap && lane == 0) buf[pos] = x;
    }
};

struct TraceWalker {
    const GridCtx& G;
    OutStream& out;
    uint8_t* win;     // shared memory, WIN*WIN bytes, win[(row - wRow0) * WIN + (col - wCol0)]
    int wRow0, wCol0; // matrix coordinates of the window origin; wRow0 < 0: no window loaded
    int wRowEnd, wColEnd;  // last row / column staged
    int pc, pv;       // navigator position: column, storage row
    int nSegs;        // segments emitted for the current trace
    bool emitOn;
    bool bad;         // undefined trace value (reference: endless loop / assert)

    __device__ TraceWalker(const GridCtx& g, OutStream& o, uint8_t* w)
        : G(g), out(o), win(w), wRow0(-1), wCol0(0), wRowEnd(-1), wColEnd(-1), pc(0), pv(0), nSegs(0), emitOn(true), bad(false) {}

    // Stage the window whose bottom-right corner is the 8-row group of (i, j).
    __device__ void loadWindow(int i, int j) {
        const int lane = threadIdx.x & 31;
        const int gBot = (i - 1) / R;                       // row group of row i (rows 8g+1 .. 8g+8)
        const int gTop = imax(0, gBot - (WIN / R - 1));
        wRow0 = gTop * R + 1;
        wCol0 = imax(1, j - (WIN - 1));
        wRowEnd = (gBot + 1) * R;
        wColEnd = j;
        __syncwarp();
        const int nGroups = gBot - gTop + 1;
        const int nCols = j - wCol0 + 1;
        for (int idx = lane; idx < nGroups * nCols; idx += 32) {
            const int gq = gTop + idx / nCols;
            const int col = wCol0 + idx % nCols;
            const int row = gq * R + 1;
            const int s = (row - 1) / SH;
            const int t = ((row - 1) - s * SH) / R;
            const int jlo = stripJlo(G.g, s), jhi = stripJhi(G.g, s);
            uint2 v = make_uint2(0u, 0u);
            if (col >= jlo && col <= jhi && row <= G.g.nV && s < G.NS) {
                const int k = col - jlo + t;
                v = *reinterpret_cast<const uint2*>(G.trace + (size_t)G.stripBase[s] + ((size_t)k * 32 + t) * R);
            }
            uint8_t* dst = win + (size_t)(row - wRow0) * WIN + (col - wCol0);
#pragma unroll
            for (int r = 0; r < 4; ++r) dst[r * WIN] = (uint8_t)(v.x >> (8 * r));
#pragma unroll
            for (int r = 0; r < 4; ++r) dst[(4 + r) * WIN] = (uint8_t)(v.y >> (8 * r));
        }
        __syncwarp();
    }

    __device__ __forceinline__ uint32_t tvHere() {
        const int i = pv - storageOffset(G.g, pc);
        const int j = pc;
        if (i <= 0 || j <= 0 || i > G.g.nV || j > G.g.nH) return 0;
        if (wRow0 < 0 || i < wRow0 || j < wCol0 || i > wRowEnd || j > wColEnd) loadWindow(i, j);
        return win[(size_t)(i - wRow0) * WIN + (j - wCol0)];
    }
    __device__ Coord makeCoord(int endCol, int endRow) const {  // dp_traceback_impl.h:121-141
        Coord c;
        c.currCol = pc; c.currRow = pv; c.endCol = endCol; c.endRow = endRow; c.bp1 = 0; c.bp2 = 0;
        c.inBandFlag = false;
        if (G.g.banded) {
            if (c.currCol > G.g.up) c.currRow += c.currCol - G.g.up;
            if (c.endCol > G.g.up) c.endRow += c.endCol - G.g.up;
            c.bp1 = imin(G.g.nH, imax(0, G.g.up));
            c.bp2 = imin(G.g.nH, imax(0, G.g.nV + G.g.lo));
            int mb = imin(c.bp1, c.bp2);
            if (c.currCol < mb) c.currRow -= mb - c.currCol;
            c.inBandFlag = true;
        }
        return c;
    }
    __device__ __forceinline__ void record(int h, int v, int len, uint32_t tv) {  // dp_trace_segment.h:319-337
        if (len == 0) return;
        int dir;
        if (tv & T_D) dir = T_D;
        else if (tv & T_V) dir = T_V;
        else if (tv & T_H) dir = T_H;
        else return;
        if (!emitOn) return;
        out.put(h + out.h0); out.put(v + out.v0); out.put(len); out.put(dir);
        ++nSegs;
    }
    __device__ __forceinline__ void moveH(const Coord& c) { if (c.isInBand()) { --pc; ++pv; } else --pc; }
    __device__ __forceinline__ void moveD(const Coord& c) { if (c.isInBand()) { --pc; } else { --pc; --pv; } }
    __device__ __forceinline__ void moveV() { --pv; }

    __device__ void doTraceback(uint32_t& tv, uint32_t& last, int& frag, Coord& c) {  // dp_traceback_impl.h:335-431
        const bool aff = G.affine;
        if (tv & T_D) {
            if (!(last & T_D)) { record(c.currCol, c.currRow, frag, last); last = T_D; frag = 0; }
            moveD(c); tv = tvHere(); --c.currCol; --c.currRow; ++frag;
        } else if ((tv & T_MV) && (tv & T_V)) {
            if (!(last & T_V)) { record(c.currCol, c.currRow, frag, last); last = T_V; frag = 0; }
            if (aff) {
                while ((!(tv & T_VO) || (tv & T_V)) && c.currRow != 1) { moveV(); tv = tvHere(); --c.currRow; ++frag; }
                moveV(); tv = tvHere(); --c.currRow; ++frag;
            } else { moveV(); tv = tvHere(); --c.currRow; ++frag; }
        } else if ((tv & T_MV) && (tv & T_VO)) {
            if (!(last & T_V)) { record(c.currCol, c.currRow, frag, last); last = T_V; frag = 0; }
            moveV(); tv = tvHere(); --c.currRow; ++frag;
        } else if ((tv & T_MH) && (tv & T_H)) {
            if (!(last & T_H)) { record(c.currCol, c.currRow, frag, last); last = T_H; frag = 0; }
            if (aff) {
                while ((!(tv & T_HO) || (tv & T_H)) && c.currCol != 1) { moveH(c); tv = tvHere(); --c.currCol; ++frag; }
                moveH(c); tv = tvHere(); --c.currCol; ++frag;
            } else { moveH(c); tv = tvHere(); --c.currCol; ++frag; }
        } else if ((tv & T_MH) && (tv & T_HO)) {
            if (!(last & T_H)) { record(c.currCol, c.currRow, frag, last); last = T_H; frag = 0; }
            moveH(c); tv = tvHere(); --c.currCol; ++frag;
        } else {
            if (tv != T_NONE) { bad = true; tv = T_NONE; }
        }
    }
    __device__ static uint32_t initialDirection(uint32_t& tv, bool prefer) {  // dp_traceback_impl.h:433-461
        if (prefer) {
            if (tv & T_MV) { tv &= (T_V | T_VO | T_MV); return T_V; }
            if (tv & T_MH) { tv &= (T_H | T_HO | T_MH); return T_H; }
            return T_D;
        }
        if (tv & T_D) return T_D;
        if (tv & (T_V | T_MV)) return T_V;
        if (tv & (T_H | T_MH)) return T_H;
        return T_NONE;
    }
    // generic _computeTraceback (dp_traceback_impl.h:463-526); tvOverride >= 0 replaces the
    // start cell's trace value (the SingleTrace _correctTraceValue patch, dp_algorithm_impl.h:1354-1370)
    __device__ void generic(bool prefer, bool head, bool tail, int tvOverride) {
        uint32_t tv = tvOverride >= 0 ? (uint32_t)tvOverride : tvHere();
        uint32_t last = initialDirection(tv, prefer);
        Coord c = makeCoord(0, 0);
        const int nH = G.g.nH, nV = G.g.nV;
        if (tail) {
            if (c.currRow != nV) record(nH, c.currRow, nV - c.currRow, T_V);
            if (c.currCol != nH) record(c.currCol, c.currRow, nH - c.currCol, T_H);
        }
        int frag = 0;
        while (!c.reachedEnd() && tv != T_NONE) doTraceback(tv, last, frag, c);
        record(c.currCol, c.currRow, frag, last);
        if (head) {
            if (c.currRow != 0) record(0, 0, c.currRow, T_V);
            if (c.currCol != 0) record(0, 0, c.currCol, T_H);
        }
    }
};

}  // namespace ub200
