// ORACLE — TEST INFRASTRUCTURE ONLY (see dp_oracle.hpp for scope and parity status).
//
// Literal scalar emulation of SeqAn 2.3.1's column-wise DP as patched and used by
// Unicycler v0.5.1.  Citations are relative to /root/reference/unicycler/include/seqan.
#include "dp_oracle.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace orc {

CellHook g_cellHook = nullptr;
GridHook g_gridHook = nullptr;

namespace {

enum ColProp { INITIAL, INNER, FINAL };           // align/dp_meta_info.h:60-65
enum ColLoc { FULL, TOP, MIDDLE, BOTTOM };        // align/dp_meta_info.h:72-79
enum CellT { FIRST, INNERC, LAST };               // align/dp_meta_info.h:80-85
enum Rec { R_ZERO, R_H, R_V, R_ALL, R_UD, R_LD }; // align/dp_formula.h:60-71

// align/dp_traceback_impl.h:57-140 (TracebackCoordinator_)
struct Coord {
    long currCol, currRow, endCol, endRow, bp1, bp2;
    bool inBandFlag;
    bool reachedEnd() const { return currCol <= endCol || currRow <= endRow; }  // :107-111
    bool isInBand() const {                                                     // :143-150
        if (!inBandFlag) return false;
        return currCol > bp1 || currCol <= bp2;
    }
};

struct DPRun {
    const DPProblem& p;
    bool affine;
    long dimH, dimV;
    std::vector<uint8_t> trace;
    std::vector<Cell> col;
    // trace-matrix navigator (align/dp_matrix_navigator_trace_matrix.h)
    long tpos = 0;
    long tLeap = 0;
    // sparse score-matrix navigator (align/dp_matrix_navigator_score_matrix_sparse.h)
    long sAct = 0, sPrev = 0, sLeap = 0;
    Cell pD, pH, pV;
    // default scout (align/dp_scout.h:86-99) and chain scout (seeds/banded_chain_alignment_scout.h:81-103)
    Cell maxCell;
    unsigned maxPos = (unsigned)NEG;
    std::vector<unsigned> maxPositions;

    explicit DPRun(const DPProblem& prob) : p(prob), affine(prob.sc.affine()) {}

    long coordH() const { return tpos / dimV; }  // align/dp_matrix.h coordinate(): column-major host
    long coordV() const { return tpos % dimV; }

    // ------------------------------------------------------------------ navigators
    void initNavigators() {
        // align/dp_algorithm_impl.h:1547-1560
        dimH = p.nH + 1 - std::max(0L, p.banded ? p.lower : 0L);
        if (!p.banded)
            dimV = p.nV + 1;
        else {
            long bandSize = std::min(p.nH, p.upper) - std::max(p.lower, -p.nV) + 1;
            dimV = std::min(p.nV + 1, bandSize);
        }
        col.assign((size_t)dimV, Cell());
        trace.assign((size_t)(dimH * dimV), 0);
        if (!p.banded) {
            // score nav :72-81, trace nav :72-84
            sAct = 0; sPrev = 0; sLeap = 1 - dimV;
            tpos = 0; tLeap = 1;
        } else if (p.lower >= 0) {
            sLeap = 0; sAct = dimV - 1;
            tLeap = std::min(dimV, p.upper - p.lower + 1); tpos = dimV - 1;
        } else if (p.upper <= 0) {
            sLeap = 1 - dimV; sAct = 0;
            tLeap = 1; tpos = 0;
        } else {
            sLeap = std::max(p.lower, 1 - dimV);
            sAct = dimV + sLeap - 1;
            long lastPos = std::max(-(dimV - 1), p.lower);
            tLeap = dimV + lastPos;
            tpos = tLeap - 1;
        }
        sPrev = sAct;
        col[sAct] = Cell();
        trace[tpos] = 0;
    }

    void goNext(ColProp cp, ColLoc cl, CellT ct) {
        // ---- score navigator (sparse) :113-366
        if (ct == FIRST) {
            if (cp == INITIAL) {
            } else if (cl == TOP) {
                --sLeap; sAct += sLeap; sPrev = sAct; pH = col[++sPrev];
            } else if (cl == FULL) {
                sAct += sLeap; pH = col[sAct];
            } else {
                sAct += sLeap; sPrev = sAct; pD = col[sPrev]; pH = col[++sPrev];
            }
        } else if (ct == INNERC) {
            if (cp == INITIAL) {
                pV = col[sAct]; ++sAct;
            } else if (cl == FULL) {
                pD = pH; pV = col[sAct]; pH = col[++sAct];
            } else {
                pD = pH; pV = col[sAct]; pH = col[++sPrev]; ++sAct;
            }
        } else {
            if (cp == INITIAL) {
                pV = col[sAct]; ++sAct;
            } else if (cl == BOTTOM) {
                pD = pH; pV = col[sAct]; pH = col[++sPrev]; ++sAct; ++sLeap;
            } else if (cl == FULL) {
                pD = pH; pV = col[sAct]; pH = col[++sAct];
            } else {
                pD = pH; pV = col[sAct]; ++sAct;
            }
        }
        // ---- trace navigator :118-196
        if (ct == FIRST) {
            if (cp == INITIAL) {
            } else if (cl == TOP) {
                --tLeap; tpos += tLeap;
            } else
                tpos += tLeap;
        } else if (ct == INNERC) {
            ++tpos;
        } else {
            ++tpos;
            if (cp != INITIAL && cl == BOTTOM) ++tLeap;
        }
    }

    // ------------------------------------------------------------------ recurrences
    int sub(uint8_t a, uint8_t b) const { return a == b ? p.sc.match : p.sc.mismatch; }  // score/score_base.h:335-340

    // align/dp_formula_affine.h:205-262 (RecursionDirectionHorizontal helper)
    uint8_t affH(Cell& a, int cmp) const {
        if (a.h < cmp) { a.s = a.h = cmp; return T_HO; }
        a.s = a.h;
        if (p.complete && a.h == cmp) return T_H | T_HO;
        return T_H;
    }
    // :302-359
    uint8_t affV(Cell& a, int cmp) const {
        if (a.v < cmp) { a.s = a.v = cmp; return T_VO; }
        a.s = a.v;
        if (p.complete && a.v == cmp) return T_V | T_VO;
        return T_V;
    }
    // :398-456
    uint8_t affMax(Cell& a) const {
        if (a.s < a.h) { a.s = a.h; return T_MH; }
        if (p.complete && a.s == a.h) return T_MV | T_MH;
        return T_MV;
    }
    // :91-152
    uint8_t affD(Cell& a, int cmp, uint8_t left, uint8_t gap) const {
        if (!p.complete) {
            if (a.s <= cmp) { a.s = cmp; return T_D | left; }
            return left | gap;
        }
        if (a.s < cmp) { a.s = cmp; return T_D | left; }
        if (a.s == cmp) return left | T_D | gap;
        return left | gap;
    }
    // align/dp_formula_linear.h:96-148
    uint8_t lin(Cell& a, int cmp, uint8_t left, uint8_t right) const {
        if (a.s < cmp) { a.s = cmp; return right; }
        if (p.complete && a.s == cmp) return right | left;
        return left;
    }

    uint8_t computeScore(Cell& a, uint8_t hv, uint8_t vv, Rec rec) const {
        const int go = p.sc.gapOpen, ge = p.sc.gapExtend;
        if (rec == R_ZERO) { a.s = 0; return T_NONE; }  // align/dp_formula.h:199-214
        if (affine) {
            switch (rec) {
            case R_ALL: {  // align/dp_formula_affine.h:459-496
                a.h = pH.h + ge;
                uint8_t tvGap = affH(a, pH.s + go);
                a.v = pV.v + ge;
                tvGap |= affV(a, pV.s + go);
                uint8_t tvMax = affMax(a);
                return affD(a, pD.s + sub(hv, vv), tvGap, tvMax);
            }
            case R_UD: {  // :503-531
                a.h = pH.h + ge;
                a.v = NEG;
                uint8_t tv = affH(a, pH.s + go);
                return affD(a, pD.s + sub(hv, vv), tv, T_MH);
            }
            case R_LD: {  // :538-566
                a.v = pV.v + ge;
                int t = pV.s + go;
                a.h = NEG;
                uint8_t tv = affV(a, t);
                return affD(a, pD.s + sub(hv, vv), tv, T_MV);
            }
            case R_H: {  // :573-594
                int t = pH.s + go;
                a.h = pH.h + ge;
                a.v = NEG;
                return affH(a, t) | T_MH;
            }
            case R_V: {  // :601-622
                int t = pV.s + go;
                a.v = pV.v + ge;
                a.h = NEG;
                return affV(a, t) | T_MV;
            }
            default: break;
            }
        } else {
            switch (rec) {
            case R_ALL: {  // align/dp_formula_linear.h:150-185
                a.s = pD.s + sub(hv, vv);
                uint8_t tv = lin(a, pV.s + ge, T_D, T_V | T_MV);
                return lin(a, pH.s + ge, tv, T_H | T_MH);
            }
            case R_UD:  // :192-213
                a.s = pD.s + sub(hv, vv);
                return lin(a, pH.s + ge, T_D, T_H | T_MH);
            case R_LD:  // :220-241
                a.s = pD.s + sub(hv, vv);
                return lin(a, pV.s + ge, T_D, T_V | T_MV);
            case R_H:  // :248-266
                a.s = pH.s + ge;
                return T_H | T_MH;
            case R_V:  // :273-291
                a.s = pV.s + ge;
                return T_V | T_MV;
            default: break;
            }
        }
        return T_NONE;
    }

    // ------------------------------------------------------------------ meta info
    // align/dp_meta_info.h:96-260 (default DPMetaColumn_)
    Rec globalRec(ColProp cp, ColLoc cl, CellT ct) const {
        Rec firstColRec = p.fe.firstCol ? R_ZERO : R_V;
        if (ct == FIRST) {
            if (cl == FULL || cl == TOP) return (cp == INITIAL || p.fe.firstRow) ? R_ZERO : R_H;
            return cp == INITIAL ? R_ZERO : R_UD;
        }
        if (cp == INITIAL) return firstColRec;
        if (ct == INNERC) return R_ALL;
        return (cl == TOP || cl == MIDDLE) ? R_LD : R_ALL;
    }
    bool globalTrack(ColProp cp, ColLoc cl, CellT ct) const {
        if (ct == LAST && (cl == FULL || cl == BOTTOM)) return cp == FINAL || p.fe.lastRow;
        return cp == FINAL && p.fe.lastCol;
    }
    // seeds/banded_chain_alignment_profile.h:66-121
    Rec chainRec(ColLoc cl, CellT ct) const {
        if (ct == FIRST) return (cl == FULL || cl == TOP) ? R_ZERO : R_UD;
        if (ct == INNERC) return R_ALL;
        return (cl == TOP || cl == MIDDLE) ? R_LD : R_ALL;
    }

    // ------------------------------------------------------------------ scouts
    void scoutDefault(const Cell& a) {  // align/dp_scout.h:163-179
        if (a.s > maxCell.s) { maxCell = a; maxPos = (unsigned)tpos; }
    }
    // seeds/banded_chain_alignment_impl.h:282-377 + scout.h:230-270
    void chainTracking(ColProp cp, ColLoc cl, CellT ct, const Cell& a) {
        ChainState& st = *p.st;
        bool lastCol = false, lastRow = false, storeCol = false, storeRow = false;
        long ch = coordH(), cv = coordV();
        if (ch >= (long)st.hNext) {
            if (cl == BOTTOM) {
                if (cv + tLeap == (long)st.vNext) storeRow = true;
            } else {
                if (cv == (long)st.vNext) storeRow = true;
            }
            if (ch == (long)st.hNext)
                if (cv >= (long)st.vNext) storeCol = true;
            if (ct == LAST) {
                if (p.loc == LOC_FINAL) {
                    if (p.fe.lastRow) lastRow = true;
                } else
                    lastRow = true;
            }
            if (cl == FULL) {
                if (cp == FINAL) {
                    if (ct == LAST) {
                        lastCol = lastRow = true;
                    } else if (cv >= (long)st.vNext) {
                        if (p.loc == LOC_FINAL) {
                            if (p.fe.lastCol) lastCol = true;
                        } else
                            lastCol = true;
                    }
                }
            } else {
                if (cp == FINAL) {
                    if (ct == LAST) {
                        lastCol = lastRow = true;
                    } else {
                        if (p.loc == LOC_FINAL) {
                            if (p.fe.lastCol) lastCol = true;
                        } else
                            lastCol = true;
                    }
                }
            }
        }
        if (storeCol) st.vInitNext.at((size_t)(cv - (long)st.vNext)) = a;
        if (storeRow) st.hInitNext.at((size_t)(ch - (long)st.hNext)) = a;
        if (lastCol || lastRow) {
            if (a.s >= maxCell.s) {
                if (a.s == maxCell.s)
                    maxPositions.push_back((unsigned)tpos);
                else {
                    maxPositions.resize(1);
                    maxPositions[0] = (unsigned)tpos;
                    maxCell = a;
                }
            }
        }
    }

    // ------------------------------------------------------------------ cell / track
    // align/dp_algorithm_impl.h:279-311 and seeds/banded_chain_alignment_impl.h:405-560
    long hookRow = 0;  // matrix row of the cell being computed (debug hook only)
    void computeCell(ColProp cp, ColLoc cl, CellT ct, uint8_t hv, uint8_t vv) {
        Cell& a = col[sAct];
        if (g_cellHook) g_cellHook((int)coordH(), (int)hookRow, tpos, tLeap, (int)cp, (int)cl, (int)ct, (int)dimV);
        if (p.algo == ALGO_CHAIN) {
            uint8_t tv;
            if (cp == INITIAL) {
                a = p.st->vInitCur.at((size_t)(coordV() - (tLeap - 1)));  // impl.h:243-251
                tv = T_NONE;
            } else if (ct == FIRST && (cl == TOP || cl == FULL)) {
                a = p.st->hInitCur.at((size_t)coordH());  // impl.h:253-259
                tv = T_NONE;
            } else
                tv = computeScore(a, hv, vv, chainRec(cl, ct));
            trace[tpos] = tv;
            chainTracking(cp, cl, ct, a);
        } else {
            trace[tpos] = computeScore(a, hv, vv, globalRec(cp, cl, ct));
            if (globalTrack(cp, cl, ct)) scoutDefault(a);
        }
    }

    // align/dp_algorithm_impl.h:330-396
    void computeTrack(ColProp cp, ColLoc cl, uint8_t hv, uint8_t vFirst, long vBegin, long vEnd) {
        goNext(cp, cl, FIRST);
        hookRow = (cl == FULL || cl == TOP) ? 0 : vBegin;
        computeCell(cp, cl, FIRST, hv, vFirst);
        long it = vBegin;
        long itEnd = vEnd - 1;
        for (; it != itEnd; ++it) {
            goNext(cp, cl, INNERC);
            hookRow = it + 1;
            computeCell(cp, cl, INNERC, hv, p.V[it]);
        }
        goNext(cp, cl, LAST);
        hookRow = it + 1;
        computeCell(cp, cl, LAST, hv, p.V[it]);
    }

    // align/dp_algorithm_impl.h:424-513
    void computeUnbanded() {
        computeTrack(INITIAL, FULL, p.H[0], p.V[0], 0, p.nV);
        long h = 0;
        for (; h != p.nH - 1; ++h) computeTrack(INNER, FULL, p.H[h], p.V[0], 0, p.nV);
        computeTrack(FINAL, FULL, p.H[h], p.V[0], 0, p.nV);
    }

    // Extra scouting calls of the default scout inside _computeBandedAlignment.  For the
    // banded-chain scout these resolve to the empty overload (scout.h:272-283).
    void extraScout(bool enabled) {
        if (p.algo == ALGO_GLOBAL && enabled) scoutDefault(col[sAct]);
    }

    // align/dp_algorithm_impl.h:515-860
    void computeBanded() {
        const long nH = p.nH, nV = p.nV, lo = p.lower, up = p.upper;
        long vBegin = 0 - std::min(0L, 1 + up);
        long vEnd = 0 - std::min(0L, std::max(-nV, lo));
        long hBegin = std::max(0L, std::min(nH - 1, lo));
        long hEndTop = std::min(nH - 1, std::max(0L, up));
        long hEndMid = std::min(nH - 1, std::max(0L, nV + lo));
        if (up > nV + lo) std::swap(hEndTop, hEndMid);
        long hEndBottom = std::max(0L, std::min(nH, up + nV) - 1);

        if (hBegin == nH - 1) {  // :545-558
            goNext(INITIAL, TOP, FIRST);
            hookRow = 0;
            computeCell(INITIAL, TOP, FIRST, p.H[hBegin], p.V[0]);
            extraScout(globalTrack(FINAL, TOP, FIRST));
            return;
        }
        if (hEndBottom == 0) {  // :559-573
            goNext(INITIAL, BOTTOM, FIRST);
            hookRow = vBegin;
            computeCell(INITIAL, BOTTOM, FIRST, p.H[0], p.V[vBegin]);
            extraScout(globalTrack(INITIAL, BOTTOM, LAST));
            return;
        }
        if (up < 0) {  // :574-590
            ++vBegin;
            if (lo > -nV)
                computeTrack(INITIAL, MIDDLE, p.H[0], p.V[vBegin - 1], vBegin, vEnd);
            else
                computeTrack(INITIAL, BOTTOM, p.H[0], p.V[vBegin - 1], vBegin, vEnd);
        } else if (lo >= 0) {  // :591-604
            goNext(INITIAL, TOP, FIRST);
            computeCell(INITIAL, TOP, FIRST, p.H[hBegin], p.V[0]);
            extraScout(globalTrack(INNER, TOP, FIRST));
        } else if (lo <= -nV)  // :606-611
            computeTrack(INITIAL, FULL, p.H[0], p.V[0], vBegin, vEnd);
        else  // :612-617
            computeTrack(INITIAL, TOP, p.H[0], p.V[0], vBegin, vEnd);

        long h = hBegin;
        for (; h != hEndTop; ++h) {  // :623-635
            ++vEnd;
            computeTrack(INNER, TOP, p.H[h], p.V[0], vBegin, vEnd);
        }
        if (up > nV + lo) {  // :636-652
            extraScout(globalTrack(INNER, FULL, LAST));
            for (; h != hEndMid; ++h) computeTrack(INNER, FULL, p.H[h], p.V[0], vBegin, vEnd);
        } else {  // :653-678
            for (; h != hEndMid; ++h) {
                ++vBegin;
                ++vEnd;
                computeTrack(INNER, MIDDLE, p.H[h], p.V[vBegin - 1], vBegin, vEnd);
            }
            if (globalTrack(INNER, BOTTOM, LAST))
                if (lo + nV < nH) extraScout(true);
        }
        for (; h != hEndBottom; ++h) {  // :679-691
            ++vBegin;
            computeTrack(INNER, BOTTOM, p.H[h], p.V[vBegin - 1], vBegin, vEnd);
        }
        if (h < nH - 1) {  // Case 1 :692-706
            goNext(INNER, BOTTOM, FIRST);
            hookRow = vBegin + 1;
            computeCell(INNER, BOTTOM, FIRST, p.H[h], p.V[vBegin]);
            extraScout(globalTrack(INNER, BOTTOM, LAST));
        } else if (h == nH - 1) {  // Case 2 :707-
            if (up == nH - nV) {
                goNext(FINAL, BOTTOM, FIRST);
                hookRow = vBegin + 1;
                computeCell(FINAL, BOTTOM, FIRST, p.H[h], p.V[vBegin]);
                extraScout(globalTrack(FINAL, BOTTOM, LAST));
            } else {
                if (up >= nH) {
                    if (lo + nV > nH) {
                        ++vEnd;
                        computeTrack(FINAL, TOP, p.H[h], p.V[0], vBegin, vEnd);
                    } else {
                        if (lo + nV + 1 > nH) {
                            ++vEnd;
                            computeTrack(FINAL, TOP, p.H[h], p.V[0], vBegin, vEnd);
                            extraScout(globalTrack(FINAL, FULL, LAST));
                        } else
                            computeTrack(FINAL, FULL, p.H[h], p.V[0], vBegin, vEnd);
                    }
                } else {
                    ++vBegin;
                    if (lo + nV <= nH) {
                        if (lo + nV == nH) {
                            ++vEnd;
                            computeTrack(FINAL, MIDDLE, p.H[h], p.V[vBegin - 1], vBegin, vEnd);
                            extraScout(globalTrack(FINAL, BOTTOM, LAST));
                        } else
                            computeTrack(FINAL, BOTTOM, p.H[h], p.V[vBegin - 1], vBegin, vEnd);
                    } else {
                        ++vEnd;
                        computeTrack(FINAL, MIDDLE, p.H[h], p.V[vBegin - 1], vBegin, vEnd);
                    }
                }
            }
        }
    }

    // align/dp_algorithm_impl.h:117-157
    bool validSettings() const {
        if (p.nH == 0 || p.nV == 0) return false;
        if (!p.banded) return true;
        if (p.upper < -p.nV || p.lower > p.nH) return false;
        if (p.upper < 0 && !p.fe.firstCol) return false;
        if (p.lower > 0 && !p.fe.firstRow) return false;
        if (p.upper + p.nV < p.nH && !p.fe.lastRow) return false;
        if (p.lower + p.nV > p.nH && !p.fe.lastCol) return false;
        return true;
    }

    // ------------------------------------------------------------------ traceback
    Coord makeCoord(long currCol, long currRow, long endCol, long endRow) const {  // :121-141
        Coord c{currCol, currRow, endCol, endRow, 0, 0, false};
        if (p.banded) {
            if (p.lower >= 0) c.currCol += p.lower;
            if (c.currCol > p.upper) c.currRow += c.currCol - p.upper;
            if (c.endCol > p.upper) c.endRow += c.endCol - p.upper;
            c.bp1 = std::min(p.nH, std::max(0L, p.upper));
            c.bp2 = std::min(p.nH, std::max(0L, p.nV + p.lower));
            if (c.currCol < std::min(c.bp1, c.bp2)) c.currRow -= std::min(c.bp1, c.bp2) - c.currCol;
            c.inBandFlag = true;
        }
        return c;
    }

    static void record(Trace& t, long h, long v, long len, uint8_t tv) {  // align/dp_trace_segment.h:319-337
        if (len == 0) return;
        if (tv & T_D) t.push_back(Seg{h, v, len, T_D});
        else if (tv & T_V) t.push_back(Seg{h, v, len, T_V});
        else if (tv & T_H) t.push_back(Seg{h, v, len, T_H});
    }

    void traceH(const Coord& c) { tpos -= c.isInBand() ? dimV - 1 : dimV; }       // nav :198-209
    void traceD(const Coord& c) { tpos -= c.isInBand() ? dimV : dimV + 1; }       // nav :211-222
    void traceV(const Coord&) { tpos -= 1; }                                      // nav :224-232

    // align/dp_traceback_impl.h:152-431 (GapsLeft only — that is what every caller uses)
    void doTraceback(Trace& target, uint8_t& tv, uint8_t& last, long& frag, Coord& c) {
        if (tv & T_D) {
            if (!(last & T_D)) { record(target, c.currCol, c.currRow, frag, last); last = T_D; frag = 0; }
            traceD(c); tv = trace[tpos]; --c.currCol; --c.currRow; ++frag;
        } else if ((tv & T_MV) && (tv & T_V)) {
            if (!(last & T_V)) { record(target, c.currCol, c.currRow, frag, last); last = T_V; frag = 0; }
            if (affine) {
                while ((!(tv & T_VO) || (tv & T_V)) && c.currRow != 1) {
                    traceV(c); tv = trace[tpos]; --c.currRow; ++frag;
                }
                traceV(c); tv = trace[tpos]; --c.currRow; ++frag;
            } else {
                traceV(c); tv = trace[tpos]; --c.currRow; ++frag;
            }
        } else if ((tv & T_MV) && (tv & T_VO)) {
            if (!(last & T_V)) { record(target, c.currCol, c.currRow, frag, last); last = T_V; frag = 0; }
            traceV(c); tv = trace[tpos]; --c.currRow; ++frag;
        } else if ((tv & T_MH) && (tv & T_H)) {
            if (!(last & T_H)) { record(target, c.currCol, c.currRow, frag, last); last = T_H; frag = 0; }
            if (affine) {
                while ((!(tv & T_HO) || (tv & T_H)) && c.currCol != 1) {
                    traceH(c); tv = trace[tpos]; --c.currCol; ++frag;
                }
                traceH(c); tv = trace[tpos]; --c.currCol; ++frag;
            } else {
                traceH(c); tv = trace[tpos]; --c.currCol; ++frag;
            }
        } else if ((tv & T_MH) && (tv & T_HO)) {
            if (!(last & T_H)) { record(target, c.currCol, c.currRow, frag, last); last = T_H; frag = 0; }
            traceH(c); tv = trace[tpos]; --c.currCol; ++frag;
        } else {
            // NONE: caller's loop condition ends the walk.  Any other value is the
            // reference's SEQAN_ASSERT_FAIL (compiled out with NDEBUG) — an endless loop there.
            if (tv != T_NONE) throw std::logic_error("undefined traceback value");
        }
    }

    static uint8_t initialDirection(uint8_t& tv, bool preferGapsAtEnd) {  // :433-461
        if (preferGapsAtEnd) {
            if (tv & T_MV) { tv &= (T_V | T_VO | T_MV); return T_V; }
            if (tv & T_MH) { tv &= (T_H | T_HO | T_MH); return T_H; }
            return T_D;
        }
        if (tv & T_D) return T_D;
        if (tv & (T_V | T_MV)) return T_V;
        if (tv & (T_H | T_MH)) return T_H;
        return T_NONE;
    }

    // generic _computeTraceback, align/dp_traceback_impl.h:463-526
    void tracebackGeneric(Trace& target, unsigned startPos, bool preferGapsAtEnd, bool head, bool tail) {
        tpos = startPos;
        uint8_t tv = trace[tpos];
        uint8_t last = initialDirection(tv, preferGapsAtEnd);
        Coord c = makeCoord(coordH(), coordV(), 0, 0);
        if (tail) {
            if (c.currRow != p.nV) record(target, p.nH, c.currRow, p.nV - c.currRow, T_V);
            if (c.currCol != p.nH) record(target, c.currCol, c.currRow, p.nH - c.currCol, T_H);
        }
        long frag = 0;
        while (!c.reachedEnd() && tv != T_NONE) doTraceback(target, tv, last, frag, c);
        record(target, c.currCol, c.currRow, frag, last);
        if (head) {
            if (c.currRow != 0) record(target, 0, 0, c.currRow, T_V);
            if (c.currCol != 0) record(target, 0, 0, c.currCol, T_H);
        }
    }

    // seeds/banded_chain_alignment_traceback.h:233-355 (one candidate)
    void tracebackChainOne(Trace& target, unsigned startPos) {
        ChainState& st = *p.st;
        const bool prefer = affine && p.loc == LOC_FINAL;  // traceback.h:55-61; linear+GapsLeft -> False
        tpos = startPos;
        uint8_t tv = trace[tpos];
        uint8_t last = initialDirection(tv, prefer);
        Coord c = makeCoord(coordH(), coordV(), (long)st.hNext, (long)st.vNext);
        if (p.loc == LOC_FINAL) {
            if (c.currRow != p.nV) record(target, p.nH, c.currRow, p.nV - c.currRow, T_V);
            if (c.currCol != p.nH) record(target, c.currCol, c.currRow, p.nH - c.currCol, T_H);
            tracebackGeneric(target, (unsigned)tpos, prefer, false, false);
            return;
        }
        long frag = 0;
        Trace tmp;
        while (!c.reachedEnd() && tv != T_NONE) doTraceback(tmp, tv, last, frag, c);
        long hInit = c.currCol - c.endCol;
        long vInit = c.currRow - c.endRow;
        bool inserted;
        auto correct = [&](Cell& cell) {  // traceback.h:211-231
            if (!affine) return;
            if (last & T_D) { cell.v = NEG; cell.h = NEG; }
            else if (last & T_V) cell.h = NEG;
            else cell.v = NEG;
        };
        if (vInit <= 0) {
            Cell& cell = st.hInitNext.at((size_t)hInit);
            correct(cell);
            inserted = st.nextInitCells.insert(InitCell{(unsigned)hInit, 0u, cell, affine}).second;
        } else {
            Cell& cell = st.vInitNext.at((size_t)vInit);
            correct(cell);
            inserted = st.nextInitCells.insert(InitCell{0u, (unsigned)vInit, cell, affine}).second;
        }
        if (inserted) {
            if (vInit < 0) record(target, c.currCol, c.currRow, -vInit, last);
            else if (hInit < 0) record(target, c.currCol, c.currRow, -hInit, last);
            tracebackGeneric(target, (unsigned)tpos, prefer, false, false);
        }
        if (p.loc == LOC_INITIAL) {
            long currCol = coordH(), currRow = coordV();
            if (p.banded)
                if (p.upper > 0)
                    if (currCol < c.bp1)
                        if (currCol < c.bp2) currRow -= dimV - 1 + p.lower - currCol;
            if (currRow != 0) record(target, 0, 0, currRow, T_V);
            if (currCol != 0) record(target, 0, 0, currCol, T_H);
        }
    }
};

void countCells(CellCounter* cc, const DPRun& r) {
    if (cc) { cc->cells += (long long)r.dimH * r.dimV; cc->grids += 1; }
}

// _computeAlignment for the default (global) profile.  Returns score; throws BadScore.
int runGlobal(const DPProblem& prob, Trace& out, CellCounter* cc) {
    if (g_gridHook) g_gridHook(3, prob.nH, prob.nV, prob.banded ? 1 : 0, prob.lower, prob.upper, prob.st ? (long)prob.st->hNext : 0, prob.st ? (long)prob.st->vNext : 0);
    DPRun r(prob);
    if (!r.validSettings()) return INT_MIN;  // :1543-1544 — no traceback, no throw
    r.initNavigators();
    countCells(cc, r);
    if (!prob.banded) r.computeUnbanded();
    else if (prob.upper == prob.lower) throw std::logic_error("band width 1 (_computeHammingDistance) not restated");
    else r.computeBanded();
    if (r.maxCell.s < -1000000) throw BadScore();
    if (!prob.complete && r.affine) {  // _correctTraceValue :1354-1370
        uint8_t& t = r.trace.at(r.maxPos);
        if (r.maxCell.v == r.maxCell.s) { t &= ~T_D; t |= T_MV; }
        else if (r.maxCell.h == r.maxCell.s) { t &= ~T_D; t |= T_MH; }
    }
    // PreferGapsAtEnd_: affine -> True, linear+GapsLeft -> False (align/dp_traceback_impl.h:98-105)
    r.tracebackGeneric(out, r.maxPos, r.affine, true, true);
    return r.maxCell.s;
}

// _computeAlignment for BandedChainAlignment_ profiles: fills, then one traceback per
// tied maximum (seeds/banded_chain_alignment_traceback.h:357-388).
int runChainGrid(const DPProblem& prob, std::vector<Trace>& localTraces, CellCounter* cc) {
    if (g_gridHook) g_gridHook((int)prob.loc, prob.nH, prob.nV, prob.banded ? 1 : 0, prob.lower, prob.upper, (long)prob.st->hNext, (long)prob.st->vNext);
    DPRun r(prob);
    if (!r.validSettings()) return INT_MIN;
    r.initNavigators();
    countCells(cc, r);
    if (!prob.banded) r.computeUnbanded();
    else if (prob.upper == prob.lower) throw std::logic_error("band width 1 not restated");
    else r.computeBanded();
    if (r.maxCell.s < -1000000) throw BadScore();
    prob.st->nextInitCells.clear();
    for (size_t i = 0; i < r.maxPositions.size(); ++i) {
        Trace tmp;
        r.tracebackChainOne(tmp, r.maxPositions[i]);
        if (!tmp.empty()) localTraces.push_back(tmp);
    }
    return r.maxCell.s;
}

// seeds/banded_chain_alignment_scout.h:175-221
void reinitScoutState(ChainState& st, long originH, long originV, long sizeCurH, long sizeCurV, long sizeNextH,
                      long sizeNextV) {
    st.hNext = (unsigned)originH;
    st.vNext = (unsigned)originV;
    std::fill(st.hInitCur.begin(), st.hInitCur.end(), Cell());
    std::fill(st.vInitCur.begin(), st.vInitCur.end(), Cell());
    std::fill(st.hInitNext.begin(), st.hInitNext.end(), Cell());
    std::fill(st.vInitNext.begin(), st.vInitNext.end(), Cell());
    if ((long)st.hInitCur.size() < sizeCurH) st.hInitCur.resize((size_t)sizeCurH, Cell());
    if ((long)st.vInitCur.size() < sizeCurV) st.vInitCur.resize((size_t)sizeCurV, Cell());
    if ((long)st.hInitNext.size() < sizeNextH) st.hInitNext.resize((size_t)sizeNextH, Cell());
    if ((long)st.vInitNext.size() < sizeNextV) st.vInitNext.resize((size_t)sizeNextV, Cell());
    for (const InitCell& ic : st.nextInitCells) {
        if (ic.i1 == 0) st.vInitCur.at(ic.i2) = ic.c;
        if (ic.i2 == 0) st.hInitCur.at(ic.i1) = ic.c;
    }
}

void adaptLocal(std::vector<Trace>& ts, long h0, long v0) {  // traceback.h:63-78
    for (Trace& t : ts)
        for (Seg& s : t) { s.hBeg += h0; s.vBeg += v0; }
}

void smoothGluePoint(Trace& path, size_t referenceSize) {  // traceback.h:80-94
    size_t endOld = path.size() - referenceSize;
    size_t beginNew = endOld - 1;
    if (path[endOld].dir == path[beginNew].dir) {
        path[endOld].len += path[beginNew].len;
        path.erase(path.begin() + (long)beginNew);
    }
}

void glueTracebacks(std::vector<Trace>& global, std::vector<Trace>& local) {  // traceback.h:96-203
    if (global.empty()) { global = local; return; }
    size_t lengthGlobal = global.size();
    size_t oldNum = lengthGlobal;
    std::vector<size_t> toErase;
    for (size_t j = 0; j < lengthGlobal; ++j) {
        Seg gEnd = global[j].front();
        size_t numCurr = global[j].size();
        size_t numAdded = 0;
        bool connected = false;
        for (size_t i = 0; i < local.size(); ++i) {
            const Seg& lBeg = local[i].back();
            if (gEnd.hEnd() == lBeg.hBeg && gEnd.vEnd() == lBeg.vBeg) {
                if (connected) {
                    Trace nt(local[i]);
                    nt.insert(nt.end(), global[j].end() - (long)numCurr, global[j].end());
                    global.push_back(nt);
                    ++numAdded;
                    continue;
                }
                Trace ng(local[i]);
                ng.insert(ng.end(), global[j].begin(), global[j].end());
                global[j].swap(ng);
                connected = true;
            }
        }
        if (!connected)
            toErase.push_back(j);
        else {
            smoothGluePoint(global[j], numCurr);
            for (size_t t = oldNum; t < oldNum + numAdded; ++t) smoothGluePoint(global[t], numCurr);
            oldNum += numAdded;
        }
    }
    for (size_t i = toErase.size(); i > 0; --i) global.erase(global.begin() + (long)toErase[i - 1]);
}

struct ChainCtx {
    const std::vector<uint8_t>& H;
    const std::vector<uint8_t>& V;
    Score sc;
    FreeEnds fe;
    bool affine;
    long b;  // bandExtension
    ChainState st;
    std::vector<Trace> global;
    CellCounter* cc;

    long lenH() const { return (long)H.size(); }
    long lenV() const { return (long)V.size(); }

    DPProblem mk(long h0, long h1, long v0, long v1, bool banded, long lo, long up, Algo algo, MatLoc loc) {
        DPProblem p;
        p.H = H.data() + h0; p.nH = h1 - h0;
        p.V = V.data() + v0; p.nV = v1 - v0;
        p.sc = sc; p.banded = banded; p.lower = lo; p.upper = up;
        p.complete = true; p.algo = algo; p.fe = fe; p.loc = loc; p.st = &st;
        return p;
    }

    static long hShiftBegin(const Seed& s) { return s.upperDiag - (s.beginH - s.beginV); }  // impl.h:197-205
    static long vShiftBegin(const Seed& s) { return (s.beginH - s.beginV) - s.lowerDiag; }  // :207-218
    static long hShiftEnd(const Seed& s) { return s.endH - s.endV - s.lowerDiag; }          // :225-233
    static long vShiftEnd(const Seed& s) { return s.upperDiag - s.endH + s.endV; }          // :235-246

    // impl.h:683-729
    void initBeginning(long sizeH, long sizeV) {
        const int go = sc.gapOpen, ge = sc.gapExtend;
        Cell hc;
        hc.s = 0;  // RecursionDirectionZero on a default cell
        st.nextInitCells.insert(InitCell{0u, 0u, hc, affine});
        for (long colI = 1; colI < sizeH; ++colI) {
            Cell prev = hc;
            if (fe.firstRow) hc.s = 0;
            else if (affine) {
                int t = prev.s + go; hc.h = prev.h + ge; hc.v = NEG;
                if (hc.h < t) hc.s = hc.h = t; else hc.s = hc.h;
            } else hc.s = prev.s + ge;
            st.nextInitCells.insert(InitCell{(unsigned)colI, 0u, hc, affine});
        }
        Cell vc;
        vc.s = 0;
        for (long rowI = 1; rowI < sizeV; ++rowI) {
            Cell prev = vc;
            if (fe.firstCol) vc.s = 0;
            else if (affine) {
                int t = prev.s + go; vc.v = prev.v + ge; vc.h = NEG;
                if (vc.v < t) vc.s = vc.v = t; else vc.s = vc.v;
            } else vc.s = prev.s + ge;
            st.nextInitCells.insert(InitCell{0u, (unsigned)rowI, vc, affine});
        }
    }

    void finishLocal(std::vector<Trace>& local, long h0, long v0) {
        adaptLocal(local, h0, v0);
        if (!local.empty()) glueTracebacks(global, local);
    }

    // impl.h:737-894
    int initializeChain(const Seed& seed) {
        long hShift = hShiftBegin(seed), vShift = vShiftBegin(seed);
        long hNextO = std::max(0L, seed.beginH + 1 - b);
        long vNextO = std::max(0L, seed.beginV + 1 - b);
        long up = std::min(lenH(), hNextO + (b << 1) + hShift + std::max(0L, b - seed.beginV - 1) +
                                       std::min(0L, seed.beginH + 1 - b));
        long lo = -std::min(lenV(), vNextO + (b << 1) + vShift + std::max(0L, b - seed.beginH - 1) +
                                        std::min(0L, seed.beginV + 1 - b));
        int score = 0;
        if (hNextO != 0 || vNextO != 0) {
            initBeginning(up + 1, 1 - lo);
            reinitScoutState(st, hNextO, vNextO, 1 + up, 1 - lo, 1 + up - hNextO, 1 - lo - vNextO);
            DPProblem p = mk(0, up, 0, -lo, false, 0, 0, ALGO_CHAIN, LOC_INITIAL);
            score = runChainGrid(p, global, cc);
        } else
            initBeginning(up, -lo);
        long gb1 = hNextO, gb2 = vNextO;
        long ge1 = std::min(lenH(), seed.endH + b);
        long ge2 = std::min(lenV(), seed.endV + b);
        long infH = ge1 - gb1, infV = ge2 - gb2;
        hShift = hShiftEnd(seed);
        vShift = vShiftEnd(seed);
        hNextO = std::max(0L, seed.endH - b - hShift - std::max(0L, seed.endV + b - lenV()) - gb1);
        vNextO = std::max(0L, seed.endV - b - vShift - std::max(0L, seed.endH + b - lenH()) - gb2);
        up = up - gb1;
        lo = lo + gb2;
        if (infV + lo > up) vNextO -= (infV + lo) - up;
        reinitScoutState(st, hNextO, vNextO, up + 1, 1 - lo, infH - hNextO + 1, infV - vNextO + 1);
        std::vector<Trace> local;
        if (gb1 == 0 && gb2 == 0) {
            if (ge1 == lenH() && ge2 == lenV()) {
                local.resize(1);
                DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_GLOBAL, LOC_INNER);
                score = runGlobal(p, local[0], cc);
            } else {
                DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_CHAIN, LOC_INITIAL);
                score = runChainGrid(p, local, cc);
            }
        } else {
            if (ge1 == lenH() && ge2 == lenV()) {
                DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_CHAIN, LOC_FINAL);
                score = runChainGrid(p, local, cc);
            } else {
                DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_CHAIN, LOC_INNER);
                score = runChainGrid(p, local, cc);
            }
        }
        if (score < -1000000) throw BadScore();
        if (gb1 != 0 || gb2 != 0) {
            adaptLocal(local, gb1, gb2);
            if (!local.empty()) glueTracebacks(global, local);
        } else
            global = local;
        st.hNext += (unsigned)gb1;
        st.vNext += (unsigned)gb2;
        if (infV + lo > up) st.vNext += (unsigned)((infV + lo) - up);
        return score;
    }

    // impl.h:901-959
    int gapArea(const Seed& seed) {
        long gb1 = st.hNext, gb2 = st.vNext;
        long ge1 = seed.beginH + 1 + b + hShiftBegin(seed);
        long ge2 = seed.beginV + 1 + b + vShiftBegin(seed);
        long hNextO = seed.beginH + 1 - b - gb1;
        long vNextO = seed.beginV + 1 - b - gb2;
        reinitScoutState(st, hNextO, vNextO, ge1 - gb1 + 1, ge2 - gb2 + 1, ge1 - gb1 + 1 - hNextO,
                         ge2 - gb2 + 1 - vNextO);
        std::vector<Trace> local;
        DPProblem p = mk(gb1, ge1, gb2, ge2, false, 0, 0, ALGO_CHAIN, LOC_INNER);
        int score = runChainGrid(p, local, cc);
        if (score < -1000000) throw BadScore();
        finishLocal(local, gb1, gb2);
        st.hNext += (unsigned)gb1;
        st.vNext += (unsigned)gb2;
        return score;
    }

    // impl.h:966-1027
    int anchorArea(const Seed& seed) {
        long gb1 = st.hNext, gb2 = st.vNext;
        long ge1 = seed.endH + b, ge2 = seed.endV + b;
        long infH = ge1 - gb1, infV = ge2 - gb2;
        long hShift = hShiftBegin(seed), vShift = vShiftBegin(seed);
        long hNextO = seed.endH - b - hShiftEnd(seed) - gb1;
        long vNextO = seed.endV - b - vShiftEnd(seed) - gb2;
        long lo = -(b << 1) - vShift, up = (b << 1) + hShift;
        long relV = vNextO;
        if (infV + lo > up) relV -= (infV + lo) - up;
        reinitScoutState(st, hNextO, relV, up + 1, 1 - lo, infH - hNextO + 1, infV - vNextO + 1);
        std::vector<Trace> local;
        DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_CHAIN, LOC_INNER);
        int score = runChainGrid(p, local, cc);
        finishLocal(local, gb1, gb2);
        st.hNext += (unsigned)gb1;
        st.vNext += (unsigned)gb2;
        if (infV + lo > up) st.vNext += (unsigned)((infV + lo) - up);
        return score;
    }

    // impl.h:1038-1177
    int finishChain(const Seed& seed) {
        long gb1 = st.hNext, gb2 = st.vNext;
        long hShift = hShiftBegin(seed), vShift = vShiftBegin(seed);
        long ge1 = std::min(lenH(), seed.beginH + 1 + b + hShift);
        long ge2 = std::min(lenV(), seed.beginV + 1 + b + vShift);
        long infH = ge1 - gb1, infV = ge2 - gb2;
        long hNextO = std::max(0L, seed.beginH + 1 - b - gb1);
        long vNextO = std::max(0L, seed.beginV + 1 - b - gb2);
        reinitScoutState(st, hNextO, vNextO, infH + 1, infV + 1, infH - hNextO + 1, infV - vNextO + 1);
        std::vector<Trace> local;
        {
            DPProblem p = mk(gb1, ge1, gb2, ge2, false, 0, 0, ALGO_CHAIN, LOC_INNER);
            runChainGrid(p, local, cc);
        }
        finishLocal(local, gb1, gb2);
        gb1 += hNextO;
        gb2 += vNextO;
        ge1 = std::min(lenH(), seed.endH + b);
        ge2 = std::min(lenV(), seed.endV + b);
        infH = ge1 - gb1;
        infV = ge2 - gb2;
        if (ge1 == lenH() && ge2 == lenV()) {
            long lo = -(ge2 - gb2), up = ge1 - gb1;
            reinitScoutState(st, 0, 0, up + 1, 1 - lo, up + 1, 1 - lo);
            local.clear();
            DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_CHAIN, LOC_FINAL);
            int score = runChainGrid(p, local, cc);
            finishLocal(local, gb1, gb2);
            return score;
        }
        long lo = -(b << 1) - vShift, up = (b << 1) + hShift;
        hNextO = std::max(0L, seed.endH - b - hShiftEnd(seed) - gb1 - std::max(0L, seed.endV + b - lenV()));
        vNextO = std::max(0L, seed.endV - b - vShiftEnd(seed) - gb2 - std::max(0L, seed.endH + b - lenH()));
        if (infV + lo > up) vNextO -= (infV + lo) - up;
        reinitScoutState(st, hNextO, vNextO, up + 1, 1 - lo, infH - hNextO + 1, infV - vNextO + 1);
        local.clear();
        int score;
        {
            DPProblem p = mk(gb1, ge1, gb2, ge2, true, lo, up, ALGO_CHAIN, LOC_INNER);
            score = runChainGrid(p, local, cc);
        }
        finishLocal(local, gb1, gb2);
        gb1 += hNextO;
        if (infV + lo > up) vNextO += (infV + lo) - up;
        gb2 += vNextO;
        reinitScoutState(st, 0, 0, lenH() - gb1 + 1, lenV() - gb2 + 1, lenH() - gb1 + 1, lenV() - gb2 + 1);
        local.clear();
        {
            DPProblem p = mk(gb1, lenH(), gb2, lenV(), false, 0, 0, ALGO_CHAIN, LOC_FINAL);
            score = runChainGrid(p, local, cc);
        }
        if (score < -1000000) throw BadScore();
        finishLocal(local, gb1, gb2);
        return score;
    }

    // impl.h:1212-1296
    int run(const std::vector<Seed>& seeds) {
        if (seeds.empty()) return INT_MIN;
        // _findFirstAnchor :598-621
        size_t it = 0, last = seeds.size() - 1;
        {
            size_t i = 0;
            bool found = false;
            while (i != last) {
                const Seed& s = seeds[++i];
                if (s.beginH - b <= 0) continue;
                if (s.beginV - b <= 0) continue;
                it = i - 1;
                found = true;
                break;
            }
            if (!found) it = i;
        }
        // _findLastAnchor :623-647
        size_t itEnd;
        {
            size_t i = last;
            bool found = false;
            while (i != it) {
                const Seed& s = seeds[--i];
                if (s.endH + b >= lenH()) continue;
                if (s.endV + b >= lenV()) continue;
                found = true;
                break;
            }
            (void)found;
            itEnd = i;
        }
        int score = initializeChain(seeds[it]);
        if (seeds.size() == 1 || (it == itEnd && itEnd == last)) {
            if (seeds[it].endH + b < lenH() || seeds[it].endV + b < lenV()) {
                long gbH = st.hNext, gbV = st.vNext;
                reinitScoutState(st, 0, 0, lenH() + 1 - gbH, lenV() + 1 - gbV, lenH() + 1 - gbH, lenV() + 1 - gbV);
                std::vector<Trace> local;
                DPProblem p = mk(gbH, lenH(), gbV, lenV(), false, 0, 0, ALGO_CHAIN, LOC_FINAL);
                score = runChainGrid(p, local, cc);
                adaptLocal(local, gbH, gbV);
                glueTracebacks(global, local);
            }
            return score;
        }
        while (it != itEnd) {
            ++it;
            gapArea(seeds[it]);
            anchorArea(seeds[it]);
        }
        ++it;
        return finishChain(seeds[it]);
    }
};

}  // namespace

bool globalAlignmentTrace(const std::vector<uint8_t>& H, const std::vector<uint8_t>& V, const Score& sc,
                          const FreeEnds& fe, bool banded, long lower, long upper, Trace& out, int& score,
                          CellCounter* cc) {
    DPProblem p;
    p.H = H.data(); p.nH = (long)H.size();
    p.V = V.data(); p.nV = (long)V.size();
    p.sc = sc; p.banded = banded; p.lower = lower; p.upper = upper;
    p.complete = false; p.algo = ALGO_GLOBAL; p.fe = fe; p.loc = LOC_INNER; p.st = nullptr;
    try {
        score = runGlobal(p, out, cc);
    } catch (BadScore&) {
        return false;
    }
    return true;
}

bool bandedChainAlignmentTrace(const std::vector<uint8_t>& H, const std::vector<uint8_t>& V,
                               const std::vector<Seed>& chain, const Score& sc, const FreeEnds& fe,
                               unsigned bandExtension, Trace& out, bool& traceEmpty, int& score, CellCounter* cc) {
    ChainCtx ctx{H, V, sc, fe, sc.affine(), (long)bandExtension, ChainState(), {}, cc};
    try {
        score = ctx.run(chain);
    } catch (BadScore&) {
        return false;
    }
    traceEmpty = ctx.global.empty();
    if (!traceEmpty) out = ctx.global[0];
    return true;
}

std::vector<uint8_t> toDna5(const std::string& s) {
    std::vector<uint8_t> r(s.size());
    for (size_t i = 0; i < s.size(); ++i) {
        switch (s[i]) {
        case 'A': case 'a': r[i] = 0; break;
        case 'C': case 'c': r[i] = 1; break;
        case 'G': case 'g': r[i] = 2; break;
        case 'T': case 't': case 'U': case 'u': r[i] = 3; break;
        default: r[i] = 4;
        }
    }
    return r;
}

void traceToRows(const Trace& tr, const std::vector<uint8_t>& H, const std::vector<uint8_t>& V, std::string& rowH,
                 std::string& rowV) {
    static const char* A = "ACGTN";
    rowH.clear();
    rowV.clear();
    if (tr.empty()) return;
    long h = tr.back().hBeg, v = tr.back().vBeg;
    for (size_t k = tr.size(); k > 0; --k) {
        const Seg& s = tr[k - 1];
        for (long i = 0; i < s.len; ++i) {
            if (s.dir == T_H) { rowH.push_back(A[H.at((size_t)h++)]); rowV.push_back('-'); }
            else if (s.dir == T_V) { rowH.push_back('-'); rowV.push_back(A[V.at((size_t)v++)]); }
            else { rowH.push_back(A[H.at((size_t)h++)]); rowV.push_back(A[V.at((size_t)v++)]); }
        }
    }
}

std::string scoredAlignmentString(const std::string& readAlignment, const std::string& refAlignment,
                                  const std::string& readName, const std::string& refName, int refOffset,
                                  bool startImmediately, bool goToEndSeq1, bool goToEndSeq2, const Score& sc,
                                  double* scaledOut) {
    // unicycler/src/scoredalignment.cpp:16-136
    enum CigarType { MATCH, INSERTION, DELETION, CLIP, NOTHING };
    int readStartPos = -1, refStartPos = -1, readEndPos = 0, refEndPos = 0, rawScore = 0;
    double scaled = 0.0;
    std::string cigar;
    int alignmentLength = (int)std::max(readAlignment.size(), refAlignment.size());
    if (alignmentLength > 0) {
        CigarType cur = MATCH;
        int curLen = 0, readBases = 0, refBases = 0;
        std::vector<CigarType> types;
        std::vector<int> lens;
        bool started = false, readStarted = false, refStarted = false;
        int startPos = -1, endPos = -1;
        if (startImmediately) { started = readStarted = refStarted = true; readStartPos = refStartPos = 0; startPos = 0; }
        for (int i = 0; i < alignmentLength; ++i) {
            char b1 = readAlignment[(size_t)i], b2 = refAlignment[(size_t)i];
            if (b1 != '-') readStarted = true;
            if (b2 != '-') refStarted = true;
            if (readStarted && refStarted && !started) {
                readStartPos = readBases; refStartPos = refBases; started = true; startPos = i;
            }
            CigarType t;
            if (b1 == '-') t = started ? DELETION : NOTHING;
            else if (b2 == '-') t = started ? INSERTION : CLIP;
            else t = MATCH;
            if (i == 0) cur = t;
            if (t == cur) ++curLen;
            else { types.push_back(cur); lens.push_back(curLen); cur = t; curLen = 1; }
            if (b1 != '-') ++readBases;
            if (b2 != '-') ++refBases;
        }
        endPos = alignmentLength;
        readEndPos = readBases;
        refEndPos = refBases;
        if (cur == INSERTION && !goToEndSeq1) { cur = CLIP; readEndPos -= curLen; endPos -= curLen; }
        else if (cur == DELETION && !goToEndSeq2) { cur = NOTHING; refEndPos -= curLen; endPos -= curLen; }
        types.push_back(cur);
        lens.push_back(curLen);
        int pos = 0;
        for (size_t i = 0; i < types.size(); ++i) {
            CigarType t = types[i];
            int len = lens[i];
            if (t == DELETION) cigar += std::to_string(len) + "D";
            else if (t == INSERTION) cigar += std::to_string(len) + "I";
            else if (t == CLIP) cigar += std::to_string(len) + "S";
            else if (t == MATCH) cigar += std::to_string(len) + "M";
            if (t == INSERTION || t == DELETION) rawScore += sc.gapOpen + (len - 1) * sc.gapExtend;
            else if (t == MATCH)
                for (int k = 0; k < len; ++k)
                    rawScore += (readAlignment[(size_t)(pos + k)] == refAlignment[(size_t)(pos + k)]) ? sc.match
                                                                                                     : sc.mismatch;
            pos += len;
        }
        int lenNoClips = endPos - startPos;
        int perfect = sc.match * lenNoClips, worst = sc.mismatch * lenNoClips;
        if (perfect > worst) scaled = 100.0 * double(rawScore - worst) / double(perfect - worst);
        else scaled = 0.0;
        refStartPos += refOffset;
        refEndPos += refOffset;
    }
    if (scaledOut) *scaledOut = scaled;
    // getFullString :139-156
    std::string rc = (!readName.empty() && readName.back() == '-') ? "-" : "+";
    return refName + "," + rc + "," + std::to_string(readStartPos) + "," + std::to_string(readEndPos) + "," +
           std::to_string(refStartPos) + "," + std::to_string(refEndPos) + "," + std::to_string(rawScore) + "," +
           std::to_string(scaled) + "," + "0" + "," + cigar;
}

}  // namespace orc
