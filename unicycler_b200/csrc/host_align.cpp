// Host side of the alignment path (see host_align.hpp).
#include "host_align.hpp"

#include <algorithm>
#include <cstdio>

#include "dpgeom.hpp"

namespace ub200 {

namespace {
struct Dna5Table {
    uint8_t t[256];
    Dna5Table() {
        for (int i = 0; i < 256; ++i) t[i] = 4;
        t[(int)'A'] = t[(int)'a'] = 0;
        t[(int)'C'] = t[(int)'c'] = 1;
        t[(int)'G'] = t[(int)'g'] = 2;
        t[(int)'T'] = t[(int)'t'] = t[(int)'U'] = t[(int)'u'] = 3;
    }
};
const Dna5Table kDna5;
}  // namespace

void toDna5(const char* s, size_t n, std::vector<uint8_t>& out) {
    out.resize(n);
    for (size_t i = 0; i < n; ++i) out[i] = kDna5.t[(unsigned char)s[i]];
}

// ---------------------------------------------------------------------------------------
// chain planner
// ---------------------------------------------------------------------------------------
namespace {

struct Planner {
    long lenH, lenV, b;
    long hNext = 0, vNext = 0;      // DPScoutState_::_horizontalNextGridOrigin / _verticalNextGridOrigin
    long capH = 0, capV = 0;        // lengths of the (grow-only) next-init arrays
    long zerosH = 0, zerosV = 0;    // pending _initiaizeBeginningOfBandedChain sizes
    bool ok = true;
    std::vector<GridDesc>& out;
    long long* colTabCount;   // running number of column descriptors of the job (the device generates them)

    Planner(long h, long v, long band, std::vector<GridDesc>& o, long long* ct)
        : lenH(h), lenV(v), b(band), out(o), colTabCount(ct) {}

    static long hShiftBegin(const ChainSeed& s) { return s.upperDiag - (s.beginH - s.beginV); }
    static long vShiftBegin(const ChainSeed& s) { return (s.beginH - s.beginV) - s.lowerDiag; }
    static long hShiftEnd(const ChainSeed& s) { return s.endH - s.endV - s.lowerDiag; }
    static long vShiftEnd(const ChainSeed& s) { return s.upperDiag - s.endH + s.endV; }

    // _reinitScoutState, seeds/banded_chain_alignment_scout.h:175-221 (sizes only)
    void reinit(long originH, long originV, long /*sizeCurH*/, long /*sizeCurV*/, long sizeNextH, long sizeNextV) {
        hNext = originH;
        vNext = originV;
        capH = std::max(capH, sizeNextH);
        capV = std::max(capV, sizeNextV);
        if (originH < 0 || originV < 0) ok = false;  // unsigned wrap in the reference: undefined there
    }

    void emit(int kind, long h0, long h1, long v0, long v1, bool banded, long lo, long up, int glue, int check) {
        GridDesc g;
        g.kind = kind;
        g.h0 = (int32_t)h0; g.v0 = (int32_t)v0;
        g.nH = (int32_t)(h1 - h0); g.nV = (int32_t)(v1 - v0);
        g.banded = banded ? 1 : 0; g.lo = (int32_t)lo; g.up = (int32_t)up;
        g.hNext = (int32_t)hNext; g.vNext = (int32_t)vNext;
        g.capNextH = (int32_t)capH; g.capNextV = (int32_t)capV;
        g.plantZerosH = (int32_t)zerosH; g.plantZerosV = (int32_t)zerosV;
        zerosH = zerosV = 0;
        g.glue = glue;
        g.checkScore = check;
        if (g.nH < 1 || g.nV < 1 || h0 < 0 || v0 < 0 || h1 > lenH || v1 > lenV) ok = false;
        if (banded) {
            if (!(lo < 0 && up > 0)) ok = false;
            // a band that spans the whole matrix is stored and traversed exactly like the unbanded
            // matrix (all FullColumns, dimV = nV + 1); run it on the unbanded kernel
            if (lo <= -(long)g.nV && up >= (long)g.nH) { g.banded = 0; g.lo = 0; g.up = 0; }
        }
        if (kind == GRID_CHAIN_FINAL && !g.banded && (hNext != 0 || vNext != 0)) ok = false;
        g.colTabOff = 0; g.nColTab = 0; g.persistOff = -1; g.ckTiles = 0; g.pad = 0;
        if (colTabCount && ok && g.banded && kind != GRID_GLOBAL) {
            // the tracking pass needs the column descriptors of _computeBandedAlignment right of the next grid's
            // origin (seeds/banded_chain_alignment_impl.h:282-377): only their number is planned here, the device
            // walks the columns itself (engine.cu: colTabKernel)
            g.colTabOff = (int32_t)*colTabCount;
            g.nColTab = bandColumnsFrom(makeGeom(g.nH, g.nV, g.banded, g.lo, g.up), g.hNext);
            *colTabCount += g.nColTab;
        }
        out.push_back(g);
    }

    // seeds/banded_chain_alignment_impl.h:737-894
    void initializeChain(const ChainSeed& seed) {
        long hShift = hShiftBegin(seed), vShift = vShiftBegin(seed);
        long hNextO = std::max(0L, seed.beginH + 1 - b);
        long vNextO = std::max(0L, seed.beginV + 1 - b);
        long up = std::min(lenH, hNextO + (b << 1) + hShift + std::max(0L, b - seed.beginV - 1) +
                                     std::min(0L, seed.beginH + 1 - b));
        long lo = -std::min(lenV, vNextO + (b << 1) + vShift + std::max(0L, b - seed.beginH - 1) +
                                      std::min(0L, seed.beginV + 1 - b));
        if (hNextO != 0 || vNextO != 0) {
            zerosH = up + 1; zerosV = 1 - lo;
            reinit(hNextO, vNextO, 1 + up, 1 - lo, 1 + up - hNextO, 1 - lo - vNextO);
            emit(GRID_CHAIN_INITIAL, 0, up, 0, -lo, false, 0, 0, GLUE_APPEND, 0);
        } else {
            zerosH = up; zerosV = -lo;
        }
        long gb1 = hNextO, gb2 = vNextO;
        long ge1 = std::min(lenH, seed.endH + b), ge2 = std::min(lenV, seed.endV + b);
        long infH = ge1 - gb1, infV = ge2 - gb2;
        hShift = hShiftEnd(seed);
        vShift = vShiftEnd(seed);
        hNextO = std::max(0L, seed.endH - b - hShift - std::max(0L, seed.endV + b - lenV) - gb1);
        vNextO = std::max(0L, seed.endV - b - vShift - std::max(0L, seed.endH + b - lenH) - gb2);
        up -= gb1;
        lo += gb2;
        if (infV + lo > up) vNextO -= (infV + lo) - up;
        reinit(hNextO, vNextO, up + 1, 1 - lo, infH - hNextO + 1, infV - vNextO + 1);
        if (gb1 == 0 && gb2 == 0) {
            if (ge1 == lenH && ge2 == lenV) emit(GRID_GLOBAL, gb1, ge1, gb2, ge2, true, lo, up, GLUE_ASSIGN, 1);
            else emit(GRID_CHAIN_INITIAL, gb1, ge1, gb2, ge2, true, lo, up, GLUE_ASSIGN, 1);
        } else {
            if (ge1 == lenH && ge2 == lenV) emit(GRID_CHAIN_FINAL, gb1, ge1, gb2, ge2, true, lo, up, GLUE_IF_NONEMPTY, 1);
            else emit(GRID_CHAIN_INNER, gb1, ge1, gb2, ge2, true, lo, up, GLUE_IF_NONEMPTY, 1);
        }
        hNext += gb1;
        vNext += gb2;
        if (infV + lo > up) vNext += (infV + lo) - up;
    }

    // :901-959
    void gapArea(const ChainSeed& seed) {
        long gb1 = hNext, gb2 = vNext;
        long ge1 = seed.beginH + 1 + b + hShiftBegin(seed);
        long ge2 = seed.beginV + 1 + b + vShiftBegin(seed);
        long hNextO = seed.beginH + 1 - b - gb1;
        long vNextO = seed.beginV + 1 - b - gb2;
        reinit(hNextO, vNextO, ge1 - gb1 + 1, ge2 - gb2 + 1, ge1 - gb1 + 1 - hNextO, ge2 - gb2 + 1 - vNextO);
        emit(GRID_CHAIN_INNER, gb1, ge1, gb2, ge2, false, 0, 0, GLUE_IF_NONEMPTY, 1);
        hNext += gb1;
        vNext += gb2;
    }

    // :966-1027
    void anchorArea(const ChainSeed& seed) {
        long gb1 = hNext, gb2 = vNext;
        long ge1 = seed.endH + b, ge2 = seed.endV + b;
        long infH = ge1 - gb1, infV = ge2 - gb2;
        long hNextO = seed.endH - b - hShiftEnd(seed) - gb1;
        long vNextO = seed.endV - b - vShiftEnd(seed) - gb2;
        long lo = -(b << 1) - vShiftBegin(seed), up = (b << 1) + hShiftBegin(seed);
        long relV = vNextO;
        if (infV + lo > up) relV -= (infV + lo) - up;
        reinit(hNextO, relV, up + 1, 1 - lo, infH - hNextO + 1, infV - vNextO + 1);
        emit(GRID_CHAIN_INNER, gb1, ge1, gb2, ge2, true, lo, up, GLUE_IF_NONEMPTY, 0);
        hNext += gb1;
        vNext += gb2;
        if (infV + lo > up) vNext += (infV + lo) - up;
    }

    // :1038-1177
    void finishChain(const ChainSeed& seed) {
        long gb1 = hNext, gb2 = vNext;
        long hShift = hShiftBegin(seed), vShift = vShiftBegin(seed);
        long ge1 = std::min(lenH, seed.beginH + 1 + b + hShift);
        long ge2 = std::min(lenV, seed.beginV + 1 + b + vShift);
        long infH = ge1 - gb1, infV = ge2 - gb2;
        long hNextO = std::max(0L, seed.beginH + 1 - b - gb1);
        long vNextO = std::max(0L, seed.beginV + 1 - b - gb2);
        reinit(hNextO, vNextO, infH + 1, infV + 1, infH - hNextO + 1, infV - vNextO + 1);
        emit(GRID_CHAIN_INNER, gb1, ge1, gb2, ge2, false, 0, 0, GLUE_IF_NONEMPTY, 0);
        gb1 += hNextO;
        gb2 += vNextO;
        ge1 = std::min(lenH, seed.endH + b);
        ge2 = std::min(lenV, seed.endV + b);
        infH = ge1 - gb1;
        infV = ge2 - gb2;
        if (ge1 == lenH && ge2 == lenV) {
            long lo = -(ge2 - gb2), up = ge1 - gb1;
            reinit(0, 0, up + 1, 1 - lo, up + 1, 1 - lo);
            emit(GRID_CHAIN_FINAL, gb1, ge1, gb2, ge2, true, lo, up, GLUE_IF_NONEMPTY, 0);
            return;
        }
        long lo = -(b << 1) - vShift, up = (b << 1) + hShift;
        hNextO = std::max(0L, seed.endH - b - hShiftEnd(seed) - gb1 - std::max(0L, seed.endV + b - lenV));
        vNextO = std::max(0L, seed.endV - b - vShiftEnd(seed) - gb2 - std::max(0L, seed.endH + b - lenH));
        if (infV + lo > up) vNextO -= (infV + lo) - up;
        reinit(hNextO, vNextO, up + 1, 1 - lo, infH - hNextO + 1, infV - vNextO + 1);
        emit(GRID_CHAIN_INNER, gb1, ge1, gb2, ge2, true, lo, up, GLUE_IF_NONEMPTY, 0);
        gb1 += hNextO;
        if (infV + lo > up) vNextO += (infV + lo) - up;
        gb2 += vNextO;
        reinit(0, 0, lenH - gb1 + 1, lenV - gb2 + 1, lenH - gb1 + 1, lenV - gb2 + 1);
        emit(GRID_CHAIN_FINAL, gb1, lenH, gb2, lenV, false, 0, 0, GLUE_IF_NONEMPTY, 1);
    }

    // :1212-1296
    void run(const std::vector<ChainSeed>& seeds) {
        size_t it = 0, last = seeds.size() - 1;
        {  // _findFirstAnchor :598-621
            size_t i = 0;
            bool found = false;
            while (i != last) {
                const ChainSeed& s = seeds[++i];
                if (s.beginH - b <= 0) continue;
                if (s.beginV - b <= 0) continue;
                it = i - 1;
                found = true;
                break;
            }
            if (!found) it = i;
        }
        size_t itEnd;
        {  // _findLastAnchor :623-647
            size_t i = last;
            while (i != it) {
                const ChainSeed& s = seeds[--i];
                if (s.endH + b >= lenH) continue;
                if (s.endV + b >= lenV) continue;
                break;
            }
            itEnd = i;
        }
        initializeChain(seeds[it]);
        if (seeds.size() == 1 || (it == itEnd && itEnd == last)) {
            if (seeds[it].endH + b < lenH || seeds[it].endV + b < lenV) {
                long gbH = hNext, gbV = vNext;
                reinit(0, 0, lenH + 1 - gbH, lenV + 1 - gbV, lenH + 1 - gbH, lenV + 1 - gbV);
                emit(GRID_CHAIN_FINAL, gbH, lenH, gbV, lenV, false, 0, 0, GLUE_ALWAYS, 0);
            }
            return;
        }
        while (it != itEnd) {
            ++it;
            gapArea(seeds[it]);
            anchorArea(seeds[it]);
        }
        ++it;
        if (it >= seeds.size()) { ok = false; return; }
        finishChain(seeds[it]);
    }
};

}  // namespace

bool planChain(const std::vector<ChainSeed>& chain, long lenH, long lenV, long bandExtension,
               std::vector<GridDesc>& grids, long long* colTabCount) {
    grids.clear();
    if (colTabCount) *colTabCount = 0;
    if (chain.empty() || lenH < 1 || lenV < 1) return false;
    Planner p(lenH, lenV, bandExtension, grids, colTabCount);
    p.run(chain);
    if (!p.ok) grids.clear();
    return p.ok;
}

bool planGlobal(long lenH, long lenV, bool banded, long lo, long up, bool freeFirstRow, bool freeFirstCol,
                bool freeLastRow, bool freeLastCol, std::vector<GridDesc>& grids) {
    grids.clear();
    // _isValidDPSettings / _checkBandProperties
    if (lenH < 1 || lenV < 1) return false;
    if (banded) {
        if (up < -lenV || lo > lenH) return false;
        if (up < 0 && !freeFirstCol) return false;
        if (lo > 0 && !freeFirstRow) return false;
        if (up + lenV < lenH && !freeLastRow) return false;
        if (lo + lenV > lenH && !freeLastCol) return false;
        if (!(lo < 0 && up > 0)) return false;  // not on the hot path (band width >= 3 there); see DESIGN.md
    }
    GridDesc g;
    g.kind = GRID_GLOBAL;
    g.h0 = 0; g.v0 = 0; g.nH = (int32_t)lenH; g.nV = (int32_t)lenV;
    g.banded = banded ? 1 : 0; g.lo = banded ? (int32_t)lo : 0; g.up = banded ? (int32_t)up : 0;
    if (banded && lo <= -lenV && up >= lenH) { g.banded = 0; g.lo = 0; g.up = 0; }
    g.hNext = 0; g.vNext = 0; g.capNextH = 0; g.capNextV = 0; g.plantZerosH = 0; g.plantZerosV = 0;
    g.glue = GLUE_ASSIGN; g.checkScore = 1;
    g.colTabOff = 0; g.nColTab = 0; g.persistOff = -1; g.ckTiles = 0; g.pad = 0;
    grids.push_back(g);
    return true;
}

// ---------------------------------------------------------------------------------------
// trace gluing
// ---------------------------------------------------------------------------------------
namespace {

typedef std::vector<Seg> Trace;

inline int segEndH(const Seg& s) { return s.dir == T_V ? s.hBeg : s.hBeg + s.len; }
inline int segEndV(const Seg& s) { return s.dir == T_H ? s.vBeg : s.vBeg + s.len; }

// The reference keeps a trace with its LAST alignment segment first and glues a local trace in front of the global
// one (seeds/banded_chain_alignment_impl.h, _glueTracebacks): a copy of the whole global trace per grid.  Here the
// global traces are kept in alignment order while gluing, so a local trace is appended in O(its own length); the
// segment arithmetic (connection test, merge of equal directions at the glue point, order of the resulting
// traces, erasure of unconnected ones) is the reference's.
typedef std::vector<Seg> TraceFwd;   // alignment order: first segment of the alignment first

inline void appendLocal(TraceFwd& g, const Trace& local) {
    // the glue point: the global trace's last segment and the local trace's first one (stored last)
    size_t k = local.size();
    if (!g.empty() && k > 0 && g.back().dir == local[k - 1].dir) {   // _smoothGluePoint
        g.back().len += local[k - 1].len;
        --k;
    }
    for (; k > 0; --k) g.push_back(local[k - 1]);
}

void glueTracebacksFwd(std::vector<TraceFwd>& global, const std::vector<Trace>& local) {
    if (global.empty()) {
        for (const Trace& t : local) global.emplace_back(t.rbegin(), t.rend());
        return;
    }
    const size_t lengthGlobal = global.size();
    std::vector<size_t> toErase;
    for (size_t j = 0; j < lengthGlobal; ++j) {
        const Seg gEnd = global[j].back();
        const size_t numCurr = global[j].size();
        bool connected = false;
        for (size_t i = 0; i < local.size(); ++i) {
            const Seg& lBeg = local[i].back();
            if (segEndH(gEnd) != lBeg.hBeg || segEndV(gEnd) != lBeg.vBeg) continue;
            if (connected) {
                // a further local trace continues the same global trace: a new trace from the old part
                TraceFwd joined(global[j].begin(), global[j].begin() + (long)numCurr);
                joined.back() = gEnd;
                appendLocal(joined, local[i]);
                global.push_back(std::move(joined));
            } else {
                appendLocal(global[j], local[i]);
                connected = true;
            }
        }
        if (!connected) toErase.push_back(j);
    }
    for (size_t i = toErase.size(); i > 0; --i) global.erase(global.begin() + (long)toErase[i - 1]);
}

}  // namespace

void glueChain(const std::vector<GridDesc>& grids, const JobResult& res, std::vector<Seg>& trace, bool& empty) {
    std::vector<TraceFwd> global;
    const bool dump = getenv("UNICYCLER_B200_DUMPTRACES") != nullptr;   // developer aid
    for (size_t k = 0; k < grids.size(); ++k) {
        const std::vector<Trace>& local = res.gridTraces[k];
        if (dump) {
            fprintf(stderr, "[ub200 glue] grid %zu kind %d glue %d origin (%d,%d) %dx%d banded %d: %zu local traces, %zu global before;", k, grids[k].kind, grids[k].glue, grids[k].h0, grids[k].v0, grids[k].nH, grids[k].nV, grids[k].banded, local.size(), global.size());
            for (const Trace& t : local)
                if (!t.empty()) fprintf(stderr, " [%zu segs: last-stored (%d,%d,len %d,dir %d) first-stored (%d,%d,len %d,dir %d)]", t.size(), t.back().hBeg, t.back().vBeg, t.back().len, t.back().dir, t.front().hBeg, t.front().vBeg, t.front().len, t.front().dir);
            fprintf(stderr, "\n");
        }
        switch (grids[k].glue) {
        case GLUE_APPEND:
            for (const Trace& t : local) global.emplace_back(t.rbegin(), t.rend());
            break;
        case GLUE_ASSIGN:
            global.clear();
            for (const Trace& t : local) global.emplace_back(t.rbegin(), t.rend());
            break;
        case GLUE_IF_NONEMPTY:
            if (!local.empty()) glueTracebacksFwd(global, local);
            break;
        case GLUE_ALWAYS:
            if (local.empty()) {
                // _glueTracebacks with an empty local set: every global trace is unconnected and erased
                if (!global.empty()) global.clear();
            } else
                glueTracebacksFwd(global, local);
            break;
        }
    }
    empty = global.empty();
    if (!empty) trace.assign(global[0].rbegin(), global[0].rend());
    else trace.clear();
}

// ---------------------------------------------------------------------------------------
// ScoredAlignment
// ---------------------------------------------------------------------------------------
void scoreAlignment(const std::vector<Seg>& trace, bool traceEmpty, const uint8_t* H, long lenH, const uint8_t* V,
                    long lenV, int refOffset, bool startImmediately, bool goToEndSeq1, bool goToEndSeq2,
                    const Scoring& sc, AlignmentRecord& rec) {
    (void)lenH; (void)lenV;
    rec = AlignmentRecord();
    if (traceEmpty || trace.empty()) return;  // reference: rows left unaligned -> undefined output
    enum CigarType { MATCH, INSERTION, DELETION, CLIP, NOTHING };
    // Walk the alignment columns from the first trace segment (stored last) to the end.
    long h = trace.back().hBeg, v = trace.back().vBeg;
    long total = 0;
    for (const Seg& s : trace) total += s.len;
    if (total == 0) return;
    rec.emptyAlignment = false;
    std::vector<int> types;   // run-length encoded cigar types
    std::vector<long> lens;
    std::vector<int> runScore;
    int cur = MATCH;
    long curLen = 0;
    int curScore = 0;
    bool started = startImmediately, readStarted = startImmediately, refStarted = startImmediately;
    long readBases = 0, refBases = 0, startPos = startImmediately ? 0 : -1, col = 0;
    if (startImmediately) { rec.readStart = 0; rec.refStart = 0; }
    bool first = true;
    // Every column of one trace segment has the same kind (the "started" state can only change at a segment's first
    // column), so the walk is per segment; a diagonal segment only needs its number of equal bases.
    for (size_t k = trace.size(); k > 0; --k) {
        const Seg& s = trace[k - 1];
        const long L = s.len;
        if (L <= 0) continue;
        const bool hasRead = (s.dir != T_V), hasRef = (s.dir != T_H);
        if (hasRead) readStarted = true;
        if (hasRef) refStarted = true;
        if (readStarted && refStarted && !started) {
            rec.readStart = (int)readBases; rec.refStart = (int)refBases; started = true; startPos = col;
        }
        int type;
        int segScore = 0;
        if (!hasRead) type = started ? DELETION : NOTHING;
        else if (!hasRef) type = started ? INSERTION : CLIP;
        else {
            type = MATCH;
            long equal = 0;
            const uint8_t* a = H + h;
            const uint8_t* b = V + v;
            for (long t = 0; t < L; ++t) equal += (a[t] == b[t]);
            segScore = (int)(equal * sc.match + (L - equal) * sc.mismatch);
        }
        if (first) { cur = type; first = false; }
        if (type == cur) { curLen += L; curScore += segScore; }
        else {
            types.push_back(cur); lens.push_back(curLen); runScore.push_back(curScore);
            cur = type; curLen = L; curScore = segScore;
        }
        if (hasRead) { readBases += L; h += L; }
        if (hasRef) { refBases += L; v += L; }
        col += L;
    }
    long endPos = total;
    rec.readEnd = (int)readBases;
    rec.refEnd = (int)refBases;
    if (cur == INSERTION && !goToEndSeq1) { cur = CLIP; rec.readEnd -= (int)curLen; endPos -= curLen; }
    else if (cur == DELETION && !goToEndSeq2) { cur = NOTHING; rec.refEnd -= (int)curLen; endPos -= curLen; }
    types.push_back(cur); lens.push_back(curLen); runScore.push_back(curScore);
    // (a 12 kb read at 15 % errors has ~3 600 runs: digits are written in place, no temporary strings)
    rec.cigar.reserve(types.size() * 3 + 16);
    auto run = [&](long len, char op) {
        char buf[24];
        int nd = 0;
        do { buf[nd++] = (char)('0' + len % 10); len /= 10; } while (len > 0);
        while (nd > 0) rec.cigar.push_back(buf[--nd]);
        rec.cigar.push_back(op);
    };
    for (size_t i = 0; i < types.size(); ++i) {
        const long len = lens[i];
        switch (types[i]) {
        case DELETION: run(len, 'D'); rec.rawScore += sc.gapOpen + (int)(len - 1) * sc.gapExtend; break;
        case INSERTION: run(len, 'I'); rec.rawScore += sc.gapOpen + (int)(len - 1) * sc.gapExtend; break;
        case CLIP: run(len, 'S'); break;
        case MATCH: run(len, 'M'); rec.rawScore += runScore[i]; break;
        default: break;
        }
    }
    const int lenNoClips = (int)(endPos - startPos);
    const int perfect = sc.match * lenNoClips, worst = sc.mismatch * lenNoClips;
    if (perfect > worst) rec.scaledScore = 100.0 * double(rec.rawScore - worst) / double(perfect - worst);
    else rec.scaledScore = 0.0;
    rec.refStart += refOffset;
    rec.refEnd += refOffset;
}

std::string fullString(const AlignmentRecord& rec, const std::string& readName, const std::string& refName,
                       long long milliseconds) {
    const char* strand = (!readName.empty() && readName.back() == '-') ? "-" : "+";
    std::string out;
    out.reserve(refName.size() + rec.cigar.size() + 96);
    out += refName; out += ','; out += strand; out += ',';
    out += std::to_string(rec.readStart); out += ','; out += std::to_string(rec.readEnd); out += ',';
    out += std::to_string(rec.refStart); out += ','; out += std::to_string(rec.refEnd); out += ',';
    out += std::to_string(rec.rawScore); out += ','; out += std::to_string(rec.scaledScore); out += ',';
    out += std::to_string((int)milliseconds); out += ','; out += rec.cigar;
    return out;
}

}  // namespace ub200
