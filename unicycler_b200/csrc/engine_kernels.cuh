// Device side of the B200 DP engine (sm_100a).  See engine.hpp for the job model.
//
// Execution model (one persistent kernel, one CTA of NWARPS warps per SM, every WARP is an agent):
//   * control agents (the first NCTRL warps of a CTA) pull jobs from an atomic queue and walk the job's
//     grids in order, because grid k+1 is initialised from the traceback of grid k
//     (seeds/banded_chain_alignment_traceback.h:296-330).
//   * a SMALL grid (<= 256 rows, trace fits the agent's shared-memory window) is filled by the control
//     warp itself with the full trace byte per cell (the reference's TraceBitMap_) kept in shared memory.
//   * a BIG grid is published as a task on a global FIFO board.  Its matrix is cut into strips of 256
//     rows x segments of SEG columns; every warp of the GPU claims (strip, segment) items in wavefront
//     order and runs a SCORE-ONLY fill (no trace bits: ~10 instructions per cell).  Inside an item lane t
//     owns 8 consecutive rows and walks the columns one step behind lane t-1 (anti-diagonal wavefront,
//     cell hand-off with __shfl_up).  The row between two strips and, every CKW columns, the column state
//     of a strip (plus a row every 64 rows) are written to HBM as CHECKPOINTS (0.25 B per cell instead of 1 B/cell);
//     release/acquire progress counters order producer and consumer items.
//   * traceback of a big grid: the control warp RECOMPUTES the trace bytes of the 64 x 64 tile under
//     the path from the checkpoints (bit-identical, integer arithmetic) into its shared-memory window and
//     walks it exactly like SeqAn's TracebackCoordinator_, in SeqAn's storage coordinates (dpgeom.hpp).
#pragma once
#include <cuda_runtime.h>

#include "dpgeom.hpp"
#include "engine.hpp"

namespace ub200 {

constexpr int NWARPS = 16;
constexpr int NTHREADS = NWARPS * 32;
constexpr int NCTRL = 4;                 // control-capable warps per CTA
constexpr int MODE_TASK = 0;             // score-only strip item of a published grid (checkpoints to HBM)
constexpr int MODE_TRACE = 1;            // full trace bytes into the shared-memory window
constexpr int MODE_TRACEG = 3;           // the same into global memory (a tile recomputed for another warp)
constexpr int MODE_FAST = 2;             // score-only local fill, box cells (S,H,V) into the shared-memory window
constexpr int MAXSEG = 24;               // speculative segments of one seed chain
constexpr int MAXREC = 8;                // candidates / planted cells a pass-1 grid record can hold
constexpr int NBOARD = 4;                // task boards, served in this order: [0] holds the grids with the longest remaining spine
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
    long long outOff, colTabBase;
    int outCap;
    int nSeg;                  // the chain is walked by nSeg control warps at once (speculative segments)
    int segStart[MAXSEG + 1];  // first grid of every segment; segStart[nSeg] = gridCount
    int pad2;
    long long recBase;         // GridRec index of (segment 0, grid 0); record of (p, k) = recBase + p * gridCount + k
    long long recIdxBase;      // this job's slice of the record index (KParams::recIdx)
    int recIdxCap, pad3;
};

struct JobOut {
    int status, score, outLen, pad;
    long long tSpine, tFinal;   // ns after kernel start: pass 1 resolved / job complete
    long long tP2Start, p2MaxNs, p2MaxItem, p2SumNs, p2MaxTiles, p2MaxTileCycles, p2MaxStart, tFin0, finRecords;  // developer timeline: start of the last pass-2 item, longest item
    long long prof[12];  // cycles: setup+init, local fill, task wait, track, traceback, total; tiles, tile cycles,
                         // local-grid traceback cycles, tracebacks, local grids, local-grid track cycles
};

struct ScratchLayout {  // byte offsets inside one control agent's arena
    long long rowCk, colCk, ckBase, rowProg, segDone, initRow, initCol, hInitNext, vInitNext, box, lastRow, lastCol,
        cand, planted, total;
    int maxCand, maxPlanted, maxStrips, pad;
    long long maxBox, maxRowCk, maxColCk;
    int maxCapH, maxCapV, maxNH, maxNV;
};

struct PlantedCell {
    int i1, i2;
    DCell c;
};

// Everything a grid needs.  Built by lane 0 of the control warp in shared memory; copied to the task
// board (global memory) when the grid is published, and from there into the worker warp's shared slot.
struct GridCtx {
    GridGeom g;
    int kind, h0, v0, hNext, vNext, capNextH, capNextV;
    int match, mismatch, go, ge, fe, complete, affine;
    const uint8_t* seqH;
    const uint8_t* seqV;
    // arena pointers
    int2* rowCk;        // [NS * SH / CKR][nH+1] (S,V) of every CKR-th row (row checkpoints; the last one of a strip is
                        // also the boundary the strip below starts from)
    int2* colCk;        // [ckBase[s] + c][SH] (S,H) of column (ckFirst(s)+c)*CKW
    int* ckBase;        // per strip: first checkpoint tile index
    int* rowProg;       // per strip: last boundary column written (release/acquire)
    int* segDone;       // per strip: number of the next segment that may start
    int* readyUpTo;     // task board entry: highest strip index that may be claimed (its upstream strip is far enough)
    long long taskId;   // index of the task board entry (board * maxTasks + slot)
    DCell *initRow, *initCol, *hInitNext, *vInitNext, *box, *lastRow, *lastCol;
    int* cand;
    PlantedCell* planted;
    const ColInfo* colTab;   // host-planned column descriptors of this grid (banded chain grids)
    int nColTab, pad3;
    // strips
    int NS, nSeg, local, RR;   // local: filled by the control warp with RR rows per lane, trace in shared memory
    int pitch, localJhi;       // local trace window: lanes per column (even), last column holding cells
    int colZeroMax, rrMul;     // rrMul: (x * rrMul) >> 16 == x / RR for x < 1024
    int rrs, pad2;             // bytes per (column, lane) slot of the local trace window
    // capture
    int capEdges;  // 1: last row + last column; 0: box
    int boxRow0, boxH, boxW;
    // limits
    int maxCand, maxPlanted, pad0, pad4;
    long long maxBox;
    // pass-1 fast mode: the box cells (S,H,V) of rows >= fastR0, columns >= fastC0 live in the shared-memory window
    int fastOk, fastR0, fastC0, fastPitch;
    const uint8_t* fastSeqH;   // shared memory: base codes of columns fastC0.. (index j - fastC0)
    const uint8_t* fastSeqV;   // shared memory: base codes of rows fastR0..    (index i - fastR0)
};
constexpr int FASTSEQ_H = 768, FASTSEQ_V = 256;   // capacity of the staged code windows

// Pass-1 record of a chain grid: everything pass 2 needs to redo the grid on its own.
struct GridRec {
    int state;        // 0: done in line by pass 1; 1: small grid (pass 2 fills the trace and walks all candidates);
                      // 2: big grid with a persistent block (pass 2 walks one candidate per item)
    int nCand, inserted, nPlantedIn;
    int relMax;       // grid maximum in the writing segment's score frame
    int published;    // 1: the big grid's candidates have been handed to pass 2 (by whoever flipped it from 0)
    int cand[MAXREC];
    PlantedCell plantedIn[MAXREC];
};

struct JobState {     // zeroed before every launch
    int outCursor;    // ints reserved in the job's segment stream
    int status;       // max over the statuses of pass-2 grids
    int segStopped;   // number of segments whose control warp has stopped
    int nOwner;       // resolved chain: grids [ownerFrom[t], ownerFrom[t+1]) belong to segment ownerSeg[t]
    int segProgress[MAXSEG];   // grids finished by the segment (next grid to process), release-published
    int segStop[MAXSEG];       // 1: the segment's warp has stopped
    int segClaim[MAXSEG];      // 0: not started, 1: started by its warp, 2: cancelled by an earlier segment that ran past it
    int segSyncSeg[MAXSEG];    // segment it merged into (-1: reached the end of the chain, -2: failed)
    int segSyncGrid[MAXSEG];   // grid at which it merged / failed
    int segDelta[MAXSEG];      // score of its frame minus score of the frame it merged into, at the merge cell
    int segFailStatus[MAXSEG];
    int ownerSeg[MAXSEG + 1], ownerFrom[MAXSEG + 1];
    // A segment is CONFIRMED once a confirmed segment (segment 0 is, from grid 0) has merged into it: from that grid on
    // it is on the resolved chain, and the long tracebacks of its big grids may start before the whole chain is resolved.
    int segConfClaim[MAXSEG];  // 1: some warp is recording the confirmation
    int segConfFrom1[MAXSEG];  // grid from which the segment is confirmed, plus one (0: not confirmed)
    int nRecIdx;      // records committed to the job's record index
    int p2Done;       // pass-2 items finished (the big ones may start before the chain is resolved)
    int p2Need;       // items that make the job complete (0 until the chain is resolved)
};

struct P2Entry {      // pass-2 board entry: one job whose pass 1 is complete
    int jobIdx, nItems, ready, nextItem, doneItems, pad0, pad1, pad2;
};

struct TaskDesc {
    GridCtx ctx;
    int nItems;
    int ready;       // release-published by the control warp
    int nextItem;    // claimed by compare-and-swap, in order
    int doneItems;   // completed items (release)
    int readyUpTo;   // items <= readyUpTo may be claimed: strip s+1 becomes claimable once strip s is one chunk (32 columns) in
    int pad0, pad1, pad2;
};

// A traceback through a big grid asks idle control warps for the trace tiles it is about to enter (the tiles on
// the three tile diagonals ahead of it).  One request = one 64 x 64 tile recomputed from the checkpoints into a
// slot of the asking warp (global memory): [state | .. | int4 tile extent at +16 | 4096 trace bytes at +64].
// state = tag * 4 + {0 posted, 1 claimed by a helper, 2 done, 3 cancelled}; the tag is the slot's generation.
constexpr int TILE_QUEUES = 32;          // independent request rings (pops are compare-and-swap: spread the contention)
constexpr int TILE_RING_CAP = 2048;      // entries per ring
constexpr int TILE_SLOT_BYTES = 4096 + 64;
constexpr int TILE_SLOTS = 32;          // one per lane of the asking warp
constexpr int TILE_CANDS = 21;          // sample points ahead of the walk whose tiles are requested
constexpr int TILE_HELP_MIN_EXTENT = 9000; // nH + nV of a grid whose tracebacks ask for help
constexpr int TILE_HELPERS_PER_WALK = 4; // a walk consumes a tile in a third of the time a helper needs for three
struct ControlBlock {  // zeroed before every launch; every group of counters has its own 128-byte line
    int jobQueue, jobsDone;
    unsigned long long t0;         // %globaltimer at kernel start (developer timeline)
    int pad0[28];
    int ringHead[NBOARD], ringTail[NBOARD];  // task boards, by remaining critical path of the publishing job (GridDesc::pad)
    int pad1[32 - 2 * NBOARD];
    int tokHead[NBOARD], tokTail[NBOARD];    // token rings: one token per strip that became claimable
    int pad2[32 - 2 * NBOARD];
    int p2Head, p2Tail;            // pass-2 board
    int bigHead, bigTail;          // pass-2 items of the big grids (served before everything else of pass 2)
    int pad3[28];
    struct TileQueue { int head, tail; int pad[30]; } tq[TILE_QUEUES];   // tile-request rings (one line each)
    // (the first four are read with one 16-byte load by every idle poll)
    int idleHelpers, activeWalkers; // control warps polling for work / walking a big grid (tiles are only asked for
                                    // while the idle ones outnumber the walkers several times)
    int openTasks;                  // big grids being filled: worker warps serve tile requests only while it is 0 (the
                                    // fills are the critical path of pass 1)
    int tilePending;                // tile requests posted and not yet popped (idle warps scan the rings only when > 0)
    int idleWorkers;                // worker warps polling for work
    int pad5[27];
};

struct TileReq {
    int seq;        // 2 * turn: free, 2 * turn + 1: written (turn = ring position / TILE_RING_CAP)
    int job, gi;
    int tile;       // (64-row block << 16) | 64-column block
    int expect;     // the slot's state word when posted (tag * 4)
    int pad;
    unsigned long long slot;
};

struct KParams {
    const JobDev* jobs;
    const GridDesc* grids;
    const uint8_t* seq;
    int* out;          // segment streams as written: records reserved at their worst-case size, by many warps
    int* out2;         // the same streams compacted by finalizeJob (what the host copies)
    int4* recIdx;      // (position, ints used, grid, segment tag) of every completed record, per job
    JobOut* jobOut;
    const int* order;  // pass-1 work list: jobIdx * MAXSEG + segment, longest chains first
    int nEntries, pad6;
    const ColInfo* colTabPool;
    int nJobs;
    int nSlots;        // control agents with an arena
    int maxTasks;      // capacity of each task board
    int nHiJobs;       // (unused)
    ControlBlock* cb;
    TaskDesc* ring;    // [NBOARD][maxTasks]
    int* tokRing;      // [NBOARD][maxTokens] task id + 1 of a task with a claimable strip
    int maxTokens, pad7;
    P2Entry* p2ring;   // [nJobs]
    int2* bigRing;     // [maxBig] (job + 1, item) of one candidate of a big grid
    int maxBig, pad9;
    TileReq* tileRing; // [TILE_RING_CAP]
    uint8_t* tileSlots; // [control warps][TILE_SLOTS][TILE_SLOT_BYTES]
    JobState* jobState;
    GridRec* gridRecs; // [total grids]
    uint8_t* persist;  // persistent blocks of the big grids (GridDesc::persistOff), nullptr: disabled
    uint8_t* mini;     // per control warp: initRow/initCol for pass-2 grids
    long long miniStride, miniInitCol;
    int fastEnabled, pad5;
    uint8_t* scratch;
    long long scratchStride;
    ScratchLayout lay;
};

// Launch parameters live in constant memory (a by-reference kernel argument would be copied to the
// local-memory stack of every warp).
__constant__ KParams cP;
__device__ unsigned long long gDbg[24];
__device__ unsigned long long gStripLog[16384][4];   // developer timeline of the strips (UNICYCLER_B200_DBG & 16)
__device__ int gStripLogN;
__device__ unsigned long long gGridLog[4096][4];   // developer timeline of one job's spine (UNICYCLER_B200_TRACEJOB)
__device__ int gUbSite;                   // developer aid: source line of the first JOB_REF_UB verdict
// Index of this warp among the CTA's control-capable warps (-1: a worker warp): the highest warp ids of the CTA.
// (Giving them the lowest ids instead — the issue arbiter prefers high ids — made no measurable difference.)
__device__ __forceinline__ int ctrlIndex() {
    const int warp = (int)(threadIdx.x >> 5);
    return warp - (NWARPS - NCTRL) >= 0 ? warp - (NWARPS - NCTRL) : -1;
}
__device__ __forceinline__ int markUb(int line) { atomicCAS(&gUbSite, 0, line); return 0; }
#define UB_VERDICT (markUb(__LINE__), JOB_REF_UB)   // developer counters (cycles), lane 0 of control warps

__device__ __forceinline__ uint32_t ldsU8(uint32_t sharedAddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sharedAddr) : "memory");
    return v;
}

// All per-warp working state (GridCtx slots, windows, staged codes) lives in the kernel's dynamic shared
// memory.  Out-of-line functions receive generic pointers to it; toShared() rebases such a pointer on the
// shared array so that the compiler emits LDS/STS instead of generic loads through the L1TEX path.
extern __shared__ __align__(16) uint8_t gSmem[];
template <typename T>
__device__ __forceinline__ T* toShared(T* p) {
    return reinterpret_cast<T*>(gSmem + ((uint32_t)__cvta_generic_to_shared(p) - (uint32_t)__cvta_generic_to_shared(gSmem)));
}
template <typename T>
__device__ __forceinline__ const T* toShared(const T* p) {
    return reinterpret_cast<const T*>(gSmem + ((uint32_t)__cvta_generic_to_shared(p) - (uint32_t)__cvta_generic_to_shared(gSmem)));
}

// ---------------------------------------------------------------------------------------
// memory-order helpers
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int ldAcquire(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void stRelease(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Polling load: strong (L2-coherent) but WITHOUT the L1 invalidation an acquire load carries (CCTL.IVALL
// would wipe the L1 lines of every other warp of the SM on each poll).  Consumers branch on the polled
// value and then read the produced data with ld.cg (L2), producers publish with st.release / fence + atomic.
__device__ __forceinline__ int ldRelaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ DCell ldcgCell(const DCell* p) {
    const int* q = reinterpret_cast<const int*>(p);
    DCell c;
    c.s = __ldcg(q); c.h = __ldcg(q + 1); c.v = __ldcg(q + 2);
    return c;
}
__device__ __forceinline__ unsigned long long globalTimerNs() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int ldVolatile(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// A strip of task `taskId` became claimable: hand a token to the idle warps.
__device__ __forceinline__ void pushToken(long long taskId) {
    const int board = (int)(taskId / cP.maxTasks);
    const int pos = atomicAdd(&cP.cb->tokTail[board], 1);
    if (pos < cP.maxTokens) stRelease(&cP.tokRing[(size_t)board * cP.maxTokens + pos], (int)taskId + 1);
}

// ---------------------------------------------------------------------------------------
// cell recurrences (seqan/align/dp_formula_affine.h:459-636, dp_formula_linear.h:150-291)
// mode: 0 = RecursionDirectionAll, 1 = UpperDiagonal, 2 = LowerDiagonal, 3 = outside the band
// ---------------------------------------------------------------------------------------
template <bool AFF, bool CT, bool BANDED>
__device__ __forceinline__ uint32_t cellUpdate(int& s, int& h, int& v, int sl, int hl, int su, int vu, int sd,
                                               int sub, int go, int ge, int mode) {
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

// Score-only variant of the same recurrences (identical S/H/V values, no trace bits).
template <bool AFF, bool BANDED>
__device__ __forceinline__ void cellScore(int& s, int& h, int& v, int sl, int hl, int su, int vu, int sd, int sub,
                                          int go, int ge, int mode) {
    if (AFF) {
        int hh = __viaddmax_s32(sl, go, hl + ge);
        int vv = __viaddmax_s32(su, go, vu + ge);
        if (BANDED) {
            vv = (mode == 1) ? NEG_INF : vv;
            hh = (mode == 2) ? NEG_INF : hh;
        }
        s = __vimax3_s32(vv, hh, sd + sub);
        h = hh; v = vv;
    } else {
        int tV = su + ge, tH = sl + ge;
        if (BANDED) {
            tV = (mode == 1) ? INT32_MIN : tV;
            tH = (mode == 2) ? INT32_MIN : tH;
        }
        s = __vimax3_s32(sd + sub, tV, tH);
        h = NEG_INF; v = NEG_INF;
    }
    if (BANDED) {
        const bool outside = (mode == 3);
        s = outside ? NEG_INF : s; h = outside ? NEG_INF : h; v = outside ? NEG_INF : v;
    }
}

// S and V-matrix value of the cell just above strip s in column j (j >= 0)
template <bool BANDED, bool L2ONLY>
__device__ __forceinline__ void upBoundary(const GridCtx& G, int s, int SHR, int j, int& bS, int& bV) {
    const int rowAbove = s * SHR;
    if (BANDED) {
        const int d = j - rowAbove;
        if (d < G.g.lo || d > G.g.up) { bS = NEG_INF; bV = NEG_INF; return; }
    }
    if (s == 0) {
        const DCell c = L2ONLY ? ldcgCell(&G.initRow[j]) : G.initRow[j];
        bS = c.s; bV = c.v;
    } else {
        const int2 b = __ldcg(&G.rowCk[(size_t)(s * (SHR / CKR) - 1) * (size_t)(G.g.nH + 1) + j]);
        bS = b.x; bV = b.y;
    }
}

// Register state of one lane inside a strip.
template <int RR>
struct StripState {
    int Sl[RR], Hl[RR];
    uint32_t vm[(RR + 3) / 4];  // one-hot base masks of the lane's rows, one byte per row
    int prevUpS, pubS, pubV, curHc;
};

struct StepConsts {
    int match, mismatch, go, ge, nV, nH, lo, up;
};

// 32 wavefront steps (one chunk) of a strip.  TRACE: trace bytes to the shared-memory window;
// otherwise (score-only) boundary row + column checkpoints to HBM.  CAP: slow variant that also captures
// the cells the scouts may need (last row/column for final/global matrices, the corner box otherwise).
template <bool AFF, bool CT, bool BANDED, int RR, int MODE, bool CAP>
__device__ __forceinline__ void stripSteps(const GridCtx& G, const StepConsts& K, StripState<RR>& st, int c, int lane,
                                           int cBeg, int cEnd, int i0, int bS, int bV, int hcN, int nsteps,
                                           uint8_t* win, int winPitch, int2* rowOut, int2* ckOut) {
    constexpr bool TRACE = (MODE == MODE_TRACE || MODE == MODE_TRACEG);
    const int match = K.match, mismatch = K.mismatch, go = K.go, ge = K.ge;
    const int lo = K.lo, up = K.up;
    int capEdges = 0, hNext = 0, boxRow0 = 0, boxH = 0;
    DCell *box = nullptr, *lastRow = nullptr, *lastCol = nullptr;
    if (CAP) {
        capEdges = G.capEdges; hNext = G.hNext; boxRow0 = G.boxRow0; boxH = G.boxH;
        box = G.box; lastRow = G.lastRow; lastCol = G.lastCol;
    }
    const int kEnd = imin(32, nsteps - 32 * c);
#pragma unroll 1
    for (int kk = 0; kk < kEnd; ++kk) {
        const int k = 32 * c + kk;
        int inS = __shfl_up_sync(FULLMASK, st.pubS, 1);
        int inV = __shfl_up_sync(FULLMASK, st.pubV, 1);
        int inHc = __shfl_up_sync(FULLMASK, st.curHc, 1);
        const int l0S = __shfl_sync(FULLMASK, bS, kk);
        const int l0V = __shfl_sync(FULLMASK, bV, kk);
        const int l0Hc = __shfl_sync(FULLMASK, hcN, kk);
        if (lane == 0) { inS = l0S; inV = l0V; inHc = l0Hc; }
        st.curHc = inHc;
        const int j = cBeg + k - lane;
        const bool act = (k >= lane) && (j <= cEnd) && (i0 <= K.nV);  // lanes below the matrix stay idle
        if (act) {
            int Sd = st.prevUpS, Su = inS, Vu = inV;
            uint32_t tw[(RR + 3) / 4];
#pragma unroll
            for (int w4 = 0; w4 < (RR + 3) / 4; ++w4) tw[w4] = 0u;
            const uint32_t hcRep = (1u << st.curHc) * 0x01010101u;
            uint32_t eq[(RR + 3) / 4];
#pragma unroll
            for (int w4 = 0; w4 < (RR + 3) / 4; ++w4) eq[w4] = st.vm[w4] & hcRep;
            int vArr[CAP ? RR : 1];
#pragma unroll
            for (int r = 0; r < RR; ++r) {
                int mode = 0;
                if (BANDED) {
                    const int d = j - (i0 + r);
                    mode = (d < lo || d > up) ? 3 : (d == up ? 1 : (d == lo ? 2 : 0));
                }
                const int sub = (eq[r >> 2] & (0xffu << (8 * (r & 3)))) ? match : mismatch;
                int ns, nh, nv;
                if (TRACE) {
                    const uint32_t tv = cellUpdate<AFF, CT, BANDED>(ns, nh, nv, st.Sl[r], st.Hl[r], Su, Vu, Sd, sub, go, ge, mode);
                    tw[r >> 2] |= tv << (8 * (r & 3));
                } else {
                    cellScore<AFF, BANDED>(ns, nh, nv, st.Sl[r], st.Hl[r], Su, Vu, Sd, sub, go, ge, mode);
                }
                Sd = st.Sl[r];
                st.Sl[r] = ns; st.Hl[r] = nh; Su = ns; Vu = nv;
                if (CAP) vArr[r] = nv;
            }
            st.prevUpS = inS;
            st.pubS = Su; st.pubV = Vu;
            if constexpr (TRACE) {
                constexpr int RRS = (RR == 3) ? 4 : RR;
                uint8_t* p = win + ((j - cBeg) * winPitch + lane) * RRS;
                if constexpr (RR == 8) *reinterpret_cast<uint2*>(p) = make_uint2(tw[0], tw[(RR + 3) / 4 - 1]);
                else if constexpr (RR == 2) *reinterpret_cast<uint16_t*>(p) = (uint16_t)tw[0];
                else *reinterpret_cast<uint32_t*>(p) = tw[0];
            } else if constexpr (MODE == MODE_TASK) {
                constexpr int LPB = (RR == 2 || RR == 4 || RR == 8) ? CKR / RR : 8;   // lanes per 64-row checkpoint block
                if ((lane & (LPB - 1)) == LPB - 1) __stcg(&rowOut[(size_t)(lane / LPB) * (size_t)(K.nH + 1) + j], make_int2(Su, Vu));
                if ((j & (CKW - 1)) == 0) {
                    int2* ck = ckOut + (size_t)(j / CKW) * SH + lane * RR;
#pragma unroll
                    for (int r = 0; r < RR; r += 2)
                        __stcg(reinterpret_cast<int4*>(ck + r), make_int4(st.Sl[r], st.Hl[r], st.Sl[r + 1], st.Hl[r + 1]));
                }
            }
            if constexpr (CAP && MODE == MODE_FAST) {
                // pass-1 fast mode: (S,H,V) of the box cells into the shared-memory window
                const int fc0 = G.fastC0, fr0 = G.fastR0, fp = G.fastPitch;
                if (j >= fc0) {
                    DCell* col = reinterpret_cast<DCell*>(win) + (j - fc0) * fp - fr0;
#pragma unroll
                    for (int r = 0; r < RR; ++r) {
                        const int i = i0 + r;
                        if (i >= fr0 && i <= K.nV) col[i] = DCell{st.Sl[r], st.Hl[r], vArr[r]};
                    }
                }
            } else if (CAP) {
                if (capEdges) {
                    const int rl = K.nV - i0;
                    if (rl >= 0 && rl < RR) {
#pragma unroll
                        for (int r = 0; r < RR; ++r)
                            if (r == rl) lastRow[j] = DCell{st.Sl[r], st.Hl[r], vArr[r]};
                    }
                    if (j == K.nH) {
#pragma unroll
                        for (int r = 0; r < RR; ++r)
                            if (i0 + r <= K.nV) lastCol[i0 + r] = DCell{st.Sl[r], st.Hl[r], vArr[r]};
                    }
                } else if (j >= hNext) {
                    DCell* col = box + (size_t)(j - hNext) * boxH - boxRow0;
#pragma unroll
                    for (int r = 0; r < RR; ++r) {
                        const int i = i0 + r;
                        bool inb = true;
                        if (BANDED) { const int d = j - i; inb = (d >= lo && d <= up); }
                        // unbanded: only the perimeter of the box is ever read by the tracking pass
                        else inb = (i == boxRow0) || (i == K.nV) || (j == hNext) || (j == K.nH);
                        if (i >= boxRow0 && i <= K.nV && inb) col[i] = DCell{st.Sl[r], st.Hl[r], vArr[r]};
                    }
                }
            }
        }
    }
}

// Dynamic shared memory of the kernel: [control windows | GridCtx slots | staged codes of the fast path | lean buffers]
constexpr int CTX_STRIDE = (int)((sizeof(GridCtx) + 15) / 16 * 16);
constexpr int SMEM_CTX = NCTRL * WINBYTES;
constexpr int SMEM_FASTSEQ = SMEM_CTX + (NCTRL + NWARPS) * CTX_STRIDE;
constexpr int SMEM_LEAN = SMEM_FASTSEQ + NCTRL * (FASTSEQ_H + FASTSEQ_V);
constexpr int LEAN_BYTES = 512;   // per warp: boundary feed of one chunk (32 x int2) + one-hot column-base window (64 x u32)
constexpr int SMEM_BYTES = SMEM_LEAN + NWARPS * LEAN_BYTES;

// One chunk (32 wavefront steps) of the score-only strip fill in its steady state: affine gaps, unbanded, 8 rows per
// lane, every lane active in every step (no ramp at the strip's ends, no scout captures).  Same cell values as
// stripSteps<true,false,false,8,MODE_TASK,false>; what differs is how the step is issued:
//   * no divergent code: row checkpoints are predicated stores through a per-lane pointer, a lane's column
//     checkpoint (one step in 64) is four predicated 16-byte stores;
//   * lane 0's boundary cells come from a 32-entry shared-memory feed (one broadcast LDS per step) instead of three
//     indexed shuffles, the column's base code from a shared-memory window of one-hot masks (one LDS) instead of a
//     shuffle + shift + multiply;
//   * the loop-carried dependency is ONE op per cell: with e = max(H, D) known off the chain,
//     V(i+1) = max(S(i) + go, V(i) + ge) = max(V(i) + max(go, ge), e(i) + go) because S(i) = max(V(i), e(i)).
// colBase = first column of the chunk (lane t works on column colBase + kk - t in step kk).
// RR rows per lane: 8 for the 256-row strips; 2 or 4 for flat grids of at most 64 / 128 rows (one strip), whose
// fill is one long serial walk over the columns — the step is then a quarter / half as long.
// CAPROW: the strip holds the grid's last row and the scouts track it (final / global matrices): lane capLane also
// stores (S,H,V) of its row capR into lastRow[column] every step.  Lanes below the matrix compute on padding (their
// base mask is empty); nothing they produce is ever read.
// RAMP: the strip's first chunk — lane t joins the wavefront in step t; until then its state is held with selects.
template <int RR, bool CAPROW, bool RAMP>
__device__ __forceinline__ void leanChunk(const StepConsts& K, StripState<RR>& st, int lane, uint32_t leanOff,
                                          const uint8_t* seqH, int colBase, int bS, int bV, int2* rowOut, int2* ckTile,
                                          int capLane, int capR, DCell* lastRow) {
    constexpr int LPB = CKR / RR;    // lanes per 64-row checkpoint block
    int2* bnd = reinterpret_cast<int2*>(gSmem + leanOff);
    uint32_t* hcw = reinterpret_cast<uint32_t*>(gSmem + leanOff + 256);
    // stage: boundary (S,V) of columns colBase..colBase+31, one-hot masks of columns colBase-31..colBase+31
    bnd[lane] = make_int2(bS, bV);
    {
        const int j1 = colBase - 31 + lane;   // (< 1 only in the ramp chunk, where such columns are never used)
        hcw[lane] = (!RAMP || j1 >= 1) ? (1u << seqH[j1 - 1]) * 0x01010101u : 0u;
        if (lane < 31) hcw[32 + lane] = (1u << seqH[colBase + lane]) * 0x01010101u;   // column colBase + 1 + lane
    }
    __syncwarp();
    const int match = K.match, mismatch = K.mismatch, go = K.go, ge = K.ge;
    const int g = max(go, ge);
    const int j0 = colBase - lane;                       // this lane's column in step 0
    const int kkHit = (-j0) & (CKW - 1);                 // step in which the lane reaches a checkpoint column (< 32: in this chunk)
    int2* ckPtr = ckTile + (size_t)((j0 + kkHit) / CKW) * SH + lane * RR;
    int2* rowPtr = rowOut + (size_t)(lane / LPB) * (size_t)(K.nH + 1) + j0;
    const bool rowLane = (lane & (LPB - 1)) == LPB - 1;
    const uint32_t* hp = hcw + (31 - lane);
    uint32_t vm[(RR + 3) / 4];
#pragma unroll
    for (int w = 0; w < (RR + 3) / 4; ++w) vm[w] = st.vm[w];
    int pubS = st.pubS, pubV = st.pubV, prevUpS = st.prevUpS;
    uint32_t hr = 0, hrLast = 0;
#pragma unroll 1   // (the kernel's warps run very different code: a small loop body stays in the instruction cache; unroll 2: no gain)
    for (int kk = 0; kk < 32; ++kk) {
        int inS = __shfl_up_sync(FULLMASK, pubS, 1);
        int inV = __shfl_up_sync(FULLMASK, pubV, 1);
        const int2 b = bnd[kk];
        if (lane == 0) { inS = b.x; inV = b.y; }
        hr = hp[kk];
        const bool act = !RAMP || kk >= lane;
        uint32_t eq[(RR + 3) / 4];
#pragma unroll
        for (int w = 0; w < (RR + 3) / 4; ++w) eq[w] = vm[w] & hr;
        int Sd = prevUpS;
        int vv = __viaddmax_s32(inS, go, inV + ge);      // V of the lane's first row
        int ns = 0, capV = 0;
        int nS[RR], nH[RR];
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int sub = (eq[r >> 2] & (0xffu << (8 * (r & 3)))) ? match : mismatch;
            const int hh = __viaddmax_s32(st.Sl[r], go, st.Hl[r] + ge);
            const int e = __viaddmax_s32(Sd, sub, hh);   // max(D, H)
            ns = max(vv, e);
            if (CAPROW && r == capR) capV = vv;
            Sd = st.Sl[r];
            nS[r] = ns; nH[r] = hh;
            if (r < RR - 1) vv = __viaddmax_s32(vv, g, e + go);   // V of the next row
        }
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            st.Sl[r] = (!RAMP || act) ? nS[r] : st.Sl[r];
            st.Hl[r] = (!RAMP || act) ? nH[r] : st.Hl[r];
        }
        prevUpS = (!RAMP || act) ? inS : prevUpS;
        pubS = (!RAMP || act) ? ns : pubS;
        pubV = (!RAMP || act) ? vv : pubV;
        if (!RAMP || act) hrLast = hr;
        if (rowLane && act) __stcg(&rowPtr[kk], make_int2(ns, vv));
        if (CAPROW && lane == capLane && act) {
            int cs = st.Sl[0], ch = st.Hl[0];
#pragma unroll
            for (int r = 1; r < RR; ++r)
                if (r == capR) { cs = st.Sl[r]; ch = st.Hl[r]; }
            lastRow[j0 + kk] = DCell{cs, ch, capV};
        }
        if (kk == kkHit && act) {
#pragma unroll
            for (int r = 0; r < RR; r += 2)
                __stcg(reinterpret_cast<int4*>(ckPtr + r), make_int4(st.Sl[r], st.Hl[r], st.Sl[r + 1], st.Hl[r + 1]));
        }
    }
    st.pubS = pubS; st.pubV = pubV; st.prevUpS = prevUpS;
    if (hrLast) st.curHc = __ffs((int)(hrLast & 0xffu)) - 1;
    __syncwarp();
}

// One chunk of a trace-tile recompute (64-row block, 2 rows per lane, affine gaps, unbanded): the trace bytes of the
// tile under a traceback path, recomputed from the checkpoints.  Same values and bytes as
// stripSteps<true,CT,false,2,MODE,false>; issued like leanChunk: boundary cells and one-hot column masks come from
// shared memory, and a lane outside the tile in a step (the wavefront's ramps: most of a 64-column tile) is masked
// with selects instead of a divergent branch.  win: column cBeg at offset 0, 32 two-byte slots per column.
template <bool CT, bool GLOBALWIN>
__device__ __forceinline__ void leanTileChunk(const StepConsts& K, StripState<2>& st, int lane, uint32_t leanOff,
                                              const uint8_t* seqH, int cBeg, int cEnd, int c, int i0, int bS, int bV,
                                              int nsteps, uint8_t* win) {
    int2* bnd = reinterpret_cast<int2*>(gSmem + leanOff);
    uint32_t* hcw = reinterpret_cast<uint32_t*>(gSmem + leanOff + 256);
    const int colBase = cBeg + 32 * c;
    bnd[lane] = make_int2(bS, bV);
    {   // one-hot masks of columns colBase-31 .. colBase+31 (those outside [cBeg, cEnd] are never used)
        const int j1 = colBase - 31 + lane, j2 = colBase + 1 + lane;
        hcw[lane] = (j1 >= cBeg && j1 <= cEnd) ? (1u << seqH[j1 - 1]) * 0x01010101u : 0u;
        if (lane < 31) hcw[32 + lane] = (j2 <= cEnd) ? (1u << seqH[j2 - 1]) * 0x01010101u : 0u;
    }
    __syncwarp();
    const int match = K.match, mismatch = K.mismatch, go = K.go, ge = K.ge;
    const uint32_t* hp = hcw + (31 - lane);
    const uint32_t vm0 = st.vm[0];
    const bool rowsInside = i0 <= K.nV;   // lanes below the matrix stay idle
    int pubS = st.pubS, pubV = st.pubV, prevUpS = st.prevUpS, S0 = st.Sl[0], S1 = st.Sl[1], H0 = st.Hl[0], H1 = st.Hl[1];
    uint32_t hr = 0, hrLast = 0;
    uint16_t* wp = reinterpret_cast<uint16_t*>(win) + (size_t)(colBase - lane - cBeg) * 32 + lane;   // slot of step 0
    const int kEnd = imin(32, nsteps - 32 * c);
#pragma unroll 1
    for (int kk = 0; kk < kEnd; ++kk) {
        int inS = __shfl_up_sync(FULLMASK, pubS, 1);
        int inV = __shfl_up_sync(FULLMASK, pubV, 1);
        const int2 b = bnd[kk];
        if (lane == 0) { inS = b.x; inV = b.y; }
        hr = hp[kk];
        const int j = colBase + kk - lane;
        const bool act = rowsInside && (32 * c + kk >= lane) && (j <= cEnd);
        const uint32_t eq = vm0 & hr;
        int s0, h0, v0, s1, h1, v1;
        const uint32_t t0 = cellUpdate<true, CT, false>(s0, h0, v0, S0, H0, inS, inV, prevUpS, (eq & 0xffu) ? match : mismatch, go, ge, 0);
        const uint32_t t1 = cellUpdate<true, CT, false>(s1, h1, v1, S1, H1, s0, v0, S0, (eq & 0xff00u) ? match : mismatch, go, ge, 0);
        if (act) {
            *wp = (uint16_t)(t0 | (t1 << 8));   // (win is based on the shared array or a global slot: the caller rebased it)
            hrLast = hr;
        }
        S0 = act ? s0 : S0; H0 = act ? h0 : H0; S1 = act ? s1 : S1; H1 = act ? h1 : H1;
        prevUpS = act ? inS : prevUpS;
        pubS = act ? s1 : pubS; pubV = act ? v1 : pubV;
        wp += 32;
    }
    st.pubS = pubS; st.pubV = pubV; st.prevUpS = prevUpS; st.Sl[0] = S0; st.Sl[1] = S1; st.Hl[0] = H0; st.Hl[1] = H1;
    if (hrLast) st.curHc = __ffs((int)(hrLast & 0xffu)) - 1;
    __syncwarp();
}

// Runs columns cBeg..cEnd of strip s (rows s*32*RR+1 ...).  fromCk: the lane state left of cBeg comes
// from the column checkpoint at cBeg-1 (a multiple of CKW) instead of the grid's first column.
// nsteps <= (cEnd-cBeg+1)+31 limits the wavefront (partial tile recompute).
//   TRACE : trace bytes of the columns into `win` (column cBeg at offset 0), nothing written to HBM
//           except the scout captures when `capture` is set.
//   !TRACE: score-only; boundary row, column checkpoints, rowProg published per chunk; waits on the
//           strip above through its rowProg counter.
template <bool AFF, bool CT, bool BANDED, int RR, int MODE>
__device__ __noinline__ void runStrip(const GridCtx& Gin, int s, int cBeg, int cEnd, bool fromCk, int nsteps,
                                      bool capture, uint8_t* winIn, int winPitch) {
    const GridCtx& G = *toShared(&Gin);
    uint8_t* win = (MODE == MODE_TASK) ? nullptr : (MODE == MODE_TRACEG ? winIn : toShared(winIn));
    constexpr int SHR = 32 * RR;
    constexpr bool TRACE = (MODE == MODE_TRACE || MODE == MODE_TRACEG);
    constexpr bool L2ONLY = (MODE == MODE_TASK);   // worker warps read the arena through L2 only
    const int lane = threadIdx.x & 31;
    const GridGeom& g = G.g;
    const long long dbg0 = clock64();
    long long dbgSteps = 0, dbgWait = 0;
    bool stripLog = MODE == MODE_TASK && (cP.pad5 & 16);
    int logIdx = 0;
    unsigned long long logT0 = 0, logT1 = 0;
    if (stripLog) {
        if (lane == 0) { logIdx = atomicAdd(&gStripLogN, 1); logT0 = globalTimerNs() - cP.cb->t0; }
        logIdx = __shfl_sync(FULLMASK, logIdx, 0);
        if (logIdx >= 16384) stripLog = false;
    }
    StripState<RR> st;
    const int i0 = s * SHR + lane * RR + 1;
    const int jlo = stripJlo(g, s, SHR);
#pragma unroll
    for (int w4 = 0; w4 < (RR + 3) / 4; ++w4) st.vm[w4] = 0;
    int2* ckTile = nullptr;   // checkpoint tile pointer such that tile of column j is ckTile + (j/CKW)*SH
    // column checkpoints are kept per 256-row strip; a 64-row tile (SHR == CKR) reads its quarter of them
    const int s256 = (s * SHR) / SH;
    const int ckRowOff = (s * SHR) % SH;
    if (MODE == MODE_TASK || fromCk) ckTile = G.colCk + ((size_t)__ldcg(&G.ckBase[s256]) - (size_t)ckFirst(g, s256)) * SH;
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        const int i = i0 + r;
        const uint32_t code = (i <= g.nV) ? (uint32_t)G.seqV[i - 1] : 255u;
        st.vm[r >> 2] |= (code < 8u ? (1u << code) : 0u) << (8 * (r & 3));
    }
    if (!fromCk) {
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int i = i0 + r;
            if (jlo == 1 && i <= G.colZeroMax) {
                const DCell ic = L2ONLY ? ldcgCell(&G.initCol[i]) : G.initCol[i];
                st.Sl[r] = ic.s; st.Hl[r] = ic.h;
            }
            else { st.Sl[r] = NEG_INF; st.Hl[r] = NEG_INF; }
        }
        if (jlo == 1) st.prevUpS = (i0 - 1 <= G.colZeroMax) ? (L2ONLY ? __ldcg(&G.initCol[i0 - 1].s) : G.initCol[i0 - 1].s) : NEG_INF;
        else {
            st.prevUpS = NEG_INF;
            if (lane == 0) {
                if (MODE == MODE_TASK && s > 0) {  // the corner cell above-left of the strip comes from the strip above
                    const int need = imin(jlo - 1, stripJhi(g, s - 1, SHR));
                    while (ldRelaxed(&G.rowProg[s - 1]) < need) __nanosleep(128);
                }
                int bS, bV; upBoundary<BANDED, L2ONLY>(G, s, SHR, jlo - 1, bS, bV); st.prevUpS = bS;
            }
            __syncwarp();
        }
    } else {
        const int2* ck = ckTile + (size_t)((cBeg - 1) / CKW) * SH + ckRowOff;
#pragma unroll
        for (int r = 0; r < RR; ++r) {
            const int2 v = __ldcg(&ck[lane * RR + r]);
            st.Sl[r] = v.x; st.Hl[r] = v.y;
        }
        if (lane == 0) { int bS, bV; upBoundary<BANDED, L2ONLY>(G, s, SHR, cBeg - 1, bS, bV); st.prevUpS = bS; }
        else st.prevUpS = __ldcg(&ck[lane * RR - 1]).x;
    }
    st.pubS = NEG_INF; st.pubV = NEG_INF; st.curHc = 0;
    StepConsts K;
    K.match = G.match; K.mismatch = G.mismatch; K.go = G.go; K.ge = G.ge;
    K.nV = g.nV; K.nH = g.nH; K.lo = g.lo; K.up = g.up;
    int2* rowOut = (MODE == MODE_TASK) ? G.rowCk + (size_t)s * (SH / CKR) * (size_t)(g.nH + 1) : nullptr;
    const int nch = (nsteps + 31) / 32;
    int upProg = 0;                // cached progress of the strip above
    // The strip below becomes claimable the moment this one starts computing (its first wait for the strip above is
    // over): the warp that claims it copies the grid context and loads its first column while this strip walks its
    // first two chunks, then waits for them.  At most one claimed strip per grid is waiting at any time, and a claimed
    // strip only ever waits for strips claimed earlier (no deadlock among the persistent warps).
    bool signalled = false;
    const int upJhi = (s > 0) ? stripJhi(g, s - 1, SHR) : 0;
#pragma unroll 1
    for (int c = 0; c < nch; ++c) {
        const int jj = cBeg + 32 * c + lane;
        if (MODE == MODE_TASK && s > 0) {
            // boundary columns this chunk reads: cBeg+32c .. min(cEnd, cBeg+32c+31), all <= jhi(s-1) when in band
            const int need = imin(imin(cEnd, cBeg + 32 * c + 31), upJhi);
            if (upProg < need) {
                if (lane == 0) {
                    const long long w0 = clock64();
                    int p = ldRelaxed(&G.rowProg[s - 1]);
                    while (p < need) { __nanosleep(128); p = ldRelaxed(&G.rowProg[s - 1]); }
                    upProg = p;
                    dbgWait += clock64() - w0;
                }
                upProg = __shfl_sync(FULLMASK, upProg, 0);
            }
        }
        if (MODE == MODE_TASK && !signalled && (!(cP.pad5 & 64) || c >= 2 || c == nch - 1)) {
            signalled = true;
            if (stripLog && lane == 0) logT1 = globalTimerNs() - cP.cb->t0;
            if (lane == 31) {
                atomicMax(G.readyUpTo, s + 1);
                __threadfence();
                if (s + 1 < G.NS) pushToken(G.taskId);
            }
        }
        int bS = NEG_INF, bV = NEG_INF, hcN = 0;
        if (jj <= cEnd) { hcN = G.seqH[jj - 1]; upBoundary<BANDED, L2ONLY>(G, s, SHR, jj, bS, bV); }
        bool cap = false;
        if (capture) {
            const int jmaxChunk = imin(cEnd, cBeg + 32 * c + 31);  // largest column any lane touches in this chunk
            if (MODE == MODE_FAST) cap = jmaxChunk >= G.fastC0;
            else if (G.capEdges) cap = ((s + 1) * SHR >= g.nV) || (jmaxChunk >= g.nH);
            else cap = (jmaxChunk >= G.hNext) && ((s + 1) * SHR >= G.boxRow0);
        }
        const long long dbg1 = clock64();
        bool lean = false;
        if constexpr (MODE == MODE_TASK && AFF && !BANDED && (RR == 8 || RR == 4 || RR == 2)) {
            // steady state: every lane active in all 32 steps (c >= 1), or the ramp chunk (c == 0) with masked lanes
            const bool full = cBeg + 32 * c + 31 <= cEnd && 32 * c + 31 < nsteps && !(cP.pad5 & 8) && (c >= 1 || !fromCk);
            const uint32_t leanOff = (uint32_t)(SMEM_LEAN + (threadIdx.x >> 5) * LEAN_BYTES);
            const int rl = g.nV - (s * SHR + 1);   // row nV inside the strip: lane rl / RR, row rl % RR
            const bool capRowOnly = capture && G.capEdges && cBeg + 32 * c + 31 < g.nH;   // last strip of a final / global
            if (full && !cap) {                                                           // matrix, away from its last column
                lean = true;
                if (c >= 1) leanChunk<RR, false, false>(K, st, lane, leanOff, G.seqH, cBeg + 32 * c, bS, bV, rowOut, ckTile, 0, 0, nullptr);
                else leanChunk<RR, false, true>(K, st, lane, leanOff, G.seqH, cBeg, bS, bV, rowOut, ckTile, 0, 0, nullptr);
            } else if (full && capRowOnly) {
                lean = true;
                if (c >= 1) leanChunk<RR, true, false>(K, st, lane, leanOff, G.seqH, cBeg + 32 * c, bS, bV, rowOut, ckTile, rl / RR, rl % RR, G.lastRow);
                else leanChunk<RR, true, true>(K, st, lane, leanOff, G.seqH, cBeg, bS, bV, rowOut, ckTile, rl / RR, rl % RR, G.lastRow);
            }
        }
        if constexpr ((MODE == MODE_TRACE || MODE == MODE_TRACEG) && AFF && !BANDED && RR == 2) {
            if (!capture && winPitch == 32 && !(cP.pad5 & 8)) {
                lean = true;
                leanTileChunk<CT, MODE == MODE_TRACEG>(K, st, lane, (uint32_t)(SMEM_LEAN + (threadIdx.x >> 5) * LEAN_BYTES), G.seqH,
                                                       cBeg, cEnd, c, i0, bS, bV, nsteps, win);
            }
        }
        if (lean) {}
        else if (cap)
            stripSteps<AFF, CT, BANDED, RR, MODE, true>(G, K, st, c, lane, cBeg, cEnd, i0, bS, bV, hcN, nsteps, win,
                                                         winPitch, rowOut, ckTile);
        else
            stripSteps<AFF, CT, BANDED, RR, MODE, false>(G, K, st, c, lane, cBeg, cEnd, i0, bS, bV, hcN, nsteps, win,
                                                          winPitch, rowOut, ckTile);
        dbgSteps += clock64() - dbg1;
        if (MODE == MODE_TASK) {
            // lane 31 has finished every column <= cBeg + 32c + 31 - 31 (and cEnd after the last chunk)
            const int done = imin(cEnd, cBeg + 32 * c + imin(31, nsteps - 1 - 32 * c) - 31);
            if (lane == 31 && done >= cBeg) stRelease(&G.rowProg[s], done);
        }
    }
    if (stripLog && lane == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        gStripLog[logIdx][0] = logT0 | ((unsigned long long)s << 48) | ((unsigned long long)(G.taskId & 0xffff) << 32 << 0) * 0ull;
        gStripLog[logIdx][0] = (logT0 & 0xffffffffull) | ((unsigned long long)s << 32) | ((unsigned long long)(G.taskId & 0xffffff) << 40);
        gStripLog[logIdx][1] = (logT1 & 0xffffffffull) | ((unsigned long long)smid << 32) | ((unsigned long long)(threadIdx.x >> 5) << 48);
        gStripLog[logIdx][2] = ((globalTimerNs() - cP.cb->t0) & 0xffffffffull) | ((unsigned long long)(cEnd - cBeg + 1) << 32);
        gStripLog[logIdx][3] = (unsigned long long)dbgWait | ((unsigned long long)(unsigned)dbgSteps << 40) * 0ull;
    }
    if (MODE == MODE_TASK && lane == 0) {
        atomicAdd(&gDbg[12], (unsigned long long)(clock64() - dbg0));
        atomicAdd(&gDbg[13], (unsigned long long)dbgWait);
        atomicAdd(&gDbg[14], 1ull);
        atomicAdd(&gDbg[15], (unsigned long long)(cEnd - cBeg + 1) * 256ull);
    }
    if (MODE != MODE_TASK && lane == 0) {
        const int o = (MODE == MODE_FAST) ? 4 : 0;
        atomicAdd(&gDbg[o + 0], (unsigned long long)(clock64() - dbg0));
        atomicAdd(&gDbg[o + 1], (unsigned long long)dbgSteps);
        atomicAdd(&gDbg[o + 2], (unsigned long long)nsteps);
        atomicAdd(&gDbg[o + 3], 1ull);
    }
}

// ---------------------------------------------------------------------------------------
// traceback (seqan/align/dp_traceback_impl.h, seeds/banded_chain_alignment_traceback.h)
// in SeqAn storage coordinates (col, cv).  Executed by ALL lanes of the control warp with identical
// (warp-uniform) control flow; only lane 0 writes results.  Trace bytes come from the warp's
// shared-memory window: the whole grid for local grids, the recomputed 64 x 64 tile otherwise.
// ---------------------------------------------------------------------------------------
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
    // one trace segment (four ints) with a single bounds check
    __device__ __forceinline__ void put4(int a, int b, int c, int d) {
        if (len + 4 <= cap) {
            if (lane == 0) { int* p = buf + len; p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
        } else overflow = true;
        len += 4;
    }
    __device__ __forceinline__ void patch(int pos, int x) {
        if (pos < cap && lane == 0) buf[pos] = x;
    }
};

template <bool AFF, bool CT, int MODE>
__device__ __forceinline__ void tileDispatch(const GridCtx& G, int s, int cBeg, int cEnd, bool fromCk, int nsteps,
                                             uint8_t* win) {
    if (G.g.banded) runStrip<AFF, CT, true, 2, MODE>(G, s, cBeg, cEnd, fromCk, nsteps, false, win, 32);
    else runStrip<AFF, CT, false, 2, MODE>(G, s, cBeg, cEnd, fromCk, nsteps, false, win, 32);
}

// Recomputes the trace bytes of the tile that holds (i, j): rows of strip s up to the lane owning row i,
// columns from the checkpoint left of j up to j.  Returns (strip, first column, last column, last row).
// MODE_TRACE: into the warp's shared-memory window; MODE_TRACEG: into global memory (tile helpers).
template <int MODE>
__device__ __noinline__ int4 computeTileT(const GridCtx& Gin, uint8_t* win, int i, int j) {
    const GridCtx& G = *toShared(&Gin);
    const GridGeom& g = G.g;
    const int b = (i - 1) / CKR;                    // 64-row block (32 lanes x 2 rows)
    const int jlo = stripJlo(g, b, CKR);
    const int c0 = ((j - 1) / CKW) * CKW;          // checkpoint column left of j (tile = c0+1 .. c0+CKW)
    bool fromCk = true;
    int cBeg = c0 + 1;
    if (c0 < jlo) { fromCk = false; cBeg = jlo; }
    const int laneOfI = ((i - 1) - b * CKR) / 2;
    const int nsteps = (j - cBeg + 1) + laneOfI;
    __syncwarp();
    if (G.affine) { if (G.complete) tileDispatch<true, true, MODE>(G, b, cBeg, j, fromCk, nsteps, win); else tileDispatch<true, false, MODE>(G, b, cBeg, j, fromCk, nsteps, win); }
    else { if (G.complete) tileDispatch<false, true, MODE>(G, b, cBeg, j, fromCk, nsteps, win); else tileDispatch<false, false, MODE>(G, b, cBeg, j, fromCk, nsteps, win); }
    __syncwarp();
    return make_int4(b, cBeg, j, b * CKR + (laneOfI + 1) * 2);
}
__device__ __forceinline__ int4 computeTileFn(const GridCtx& Gin, uint8_t* win, int i, int j) {
    return computeTileT<MODE_TRACE>(Gin, win, i, j);
}

// Pass-1 lazy trace value of cell (i, j): derived from the (S,H,V) of its three neighbours exactly like the
// trace fill would (same cellUpdate).  One out-of-line copy (the walker calls it from many places).
// Bit 31 of the result: a neighbour lies outside the captured box.
// ---- tile helpers: the asking side (a control warp walking a traceback through a big grid) ----
struct TileFetch { int4 t; int4 hist; int myTile, myTag; };

// Cancels (or waits for) the request of this lane's slot; returns true when the slot is free again.
__device__ __forceinline__ bool releaseTileSlot(uint8_t* slots, int lane, int myTag, bool wait) {
    int* st = reinterpret_cast<int*>(slots + (size_t)lane * TILE_SLOT_BYTES);
    int v = atomicCAS(st, myTag * 4, myTag * 4 + 3);
    if (v == myTag * 4 || v == myTag * 4 + 2) return true;
    if (!wait) return false;
    while (ldRelaxed(st) != myTag * 4 + 2) __nanosleep(64);
    return true;
}

// The trace tile that holds (i, j): taken from a helper's slot when one was asked for it in time, recomputed by
// this warp otherwise; then the tiles ahead on the path's diagonal are requested.  myTile / myTag: this lane's slot.
__device__ __noinline__ TileFetch fetchTileFn(const GridCtx& Gin, uint8_t* win, int i, int j, uint8_t* slots, int job, int gi,
                                             int myTile, int myTag, int4 hist) {
    const GridCtx& G = *toShared(&Gin);
    const GridGeom& g = G.g;
    const KParams& P = cP;
    const int lane = threadIdx.x & 31;
    const int b = (i - 1) / CKR, cbk = (j - 1) / CKW;
    const int id = (b << 16) | cbk;
    TileFetch r;
    bool have = false;
    const unsigned match = __ballot_sync(FULLMASK, myTile == id);
    if (match) {
        const int s = __ffs(match) - 1;
        const int tag = __shfl_sync(FULLMASK, myTag, s);
        uint8_t* slot = slots + (size_t)s * TILE_SLOT_BYTES;
        int v = 0;
        if (lane == 0) {
            int* st = reinterpret_cast<int*>(slot);
            v = atomicCAS(st, tag * 4, tag * 4 + 3);          // not claimed yet: take it back
            if (v == tag * 4 + 1) { while ((v = ldRelaxed(st)) != tag * 4 + 2) __nanosleep(64); }
            __threadfence();
            if (v == tag * 4 + 2) atomicAdd(&gDbg[8], 1ull);
            else atomicAdd(&gDbg[10], 1ull);
        }
        v = __shfl_sync(FULLMASK, v, 0);
        if (v == tag * 4 + 2) {
            const int4* src = reinterpret_cast<const int4*>(slot + 64);
            int4* dst = reinterpret_cast<int4*>(toShared(win));
#pragma unroll
            for (int k = 0; k < 8; ++k) dst[k * 32 + lane] = __ldcg(src + k * 32 + lane);
            r.t = __ldcg(reinterpret_cast<const int4*>(slot + 16));
            have = true;
        }
        if (lane == s) myTile = -1;
        __syncwarp();
    }
    if (!have) r.t = computeTileFn(Gin, win, i, j);
    // slots of tiles the (monotone) path has left behind are recycled
    if (myTile >= 0 && ((myTile >> 16) > b || (myTile & 0xffff) > cbk)) {
        if (releaseTileSlot(slots, lane, myTag, false)) myTile = -1;
    }
    // ask for the tiles ahead while idle control warps exist and the ring has room
    int go = 0;
    if (lane == 0)
        go = ldRelaxed(&P.cb->idleHelpers) + (ldRelaxed(&P.cb->openTasks) == 0 ? ldRelaxed(&P.cb->idleWorkers) : ldRelaxed(&P.cb->idleWorkers) / 4) >=
             TILE_HELPERS_PER_WALK * ldRelaxed(&P.cb->activeWalkers);
    go = __shfl_sync(FULLMASK, go, 0);
    if (go) {
        // the path is extrapolated along the direction of its last two tile entries (hist); lane = sample point
        // (48 cells apart) x {on the line, 16 cells to either side of it}
        int di = (hist.z >= 0 ? hist.z : hist.x) - i, dj = (hist.z >= 0 ? hist.w : hist.y) - j;
        if (hist.x < 0 || (di <= 0 && dj <= 0)) { di = 1; dj = 1; }
        di = imax(di, 0); dj = imax(dj, 0);
        const int m = imax(di, dj);
        const int si = di * 48 / m, sj = dj * 48 / m;
        int cand = -1;
        if (lane < TILE_CANDS) {
            const int t = lane / 3 + 1, ty = lane - (lane / 3) * 3;
            const int pi = i - t * si + (ty == 1 ? 16 : (ty == 2 ? -16 : 0));
            const int pj = j - t * sj + (ty == 1 ? -16 : (ty == 2 ? 16 : 0));
            if (pi >= 1 && pj >= 1 && pi <= i && pj <= j) {
                const int tb = (pi - 1) / CKR, tc = (pj - 1) / CKW;
                if ((tb != b || tc != cbk) && pj >= stripJlo(g, tb, CKR) && pj <= stripJhi(g, tb, CKR)) cand = (tb << 16) | tc;
            }
        }
        int pushSlot = -1, pushTag = 0;
        for (int c = 0; c < TILE_CANDS; ++c) {
            const int cid = __shfl_sync(FULLMASK, cand, c);
            if (cid < 0) continue;
            if (__ballot_sync(FULLMASK, myTile == cid)) continue;
            const unsigned freeM = __ballot_sync(FULLMASK, myTile < 0);
            if (!freeM) break;
            const int s = __ffs(freeM) - 1;
            if (lane == s) {
                myTile = cid;
                ++myTag;
                __stcg(reinterpret_cast<int*>(slots + (size_t)s * TILE_SLOT_BYTES), myTag * 4);
            }
            const int tag = __shfl_sync(FULLMASK, myTag, s);
            if (lane == c) { pushSlot = s; pushTag = tag; }
        }
        __syncwarp();
        if (pushSlot >= 0) {
            // one request per ring, starting at a ring that rotates with the tile: helpers pop by compare-and-swap
            const int q = (id + pushSlot * 5 + (int)blockIdx.x) & (TILE_QUEUES - 1);
            ControlBlock::TileQueue* tq = &P.cb->tq[q];
            {
                // (the rings hold far more than the 32 requests a walk can have outstanding: the entry is free)
                const int pos = atomicAdd(&tq->tail, 1);
                atomicAdd(&P.cb->tilePending, 1);
                TileReq* e = &P.tileRing[(size_t)q * TILE_RING_CAP + (pos & (TILE_RING_CAP - 1))];
                const int turn = pos / TILE_RING_CAP;
                while (ldRelaxed(&e->seq) != 2 * turn) __nanosleep(64);
                e->job = job; e->gi = gi; e->tile = cand; e->expect = pushTag * 4; e->pad = (int)(unsigned)globalTimerNs();
                e->slot = (unsigned long long)(slots + (size_t)pushSlot * TILE_SLOT_BYTES);
                __threadfence();
                stRelease(&e->seq, 2 * turn + 1);
            }
        }
        __syncwarp();
    }
    r.myTile = myTile; r.myTag = myTag;
    r.hist = make_int4(i, j, hist.x, hist.y);
    return r;
}

__device__ __forceinline__ DCell fastCellAt(const GridCtx& G, const uint8_t* win, int i, int j, bool& outOfBox) {
    const GridGeom& g = G.g;
    if (g.banded) { const int d = j - i; if (d < g.lo || d > g.up) return DCell{NEG_INF, NEG_INF, NEG_INF}; }
    if (i == 0) return G.initRow[j];
    if (j == 0) return G.initCol[i];
    if (i < G.fastR0 || j < G.fastC0) { outOfBox = true; return DCell{NEG_INF, NEG_INF, NEG_INF}; }
    return reinterpret_cast<const DCell*>(win)[(j - G.fastC0) * G.fastPitch + (i - G.fastR0)];
}
__device__ __noinline__ uint32_t lazyTvFn(const GridCtx& Gin, const uint8_t* winIn, int i, int j) {
    const GridCtx& G = *toShared(&Gin);
    const uint8_t* win = toShared(winIn);
    const GridGeom& g = G.g;
    if (g.banded) { const int d = j - i; if (d < g.lo || d > g.up) return 0; }
    bool oob = false;
    const DCell L = fastCellAt(G, win, i, j - 1, oob), U = fastCellAt(G, win, i - 1, j, oob), D = fastCellAt(G, win, i - 1, j - 1, oob);
    // (i, j) inside the box: the staged base codes cover it
    const int sub = (i >= G.fastR0 && j >= G.fastC0)
                        ? ((toShared(G.fastSeqH)[j - G.fastC0] == toShared(G.fastSeqV)[i - G.fastR0]) ? G.match : G.mismatch)
                        : ((G.seqH[j - 1] == G.seqV[i - 1]) ? G.match : G.mismatch);
    int mode = 0;
    if (g.banded) { const int d = j - i; mode = (d == g.up) ? 1 : (d == g.lo ? 2 : 0); }
    int ns, nh, nv;
    const int go = G.go, ge = G.ge;
    uint32_t tv;
    if (G.affine) {
        if (g.banded) tv = cellUpdate<true, true, true>(ns, nh, nv, L.s, L.h, U.s, U.v, D.s, sub, go, ge, mode);
        else tv = cellUpdate<true, true, false>(ns, nh, nv, L.s, L.h, U.s, U.v, D.s, sub, go, ge, mode);
    } else {
        if (g.banded) tv = cellUpdate<false, true, true>(ns, nh, nv, L.s, L.h, U.s, U.v, D.s, sub, go, ge, mode);
        else tv = cellUpdate<false, true, false>(ns, nh, nv, L.s, L.h, U.s, U.v, D.s, sub, go, ge, mode);
    }
    return tv | (oob ? 0x80000000u : 0u);
}

// TILEONLY: the walker of a big (task) grid — the local-window and lazy modes are compiled out.
template <bool TILEONLY>
struct TraceWalkerT {
    const GridCtx& G;
    OutStream& out;
    const uint8_t* __restrict__ win;     // shared-memory trace window of this control warp
    uint8_t* winW;
    uint32_t winS;                       // the window's 32-bit shared-space address
    // register copies of the hot GridCtx fields (G lives in shared memory)
    const GridGeom g;
    const int local, rrMul, rr, rrs, pitch, localJhi, affine;
    // pass-1 lazy mode: no trace bytes exist; the trace value of a cell is derived on demand from the
    // (S,H,V) of its three neighbours held in the shared-memory box (same cellUpdate as the trace fill)
    bool lazy, outOfBox;
    // cached tile (task grids): strip, first column, valid extent
    int tS, tC0, tMaxRow, tMaxCol;
    int pc, pv;       // navigator position: column, storage row
    int nSegs;        // segments emitted for the current trace
    bool emitOn;
    bool bad;         // undefined trace value (reference: endless loop / assert)
    long long tilesComputed, tileCycles;
    // tile helpers (big-grid pass-2 tracebacks): this warp's slots, and this lane's slot (requested tile, generation)
    uint8_t* hSlots;
    int hJob, hGi, myTile, myTag;
    int4 hHist;       // entry points (row, column) of the last two tiles: the direction the path is heading

    __device__ __forceinline__ TraceWalkerT(const GridCtx& g, OutStream& o, uint8_t* w)
        : G(g), out(o), win(toShared(w)), winW(w), winS((uint32_t)__cvta_generic_to_shared(w)), g(g.g), local(g.local), rrMul(g.rrMul), rr(g.RR), rrs(g.rrs), pitch(g.pitch), localJhi(g.localJhi),
          affine(g.affine), lazy(false), outOfBox(false), tS(-1), tC0(0), tMaxRow(-1), tMaxCol(-1), pc(0), pv(0), nSegs(0), emitOn(true),
          bad(false), tilesComputed(0), tileCycles(0), hSlots(nullptr), hJob(0), hGi(0), myTile(-1), myTag(0), hHist(make_int4(-1, -1, -1, -1)) {}

    __device__ __forceinline__ void enableHelp(uint8_t* slots, int job, int gi) {
        if (g.nV / CKR >= 32767 || g.nH / CKW >= 65535) return;
        // only the long walks are worth it: they end the kernel, and helpers are a shared resource
        if (g.nH + g.nV < TILE_HELP_MIN_EXTENT) return;
        hSlots = slots; hJob = job; hGi = gi; myTile = -1; hHist = make_int4(-1, -1, -1, -1);
        if ((threadIdx.x & 31) == 0) atomicAdd(&cP.cb->activeWalkers, 1);
        myTag = __ldcg(reinterpret_cast<const int*>(slots + (size_t)(threadIdx.x & 31) * TILE_SLOT_BYTES)) >> 2;
        if (myTag < 0 || myTag > (1 << 28)) myTag = 0;
    }
    // every outstanding request is cancelled or waited for: nothing may write into the slots after the walk
    __device__ __forceinline__ void drainHelp() {
        if (hSlots == nullptr) return;
        if (myTile >= 0) { releaseTileSlot(hSlots, threadIdx.x & 31, myTag, true); myTile = -1; }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) atomicSub(&cP.cb->activeWalkers, 1);
        hSlots = nullptr;
    }

    // Recompute the trace bytes of the tile that holds (i, j) (free function: the walker must stay in registers)
    __device__ __forceinline__ void computeTile(int i, int j) {
        const long long t0 = clock64();
        int4 t;
        if (hSlots != nullptr) {
            const TileFetch f = fetchTileFn(G, winW, i, j, hSlots, hJob, hGi, myTile, myTag, hHist);
            t = f.t; myTile = f.myTile; myTag = f.myTag; hHist = f.hist;
        } else {
            t = computeTileFn(G, winW, i, j);
        }
        tS = t.x; tC0 = t.y; tMaxCol = t.z; tMaxRow = t.w;
        ++tilesComputed;
        tileCycles += clock64() - t0;
    }

    __device__ __forceinline__ uint32_t tvHere() {
        if constexpr (TILEONLY) {
            // unbanded task grid (storage = matrix coordinates): a cell of the current tile is one LDS away
            if (!g.banded && tS >= 0) {
                const unsigned ri = (unsigned)(pv - 1 - tS * CKR), cj = (unsigned)(pc - tC0);
                if (ri < (unsigned)(tMaxRow - tS * CKR) && cj <= (unsigned)(tMaxCol - tC0)) return ldsU8(winS + cj * CKR + ri);
            }
        }
        const int i = pv - storageOffset(g, pc);
        const int j = pc;
        if (i <= 0 || j <= 0 || i > g.nV || j > g.nH) return 0;
        if constexpr (!TILEONLY) {
            if (lazy) {
                const uint32_t r = lazyTvFn(G, win, i, j);
                if (r >> 31) outOfBox = true;
                return r & 0xffu;
            }
            if (local) {
                if (j > localJhi) return 0;
                const int q = ((i - 1) * rrMul) >> 16;
                return win[((j - 1) * pitch + q) * rrs + ((i - 1) - q * rr)];
            }
        }
        const int b = (i - 1) / CKR;
        if (b != tS || j < tC0 || j > tMaxCol || i > tMaxRow) {
            // columns outside the block's band range hold no computed cells
            if (j < stripJlo(g, b, CKR) || j > stripJhi(g, b, CKR)) return 0;
            computeTile(i, j);
        }
        return ldsU8(winS + (uint32_t)((j - tC0) * CKR + (i - 1) - b * CKR));
    }
    __device__ __forceinline__ Coord makeCoord(int endCol, int endRow) const {  // dp_traceback_impl.h:121-141
        Coord c;
        c.currCol = pc; c.currRow = pv; c.endCol = endCol; c.endRow = endRow; c.bp1 = 0; c.bp2 = 0;
        c.inBandFlag = false;
        if (g.banded) {
            if (c.currCol > g.up) c.currRow += c.currCol - g.up;
            if (c.endCol > g.up) c.endRow += c.endCol - g.up;
            c.bp1 = imin(g.nH, imax(0, g.up));
            c.bp2 = imin(g.nH, imax(0, g.nV + g.lo));
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
        out.put4(h + out.h0, v + out.v0, len, dir);
        ++nSegs;
    }
    __device__ __forceinline__ void moveH(const Coord& c) { if (c.isInBand()) { --pc; ++pv; } else --pc; }
    __device__ __forceinline__ void moveD(const Coord& c) { if (c.isInBand()) { --pc; } else { --pc; --pv; } }
    __device__ __forceinline__ void moveV() { --pv; }

    __device__ __forceinline__ void doTraceback(uint32_t& tv, uint32_t& last, int& frag, Coord& c) {  // dp_traceback_impl.h:335-431
        const bool aff = affine;
        if (tv & T_D) {
            if (!(last & T_D)) { record(c.currCol, c.currRow, frag, last); last = T_D; frag = 0; }
            // run of diagonal steps (identical to re-entering this branch once per step)
            if (TILEONLY && !g.banded) {
                // unbanded task grid: storage = matrix coordinates.  While the run stays inside the current 64 x 64
                // tile its bytes are read straight from the window: one LDS + test per cell.
                do {
                    const int nT = imin(pv - 1 - tS * CKR, pc - tC0);
                    if (tS >= 0 && nT > 0 && pv - 1 <= tMaxRow && pc - 1 <= tMaxCol) {
                        const int n = imin(nT, imin(c.currCol - c.endCol, c.currRow - c.endRow));
                        uint32_t a = winS + (uint32_t)((pc - tC0) * CKR + (pv - 1 - tS * CKR));
                        int k = 0;
                        do { a -= CKR + 1; tv = ldsU8(a); ++k; } while ((tv & T_D) && k < n);
                        pc -= k; pv -= k; c.currCol -= k; c.currRow -= k; frag += k;
                    } else {
                        --pc; --pv;
                        tv = tvHere();
                        --c.currCol; --c.currRow; ++frag;
                    }
                } while ((tv & T_D) && !c.reachedEnd());
            } else {
                do { moveD(c); tv = tvHere(); --c.currCol; --c.currRow; ++frag; } while ((tv & T_D) && !c.reachedEnd());
            }
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
    __device__ __forceinline__ bool flatWalk() const { return TILEONLY && !g.banded; }
    // The loop `while (!c.reachedEnd() && tv != T_NONE) doTraceback(tv, last, frag, c)` for an unbanded task grid
    // (storage = matrix coordinates, c.currCol == pc, c.currRow == pv) as a flat state machine: every iteration is
    // one move and one trace-byte read, with a single read site (one LDS while the path stays inside the tile).
    // Same decisions, in the same order, as doTraceback (dp_traceback_impl.h:335-431).
    __device__ __forceinline__ void walkFlat(uint32_t& tvIO, uint32_t& lastIO, int& fragIO, Coord& c) {
        enum { ST_DISPATCH = 0, ST_DRUN, ST_VRUN, ST_HRUN, ST_ONE };
        const bool aff = affine;
        const int endCol = c.endCol, endRow = c.endRow;
        int col = pc, row = pv, frag = fragIO;
        uint32_t tv = tvIO, last = lastIO;
        int tRowLo = tS * CKR, tRows = tMaxRow - tS * CKR, tCols = tMaxCol - tC0 + 1;
        if (tS < 0) { tRows = 0; tCols = 0; }
        int state = ST_DISPATCH;
        for (;;) {
            int dr = 0, dc = 0;
            if (state == ST_DISPATCH) {
                if (col <= endCol || row <= endRow || tv == T_NONE) break;
                uint32_t dir;
                if (tv & T_D) { dir = T_D; state = ST_DRUN; }
                else if ((tv & T_MV) && (tv & T_V)) { dir = T_V; state = aff ? ST_VRUN : ST_ONE; dr = 1; }
                else if ((tv & T_MV) && (tv & T_VO)) { dir = T_V; state = ST_ONE; dr = 1; }
                else if ((tv & T_MH) && (tv & T_H)) { dir = T_H; state = aff ? ST_HRUN : ST_ONE; dc = 1; }
                else if ((tv & T_MH) && (tv & T_HO)) { dir = T_H; state = ST_ONE; dc = 1; }
                else { bad = true; tv = T_NONE; break; }
                if (!(last & dir)) { record(col, row, frag, last); last = dir; frag = 0; }
            }
            // runs inside the current tile: as many of the steps below as stay in the tile, one LDS each
            {
                const unsigned ri = (unsigned)(row - 1 - tRowLo), cj = (unsigned)(col - tC0);
                if (state != ST_DISPATCH && state != ST_ONE && ri < (unsigned)tRows && cj < (unsigned)tCols) {
                    uint32_t a = winS + cj * CKR + ri;
                    int k = 0;
                    if (state == ST_DRUN) {
                        // here tv has T_D and the end is not reached: move while that holds
                        const int kTile = imin((int)ri, (int)cj), kEnd = imin(col - endCol, row - endRow);
                        if (kTile >= 1) {
                            do { a -= CKR + 1; tv = ldsU8(a); ++k; } while ((tv & T_D) && k < kEnd && k < kTile);
                            row -= k; col -= k; frag += k;
                            if (!(tv & T_D) || k >= kEnd) { state = ST_DISPATCH; continue; }
                        }
                    } else if (state == ST_VRUN) {
                        while (((tv & (T_VO | T_V)) != T_VO) && row - k != 1 && k < (int)ri) { a -= 1; tv = ldsU8(a); ++k; }
                        row -= k; frag += k;
                    } else {
                        while (((tv & (T_HO | T_H)) != T_HO) && col - k != 1 && k < (int)cj) { a -= CKR; tv = ldsU8(a); ++k; }
                        col -= k; frag += k;
                    }
                }
            }
            if (state == ST_DRUN) { dr = 1; dc = 1; }
            else if (state == ST_VRUN) {
                dr = 1; dc = 0;
                // while ((!(tv & T_VO) || (tv & T_V)) && row != 1) step;  then one more step
                if (!((!(tv & T_VO) || (tv & T_V)) && row != 1)) state = ST_ONE;
            } else if (state == ST_HRUN) {
                dr = 0; dc = 1;
                if (!((!(tv & T_HO) || (tv & T_H)) && col != 1)) state = ST_ONE;
            }
            row -= dr; col -= dc; ++frag;
            {   // the trace byte of (row, col)
                const unsigned ri = (unsigned)(row - 1 - tRowLo), cj = (unsigned)(col - tC0);
                if (ri < (unsigned)tRows && cj < (unsigned)tCols) {
                    tv = ldsU8(winS + cj * CKR + ri);
                } else {
                    pc = col; pv = row;
                    tv = tvHere();
                    tRowLo = tS * CKR; tRows = tMaxRow - tS * CKR; tCols = tMaxCol - tC0 + 1;
                    if (tS < 0) { tRows = 0; tCols = 0; }
                }
            }
            if (state == ST_ONE) state = ST_DISPATCH;
            else if (state == ST_DRUN) { if (!((tv & T_D) && col > endCol && row > endRow)) state = ST_DISPATCH; }
        }
        pc = col; pv = row; c.currCol = col; c.currRow = row;
        tvIO = tv; lastIO = last; fragIO = frag;
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
    __device__ __forceinline__ void generic(bool prefer, bool head, bool tail, int tvOverride) {
        uint32_t tv = tvOverride >= 0 ? (uint32_t)tvOverride : tvHere();
        uint32_t last = initialDirection(tv, prefer);
        Coord c = makeCoord(0, 0);
        const int nH = g.nH, nV = g.nV;
        if (tail) {
            if (c.currRow != nV) record(nH, c.currRow, nV - c.currRow, T_V);
            if (c.currCol != nH) record(c.currCol, c.currRow, nH - c.currCol, T_H);
        }
        int frag = 0;
        if (TILEONLY && !g.banded) walkFlat(tv, last, frag, c);
        else while (!c.reachedEnd() && tv != T_NONE) doTraceback(tv, last, frag, c);
        record(c.currCol, c.currRow, frag, last);
        if (head) {
            if (c.currRow != 0) record(0, 0, c.currRow, T_V);
            if (c.currCol != 0) record(0, 0, c.currCol, T_H);
        }
    }
};
typedef TraceWalkerT<false> TraceWalker;

}  // namespace ub200
