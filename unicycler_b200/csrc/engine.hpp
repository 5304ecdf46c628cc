// B200 DP engine: host-visible descriptors and the launcher interface.
//
// One "job" is one pairwise alignment as the reference's C ABI sees it (one
// fullyGlobalAlignment / pathAlignment call, or one bandedChainAlignment inside
// semiGlobalAlignment).  A job is a short list of "grids" (sub-DPs): exactly the
// sequence of SeqAn _computeAlignment calls the reference would make
// (seeds/banded_chain_alignment_impl.h:1212-1296), planned on the host because the
// geometry does not depend on DP results.  The device fills each grid (affine or linear
// max-plus recurrences, trace bytes to HBM), finds the tied maxima, walks the traceback
// and hands the crossing cells to the next grid — all inside one persistent kernel.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "dpgeom.hpp"

namespace ub200 {

static const int NEG_INF = INT32_MIN / 2;  // seqan/align/dp_cell.h:122-124

// trace bits, seqan/align/dp_profile.h:146-153
enum : uint8_t { T_NONE = 0, T_D = 1, T_H = 2, T_V = 4, T_HO = 8, T_VO = 16, T_MH = 32, T_MV = 64 };

enum GridKind : int32_t {
    GRID_CHAIN_INITIAL = 0,  // BandedChainInitialDPMatrix
    GRID_CHAIN_INNER = 1,    // BandedChainInnerDPMatrix
    GRID_CHAIN_FINAL = 2,    // BandedChainFinalDPMatrix
    GRID_GLOBAL = 3          // GlobalAlignment_<FreeEndGaps_> with the default scout
};

enum GlueMode : int32_t {
    GLUE_APPEND = 0,         // traces are appended straight to the global set (initial rectangle)
    GLUE_ASSIGN = 1,         // globalTraceSet = localTraceSet
    GLUE_IF_NONEMPTY = 2,    // if (!empty(local)) _glueTracebacks(global, local)
    GLUE_ALWAYS = 3          // _glueTracebacks(global, local)
};

struct GridDesc {
    int32_t kind;
    int32_t h0, v0;          // offset of the infixes in the job's sequences (grid origin, global coords)
    int32_t nH, nV;          // infix lengths
    int32_t banded, lo, up;  // DPBandConfig
    int32_t hNext, vNext;    // _reinitScoutState origin of the next grid (navigator coordinates)
    int32_t capNextH, capNextV;  // lengths of _horizontalInitNextMatrix / _verticalInitNextMatrix
    int32_t plantZerosH, plantZerosV;  // >0: _initiaizeBeginningOfBandedChain(sizeH, sizeV) precedes this grid
    int32_t glue;
    int32_t checkScore;      // 1: the explicit "score < -1000000 -> throw" after this grid (always true inside too)
    // banded chain grids: the column descriptors of _computeBandedAlignment for the columns >= hNext
    // (host-side BandWalker), index into the job's / batch's ColInfo pool
    int32_t colTabOff, nColTab;
    // big grids (filled by worker items): the grid's persistent block (checkpoints, progress counters, init
    // row/column) inside the batch's persistent buffer, set by Engine::upload; -1: none (per-agent arena)
    int64_t persistOff;
    int32_t ckTiles;         // number of 256-row column-checkpoint tiles of the grid
    int32_t pad;             // big grids: task board (0 = served first), set by Engine::upload from the remaining spine latency
};

struct Seg {                 // seqan/align/dp_trace_segment.h (TraceSegment_)
    int32_t hBeg, vBeg, len, dir;
};

enum JobStatus : int32_t {
    JOB_OK = 0,
    JOB_BAD_SCORE = 1,       // reference throws "Bad Seqan alignment score" -> no alignment
    JOB_OUT_OVERFLOW = 2,    // segment buffer too small: host retries with a larger one
    JOB_REF_UB = 3,          // reference would index out of bounds (undefined there); reported loudly
    JOB_INVALID = 4          // _isValidDPSettings false
};

struct JobResult {
    int32_t status = JOB_INVALID;
    int32_t score = 0;
    // per grid: local trace sets exactly as the reference's _computeTraceback would produce them
    std::vector<std::vector<std::vector<Seg>>> gridTraces;
    // for GRID_GLOBAL single-grid jobs the one trace is gridTraces[0][0]
};

struct Job {
    const uint8_t* H = nullptr;  // Dna5 codes 0..4, host memory
    int32_t lenH = 0;
    const uint8_t* V = nullptr;
    int32_t lenV = 0;
    int32_t match = 0, mismatch = 0, gapOpen = 0, gapExtend = 0;
    int32_t freeFirstRow = 0, freeFirstCol = 0, freeLastRow = 0, freeLastCol = 0;
    int32_t complete = 0;        // CompleteTrace (chain) vs SingleTrace (global/path)
    std::vector<GridDesc> grids;
    long long colTabCount = 0;    // column descriptors of the job's banded chain grids (GridDesc::colTabOff is relative to the job; generated on the device)
    int32_t outScale = 1;         // multiplier of the segment-stream capacity (raised when a run overflowed)
    JobResult result;
    // DP cells as the reference counts them (dimH*dimV per sub-DP)
    int64_t cells = 0;
};

struct EngineStats {
    double kernelMs = 0.0;       // CUDA-event time of the DP kernel(s) of the last run()
    double h2dMs = 0.0, d2hMs = 0.0;
    int64_t cells = 0;
    int64_t launches = 0;
    int64_t traceBytes = 0;      // bytes of trace written (algorithmic 1 B/cell incl. padding)
    int64_t h2dBytes = 0, d2hBytes = 0;
    int ctas = 0;
};

// Reference cell count of one grid (seqan/align/dp_algorithm_impl.h:1547-1560).
int64_t referenceCells(const GridDesc& g);

class Engine {
public:
    // device < 0: current device.  Throws std::runtime_error when no CUDA device is usable:
    // there is deliberately no CPU fallback.
    explicit Engine(int device = -1);
    ~Engine();
    // Runs all jobs to completion (results in job.result).  Thread-safe (serialised).
    void run(std::vector<Job*>& jobs);
    // The same in two halves, so that a caller can overlap host work (seeding the next chunk, formatting the previous
    // one) with the kernel: begin() stages, uploads and launches and returns at once; end() waits for the kernel,
    // fetches the results and reruns jobs whose output stream overflowed.  The engine stays locked between the two
    // calls (same thread).
    void begin(std::vector<Job*>& jobs);
    void end(std::vector<Job*>& jobs);
    // begin() in two steps: stage() locks, stages and uploads, start() launches.  A caller puts other device work
    // (the k-mer join of the next chunk) between the two: once the persistent DP kernel is queued nothing else gets an SM.
    void stage(std::vector<Job*>& jobs);
    void start();
    EngineStats lastStats() const;
    int device() const;
    static int deviceCount();   // CUDA devices visible to the process (0 when there is none)
    // Other device allocations of this process (k-mer join buffers, resident references) tell the engines that their
    // cached view of the free device memory is stale.
    static void noteDeviceAllocation();
    // Device-resident benchmark mode: upload()+plan once, then launch() repeatedly with
    // inputs already in HBM; fetch() copies results back.
    void upload(std::vector<Job*>& jobs);
    void launch();
    // launches `steps` times back to back; returns the CUDA-event time of all launches in ms
    double launchTimed(int steps);
    void fetch(std::vector<Job*>& jobs);

private:
    struct Impl;
    Impl* impl_;
};

// Integer-pipe microbenchmark: sustained int32 ops/s over all SMs (roofline denominator).
double measureIntPeak(int device);

}  // namespace ub200
