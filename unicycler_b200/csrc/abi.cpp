// extern "C" surface of libunicycler_b200.so (declared in include/unicycler_b200.h).
// Host orchestration only: parsing, seeding, planning, formatting.  All DP runs on the GPU engine;
// if the engine cannot be created the calls fail loudly (no CPU fallback).
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <random>
#include <sstream>
#include <atomic>
#include <thread>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/unicycler_b200.h"
#include "engine.hpp"
#include "host_align.hpp"
#include "hostpool.hpp"
#include "kmerjoin.hpp"
#include "seeding.hpp"

using namespace ub200;
namespace ub200 { extern std::atomic<long long> g_seedProf[6]; extern std::atomic<long long> g_lt[10]; }

typedef std::unordered_map<std::string, std::string> SeqMap;  // include/ref_seqs.h:18

namespace {

int g_device = -1;
std::mutex g_engineMu;
// Engines of this process: two per device (own buffers and stream each), so that a chunked batch call can overlap
// the host stages of one chunk with the kernel of another.  The device list is UNICYCLER_B200_DEVICES ("0,1,2,3" or
// "all"; default: the one device selected by ub200_setDevice / UNICYCLER_B200_DEVICE / the current device): with
// several devices the chunks of a batch call go round the devices, one host thread drives them all (jobs are
// independent and carry their own reference window, so nothing is replicated or exchanged between the GPUs).
// Everything that is not a chunked batch uses engine 0.
std::vector<std::unique_ptr<Engine> > g_engines;
std::vector<int> g_deviceList;
std::atomic<bool> g_lastCallUsedBoth(false);   // the last call was a chunked batch: its counters are in g_batchStats
EngineStats g_batchStats;

void resolveDevices() {   // under g_engineMu
    if (!g_deviceList.empty()) return;
    const char* e = getenv("UNICYCLER_B200_DEVICES");
    if (e && g_device < 0) {
        std::string v(e);
        if (v == "all") {
            const int n = Engine::deviceCount();
            for (int d = 0; d < n; ++d) g_deviceList.push_back(d);
        } else {
            std::stringstream ss(v);
            std::string tok;
            while (getline(ss, tok, ',')) if (!tok.empty()) g_deviceList.push_back(atoi(tok.c_str()));
        }
    }
    if (g_deviceList.empty()) {
        int dev = g_device;
        if (dev < 0) { const char* d = getenv("UNICYCLER_B200_DEVICE"); if (d) dev = atoi(d); }
        g_deviceList.push_back(dev);
    }
    g_engines.resize(2 * g_deviceList.size());
}

int engineCount() {
    std::lock_guard<std::mutex> lock(g_engineMu);
    resolveDevices();
    return (int)g_engines.size();
}

Engine& engine(int slot = 0) {
    std::lock_guard<std::mutex> lock(g_engineMu);
    resolveDevices();
    // slot k: device k % D, buffer set k / D (consecutive chunks of a batch go to different devices)
    const size_t D = g_deviceList.size();
    if (!g_engines[(size_t)slot]) g_engines[(size_t)slot].reset(new Engine(g_deviceList[(size_t)slot % D]));
    return *g_engines[(size_t)slot];
}

// Device k-mer join of the batch path (kmerjoin.hpp), on the first device of the list.
std::unique_ptr<KmerJoiner> g_joiner;
KmerJoiner& joiner() {
    std::lock_guard<std::mutex> lock(g_engineMu);
    resolveDevices();
    if (!g_joiner) g_joiner.reset(new KmerJoiner(g_deviceList[0]));
    return *g_joiner;
}
JoinStats g_lastJoinStats;

[[noreturn]] void fatal(const std::string& msg) {
    fprintf(stderr, "unicycler_b200: fatal: %s\n", msg.c_str());
    fflush(stderr);
    abort();
}

// Request coalescer in front of Engine::run (SURVEY.md 8b "Threading"): the reference's callers invoke the per-read
// ABI from a pool of Python threads (unicycler_align.py:203-225).  Concurrent callers enqueue their jobs; whoever
// finds the engine free becomes the leader, takes EVERYTHING that is pending (group commit) and runs it as one
// device batch; callers that arrive while a batch is running form the next one.  A leader that knows other calls
// are still in their host stage waits a short window (UNICYCLER_B200_COALESCE_US, default 300) for them.
struct RunRequest {
    std::vector<Job*>* jobs;
    bool done = false;
    std::exception_ptr err;
};
std::mutex g_coMu;
std::condition_variable g_coCv;
std::deque<RunRequest*> g_coPending;
bool g_coRunning = false;
std::atomic<int> g_hostStageCalls(0);   // ABI calls currently between entry and submission
std::atomic<long long> g_coBatches(0), g_coRequests(0);

struct HostStageScope {
    HostStageScope() { g_hostStageCalls.fetch_add(1); }
    ~HostStageScope() { leave(); }
    void leave() { if (in) { in = false; g_hostStageCalls.fetch_sub(1); std::lock_guard<std::mutex> lk(g_coMu); g_coCv.notify_all(); } }
    bool in = true;
};

void runCoalesced(std::vector<Job*>& jobs) {
    if (jobs.empty()) return;
    static const long windowUs = [] { const char* e = getenv("UNICYCLER_B200_COALESCE_US"); return e ? atol(e) : 300L; }();
    RunRequest r;
    r.jobs = &jobs;
    std::unique_lock<std::mutex> lk(g_coMu);
    g_coPending.push_back(&r);
    g_coCv.notify_all();
    while (!r.done) {
        if (g_coRunning) { g_coCv.wait(lk); continue; }
        g_coRunning = true;   // leader
        if (windowUs > 0) {
            const auto deadline = std::chrono::steady_clock::now() + std::chrono::microseconds(windowUs);
            while (g_hostStageCalls.load() > 0 && g_coCv.wait_until(lk, deadline) != std::cv_status::timeout) {}
        }
        std::vector<RunRequest*> batch(g_coPending.begin(), g_coPending.end());
        g_coPending.clear();
        lk.unlock();
        std::exception_ptr err;
        g_lastCallUsedBoth = false;
        try {
            if (batch.size() == 1) engine().run(*batch[0]->jobs);
            else {
                std::vector<Job*> all;
                for (RunRequest* q : batch) all.insert(all.end(), q->jobs->begin(), q->jobs->end());
                engine().run(all);
            }
        } catch (...) { err = std::current_exception(); }
        g_coBatches.fetch_add(1);
        g_coRequests.fetch_add((long long)batch.size());
        lk.lock();
        g_coRunning = false;
        for (RunRequest* q : batch) { q->done = true; q->err = err; }
        g_coCv.notify_all();
    }
    lk.unlock();
    if (r.err) std::rethrow_exception(r.err);
}

char* dupString(const std::string& s) {  // cppStringToCString, src/string_functions.cpp:19-24
    char* p = (char*)malloc(s.size() + 1);
    memcpy(p, s.data(), s.size());
    p[s.size()] = '\0';
    return p;
}

// One line on stderr the first time a job is dropped because the reference's own behaviour is undefined for it.
void noteUndefined(int status) {
    static std::atomic<bool> said(false);
    if (!said.exchange(true))
        fprintf(stderr, "unicycler_b200: note: an alignment was dropped (status %d: the reference indexes out of bounds or "
                        "throws for this input; it returns no alignment there too)\n", status);
}

long long nowMs() {
    return std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
}

std::vector<std::string> splitString(const std::string& in, char delim) {  // src/string_functions.cpp:37-49
    std::vector<std::string> result;
    if (in.empty()) return result;
    std::stringstream ss(in);
    while (ss.good()) {
        std::string sub;
        getline(ss, sub, delim);
        result.push_back(sub);
    }
    return result;
}

// Per-read host stages run on the process-wide pool (hostpool.hpp); the reference gets the same parallelism from
// Python threads (unicycler_align.py:203-225).

double nowSec() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ------------------------------------------------------------------ global / path
struct PairJob {
    std::vector<uint8_t> H, V;
    Job job;
    bool planned = false;
    long long startMs = 0;
};

void buildPairJob(PairJob& pj, const char* s1, const char* s2, const Scoring& sc, bool useBanding, int bandSize,
                  bool path) {
    pj.startMs = nowMs();
    toDna5(s1, strlen(s1), pj.H);
    toDna5(s2, strlen(s2), pj.V);
    long lenH = (long)pj.H.size(), lenV = (long)pj.V.size();
    long lo = -bandSize, up = bandSize;
    long diff = lenV - lenH;
    if (!path) {  // src/global_align.cpp:55-66
        if (diff > 0) lo -= diff;
        else if (diff < 0) up -= diff;
    } else {  // src/path_align.cpp:58-63
        if (diff < 0) up -= diff;
    }
    Job& j = pj.job;
    j.H = pj.H.data(); j.lenH = (int32_t)lenH;
    j.V = pj.V.data(); j.lenV = (int32_t)lenV;
    j.match = sc.match; j.mismatch = sc.mismatch; j.gapOpen = sc.gapOpen; j.gapExtend = sc.gapExtend;
    j.freeFirstRow = j.freeFirstCol = j.freeLastRow = 0;
    j.freeLastCol = path ? 1 : 0;  // AlignConfig<false,false,true,false>
    j.complete = 0;                // SingleTrace (AlignConfig2 default, seqan/align/dp_profile.h:335-336)
    pj.planned = planGlobal(lenH, lenV, useBanding, lo, up, false, false, false, path, j.grids);
}

std::string finishPairJob(PairJob& pj, const Scoring& sc, bool path) {
    if (!pj.planned) return "";  // see DESIGN.md: empty inputs / band width < 3 are undefined in the reference
    const JobResult& r = pj.job.result;
    if (r.status == JOB_BAD_SCORE) return "";  // catch (...) -> return 0 (global_align.cpp:69-82)
    if (r.status == JOB_REF_UB || r.status == JOB_INVALID) { noteUndefined(r.status); return ""; }
    if (r.status != JOB_OK) fatal("DP job failed with status " + std::to_string(r.status));
    if (path && r.score < -1000000) return "";  // path_align.cpp:84-85
    const std::vector<Seg>& trace = r.gridTraces[0][0];
    AlignmentRecord rec;
    scoreAlignment(trace, false, pj.H.data(), (long)pj.H.size(), pj.V.data(), (long)pj.V.size(), 0, true, true, !path,
                   sc, rec);
    return fullString(rec, "s1", "s2", nowMs() - pj.startMs);
}

// Unbanded globalAlignment(align, score, AlignConfig<top, left, right, bottom>) as one GRID_GLOBAL job
// (semi_global_align_exhaustive.cpp:40-67, start_end_align.cpp:19-101, overlap_align.cpp:17-81).
// AlignConfig<TTop, TLeft, TRight, TBottom>: free gaps in the first row / first column / last column / last row.
void buildFreeEndJob(PairJob& pj, const std::string& s1, const std::string& s2, const Scoring& sc, bool top, bool left,
                     bool right, bool bottom) {
    pj.startMs = nowMs();
    toDna5(s1.data(), s1.size(), pj.H);
    toDna5(s2.data(), s2.size(), pj.V);
    Job& j = pj.job;
    j.H = pj.H.data(); j.lenH = (int32_t)pj.H.size();
    j.V = pj.V.data(); j.lenV = (int32_t)pj.V.size();
    j.match = sc.match; j.mismatch = sc.mismatch; j.gapOpen = sc.gapOpen; j.gapExtend = sc.gapExtend;
    j.freeFirstRow = top; j.freeFirstCol = left; j.freeLastCol = right; j.freeLastRow = bottom;
    j.complete = 0;   // SingleTrace
    pj.planned = planGlobal((long)pj.H.size(), (long)pj.V.size(), false, 0, 0, top, left, bottom, right, j.grids);
}

// Runs the job; false when the reference's call throws / yields nothing (caught there, mapped to its failure value).
bool runFreeEndJob(PairJob& pj) {
    if (!pj.planned) return false;
    std::vector<Job*> jobs{&pj.job};
    runCoalesced(jobs);
    const JobResult& r = pj.job.result;
    if (r.status == JOB_BAD_SCORE) return false;
    if (r.status == JOB_REF_UB || r.status == JOB_INVALID) { noteUndefined(r.status); return false; }
    if (r.status != JOB_OK) fatal("DP job failed with status " + std::to_string(r.status));
    return true;
}

// Calls f(hasBase1, hasBase2) for every alignment column in order: the two gapped rows the reference streams out of
// the Align object (start_end_align.cpp:64-83), walked straight from the trace segments (stored last segment first).
template <typename F>
long forEachColumn(const std::vector<Seg>& trace, F f) {
    long cols = 0;
    for (size_t k = trace.size(); k > 0; --k) {
        const Seg& s = trace[k - 1];
        for (int t = 0; t < s.len; ++t, ++cols) f(s.dir != T_V, s.dir != T_H);
    }
    return cols;
}

// start_end_align.cpp:30-101
int startEndAlignmentImpl(const char* s1, const char* s2, bool start, const Scoring& sc) {
    HostStageScope hostStage;
    std::string sequence1(s1), sequence2(s2);
    const int trimSize = int(sequence1.length() * 1.5);
    const int trimOffset = start ? 0 : std::max(0, int(sequence2.length()) - trimSize);
    if (int(sequence2.length()) > trimSize) sequence2 = sequence2.substr((size_t)trimOffset, (size_t)trimSize);
    PairJob pj;
    if (start) buildFreeEndJob(pj, sequence1, sequence2, sc, false, false, true, false);
    else buildFreeEndJob(pj, sequence1, sequence2, sc, false, true, false, false);
    hostStage.leave();
    if (!runFreeEndJob(pj)) return -1;
    int seq2Pos = 0, seq2PosAtSeq1Start = -1, seq2PosAtSeq1End = -1;
    const long cols = forEachColumn(pj.job.result.gridTraces[0][0], [&](bool b1, bool b2) {
        if (b1) {
            if (seq2PosAtSeq1Start == -1) seq2PosAtSeq1Start = seq2Pos;
            seq2PosAtSeq1End = seq2Pos + 1;
        }
        if (b2) ++seq2Pos;
    });
    if (cols == 0) return -1;
    return start ? seq2PosAtSeq1End : seq2PosAtSeq1Start + trimOffset;
}

// ------------------------------------------------------------------ chain jobs
struct ChainJob {
    std::vector<uint8_t> H, V;
    std::string readName, refName;
    int refOffset = 0;
    Job job;
    bool planned = false;
    long long startMs = 0;
};

void buildChainJob(ChainJob& cj, const char* readSeq, size_t readLen, const char* refSeq, size_t refLen,
                   const std::vector<ChainSeed>& chain, const Scoring& sc, int bandSize) {
    cj.startMs = nowMs();
    toDna5(readSeq, readLen, cj.H);
    toDna5(refSeq, refLen, cj.V);
    Job& j = cj.job;
    j.H = cj.H.data(); j.lenH = (int32_t)cj.H.size();
    j.V = cj.V.data(); j.lenV = (int32_t)cj.V.size();
    j.match = sc.match; j.mismatch = sc.mismatch; j.gapOpen = sc.gapOpen; j.gapExtend = sc.gapExtend;
    j.freeFirstRow = j.freeFirstCol = j.freeLastRow = j.freeLastCol = 1;  // AlignConfig<true,true,true,true>
    j.complete = 1;  // CompleteTrace (seeds/banded_chain_alignment_profile.h:200-204)
    cj.planned = planChain(chain, (long)cj.H.size(), (long)cj.V.size(), bandSize, j.grids, &j.colTabCount);
}

// returns false when the reference produces no alignment (exception swallowed at semi_global_align.cpp:311)
bool finishChainJob(ChainJob& cj, const Scoring& sc, std::string& out) {
    // Geometry the reference itself cannot align (negative grid origin, empty infix: its unsigned sizes wrap and
    // the resulting bad_alloc is swallowed by the catch (...) at semi_global_align.cpp:297-311): no alignment.
    if (!cj.planned) { noteUndefined(JOB_INVALID); return false; }
    const JobResult& r = cj.job.result;
    if (r.status == JOB_BAD_SCORE) return false;
    if (r.status == JOB_REF_UB || r.status == JOB_INVALID) { noteUndefined(r.status); return false; }
    if (r.status != JOB_OK) fatal("chain DP job failed with status " + std::to_string(r.status));
    std::vector<Seg> trace;
    bool empty = true;
    glueChain(cj.job.grids, r, trace, empty);
    AlignmentRecord rec;
    scoreAlignment(trace, empty, cj.H.data(), (long)cj.H.size(), cj.V.data(), (long)cj.V.size(), cj.refOffset, false,
                   false, false, sc, rec);
    out = fullString(rec, cj.readName, cj.refName, nowMs() - cj.startMs);
    return true;
}

// ------------------------------------------------------------------ semi-global (one read)
struct RangeUnit {  // one (reference, strand, range) of a read: alignReadToReferenceRange, semi_global_align.cpp:135
    std::string refName;
    char strand = '+';
    int refStart = 0, refEnd = 0;
    std::string console;
    std::vector<std::unique_ptr<ChainJob> > jobs;
};
struct ReadWork {
    std::string readName, posSeq, negSeq, console;
    KmerPosMap posKmers, negKmers;
    SensitivityParams sp;
    std::vector<RangeUnit> units;                  // in the reference's iteration order
    std::vector<std::unique_ptr<ChainJob> > jobs;  // in the reference's output order
};

// src/semi_global_align.cpp:608-639
std::pair<int, int> getRefRange(int refStart, int refEnd, int refLen, int readStart, int readEnd, int readLen,
                                bool posStrand) {
    int halfReadLen = 1 + readLen / 2;
    int before = readStart, after = readLen - readEnd;
    if (!posStrand) std::swap(before, after);
    int ns = std::max(0, refStart - before - halfReadLen);
    int ne = std::min(refLen, refEnd + after + halfReadLen);
    return std::make_pair(ns, ne);
}
std::vector<std::pair<int, int> > simplifyRanges(std::vector<std::pair<int, int> > ranges) {
    std::sort(ranges.begin(), ranges.end());
    std::vector<std::pair<int, int> > out;
    std::pair<int, int> cur = ranges[0];
    for (size_t i = 1; i < ranges.size(); ++i) {
        if (cur.second >= ranges[i].first) cur.second = std::max(cur.second, ranges[i].second);
        else { out.push_back(cur); cur = ranges[i]; }
    }
    out.push_back(cur);
    return out;
}

// Host part of semiGlobalAlignment (semi_global_align.cpp:24-142): produces the chain jobs.
void prepareRead(ReadWork& w, const char* readNameC, const char* readSeqC, int verbosity, const char* hitsC,
                 SeqMap* refSeqs, const Scoring& sc, int sensitivityLevel, bool hostKmerIndex = true) {
    typedef std::unordered_map<std::string, std::vector<std::pair<int, int> > > RefRangeMap;
    SensitivityParams sp = sensitivityParams(sensitivityLevel);
    w.readName = readNameC;
    w.posSeq = readSeqC;
    const int readLength = (int)w.posSeq.size();
    std::vector<std::string> hits = splitString(hitsC, ';');
    if (verbosity > 2) {
        w.console += "minimap alignments:\n";
        for (const std::string& h : hits) w.console += "    " + h + "\n";
    }
    RefRangeMap refRanges;
    for (const std::string& hit : hits) {
        std::vector<std::string> p = splitString(hit, ',');
        int readStart = std::stoi(p[0]), readEnd = std::stoi(p[1]);
        char strand = p[2][0];
        std::string refName = p[3];
        int refStart = std::stoi(p[4]), refEnd = std::stoi(p[5]);
        const std::string& refSeq = refSeqs->at(refName);
        std::pair<int, int> r = getRefRange(refStart, refEnd, (int)refSeq.size(), readStart, readEnd, readLength, strand == '+');
        refRanges[refName + strand].push_back(r);
    }
    RefRangeMap simplified;
    for (const auto& r : refRanges) simplified[r.first] = simplifyRanges(r.second);
    if (verbosity > 2) {
        w.console += "Reference ranges:\n";
        for (const auto& r : simplified)
            for (const auto& rr : r.second)
                w.console += "    " + r.first + ": " + std::to_string(rr.first) + " - " + std::to_string(rr.second) + "\n";
    }
    // k-mer indexes of the strands that are needed, and the list of range units in the reference's iteration order
    w.sp = sp;
    bool havePos = false, haveNeg = false;
    for (const auto& r : simplified) {
        std::string refName = r.first;
        const char strand = refName.back();
        refName.pop_back();
        if (strand == '+') {
            if (!havePos) { if (hostKmerIndex) buildKmerPositions(w.posSeq, sp.kSize, w.posKmers); havePos = true; }
        } else if (!haveNeg) {
            w.negSeq = reverseComplement(w.posSeq);
            if (hostKmerIndex) buildKmerPositions(w.negSeq, sp.kSize, w.negKmers);
            haveNeg = true;
        }
        for (const auto& range : r.second) {
            w.units.emplace_back();
            RangeUnit& u = w.units.back();
            u.refName = refName; u.strand = strand; u.refStart = range.first; u.refEnd = range.second;
        }
    }
    (void)sc;
}

// One range unit of a read: seeding + planning of its chain jobs (independent of every other unit).
void seedUnit(ReadWork& w, RangeUnit& u, int verbosity, SeqMap* refSeqs, const Scoring& sc,
              const std::vector<JoinPoint>* joined = nullptr) {
    const std::string& refSeq = refSeqs->at(u.refName);
    const std::string& readSeq = (u.strand == '+') ? w.posSeq : w.negSeq;
    const KmerPosMap& kmers = (u.strand == '+') ? w.posKmers : w.negKmers;
    std::string trimmed = refSeq.substr((size_t)u.refStart, (size_t)(u.refEnd - u.refStart));
    RangeSeeds rs;
    static_assert(sizeof(JoinPoint) == 2 * sizeof(int32_t), "JoinPoint is an (x, y) pair of int32");
    seedRange(readSeq, kmers, trimmed, w.sp, verbosity, u.refName, u.refStart, u.refEnd, rs,
              joined ? (const int32_t*)joined->data() : nullptr, joined ? joined->size() : 0);
    u.console = rs.console;
    const long long tb0 = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
    struct TB { long long t; ~TB() { g_seedProf[5] += std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count() - t; } } tbTimer{tb0};
    for (const auto& chain : rs.chains) {
        std::unique_ptr<ChainJob> cj(new ChainJob());
        cj->readName = w.readName + u.strand;
        cj->refName = u.refName;
        cj->refOffset = u.refStart;
        buildChainJob(*cj, readSeq.data(), readSeq.size(), trimmed.data(), trimmed.size(), chain, sc, w.sp.bandSize);
        u.jobs.push_back(std::move(cj));
    }
}

// Collects the units' results in the reference's order.
void collectUnits(ReadWork& w) {
    for (RangeUnit& u : w.units) {
        w.console += u.console;
        for (auto& cj : u.jobs) w.jobs.push_back(std::move(cj));
        u.jobs.clear();
    }
}

std::string finishRead(ReadWork& w, const Scoring& sc) {
    std::string ret;
    for (auto& cj : w.jobs) {
        std::string s;
        if (finishChainJob(*cj, sc, s)) ret += s + ";";
    }
    ret += w.console;
    return ret;
}

// ------------------------------------------------------------------ forwarding of out-of-scope symbols
void* forwardSym(const char* name) {
    static void* handle = nullptr;
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!handle) {
        const char* path = getenv("UNICYCLER_B200_FORWARD_LIB");
        if (!path) fatal(std::string(name) + " is outside the accelerated path; set UNICYCLER_B200_FORWARD_LIB to the stock cpp_functions.so");
        handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!handle) fatal(std::string("cannot dlopen ") + path + ": " + dlerror());
    }
    void* sym = dlsym(handle, name);
    if (!sym) fatal(std::string("symbol not found in forward library: ") + name);
    return sym;
}

void addStats(EngineStats& s, const EngineStats& t);

// device-resident bench state
struct ChainBench {
    std::vector<std::unique_ptr<ChainJob> > jobs;
    std::vector<Job*> ptrs;
    Scoring sc;
} g_bench;

}  // namespace

extern "C" {

const char* ub200_version(void) { return "unicycler_b200 0.1 (reference ABI: Unicycler 0.5.1)"; }

int ub200_setDevice(int device) {
    // not while a batch is running or queued (the engine would be destroyed under it)
    std::unique_lock<std::mutex> co(g_coMu);
    if (g_coRunning || !g_coPending.empty()) return -1;
    std::lock_guard<std::mutex> lock(g_engineMu);
    g_engines.clear();
    g_joiner.reset();
    g_deviceList.clear();
    g_device = device;
    return 0;
}

void freeCString(char* p) { free(p); }

void* newRefSeqs(void) { return new SeqMap(); }
void addRefSeq(void* h, char* name, char* seq) { ((SeqMap*)h)->emplace(name, seq); }
void deleteRefSeqs(void* h) {
    {   // the device join keeps resident copies keyed by the sequences' host addresses
        std::lock_guard<std::mutex> lock(g_engineMu);
        if (g_joiner) g_joiner->forgetReferences();
    }
    delete (SeqMap*)h;
}

static char* pairAlignment(char* s1, char* s2, int m, int mm, int go, int ge, bool useBanding, int bandSize, bool path) {
    Scoring sc{m, mm, go, ge};
    PairJob pj;
    HostStageScope hostStage;
    buildPairJob(pj, s1, s2, sc, useBanding, bandSize, path);
    hostStage.leave();
    if (pj.planned) {
        std::vector<Job*> jobs{&pj.job};
        runCoalesced(jobs);
    }
    return dupString(finishPairJob(pj, sc, path));
}

char* fullyGlobalAlignment(char* s1, char* s2, int m, int mm, int go, int ge, bool useBanding, int bandSize) {
    return pairAlignment(s1, s2, m, mm, go, ge, useBanding, bandSize, false);
}

char* pathAlignment(char* s1, char* s2, int m, int mm, int go, int ge, bool useBanding, int bandSize) {
    return pairAlignment(s1, s2, m, mm, go, ge, useBanding, bandSize, true);
}

int ub200_globalAlignmentBatch(int n, const char* const* s1, const char* const* s2, int mode, int m, int mm, int go,
                               int ge, bool useBanding, int bandSize, char** results) {
    Scoring sc{m, mm, go, ge};
    const bool path = mode == 1;
    std::vector<std::unique_ptr<PairJob> > pjs((size_t)n);
    std::vector<Job*> jobs;
    parallelFor(n, [&](int i) {
        pjs[(size_t)i].reset(new PairJob());
        buildPairJob(*pjs[(size_t)i], s1[i], s2[i], sc, useBanding, bandSize, path);
    });
    for (int i = 0; i < n; ++i)
        if (pjs[(size_t)i]->planned) jobs.push_back(&pjs[(size_t)i]->job);
    runCoalesced(jobs);
    parallelFor(n, [&](int i) { results[i] = dupString(finishPairJob(*pjs[(size_t)i], sc, path)); });
    return 0;
}

char* getRandomSequenceAlignmentScores(int seqLength, int n, int m, int mm, int go, int ge) {
    // src/random_alignments.cpp:30-52: n unbanded global alignments of uniform random ACGT pairs
    Scoring sc{m, mm, go, ge};
    std::mt19937 gen;
    const char* seedEnv = getenv("UNICYCLER_B200_SEED");
    if (seedEnv) gen.seed((unsigned)strtoul(seedEnv, nullptr, 10));
    else { std::random_device rd; gen.seed(rd()); }
    std::uniform_int_distribution<int> dist(0, 3);
    static const char bases[4] = {'A', 'C', 'G', 'T'};
    std::vector<double> scores;
    // device batches bounded by checkpoint memory (0.25 B per cell): ~1.6e11 cells, i.e. ~40 GB, per batch
    const long long cellsPerPair = (long long)(seqLength + 1) * (seqLength + 1);
    long long perBatch = std::max(1LL, std::min<long long>(n, 160000000000LL / std::max(1LL, cellsPerPair)));
    perBatch = std::min(perBatch, 65536LL);
    for (long long done = 0; done < n; done += perBatch) {
        const long long cnt = std::min<long long>(perBatch, n - done);
        // the random stream is drawn in the reference's order (s1 then s2 of each pair); everything else is
        // independent per pair and runs on the host cores
        std::vector<std::string> as((size_t)cnt, std::string((size_t)seqLength, 'A')), bs((size_t)cnt, std::string((size_t)seqLength, 'A'));
        for (long long i = 0; i < cnt; ++i) {
            for (int k = 0; k < seqLength; ++k) as[(size_t)i][(size_t)k] = bases[dist(gen)];
            for (int k = 0; k < seqLength; ++k) bs[(size_t)i][(size_t)k] = bases[dist(gen)];
        }
        std::vector<std::unique_ptr<PairJob> > pjs((size_t)cnt);
        parallelFor((int)cnt, [&](int i) {
            pjs[(size_t)i].reset(new PairJob());
            buildPairJob(*pjs[(size_t)i], as[(size_t)i].c_str(), bs[(size_t)i].c_str(), sc, false, 0, false);
        });
        std::vector<Job*> jobs;
        for (long long i = 0; i < cnt; ++i)
            if (pjs[(size_t)i]->planned) jobs.push_back(&pjs[(size_t)i]->job);
        runCoalesced(jobs);
        std::vector<double> batchScores((size_t)cnt, 0.0);
        std::vector<char> have((size_t)cnt, 0);
        parallelFor((int)cnt, [&](int i) {
            PairJob& pj = *pjs[(size_t)i];
            if (!pj.planned || pj.job.result.status != JOB_OK) return;  // "if (alignment != 0)"
            AlignmentRecord rec;
            scoreAlignment(pj.job.result.gridTraces[0][0], false, pj.H.data(), (long)pj.H.size(), pj.V.data(),
                           (long)pj.V.size(), 0, true, true, true, sc, rec);
            batchScores[(size_t)i] = rec.scaledScore;
            have[(size_t)i] = 1;
        });
        for (long long i = 0; i < cnt; ++i)
            if (have[(size_t)i]) scores.push_back(batchScores[(size_t)i]);
    }
    double mean = 0.0, sd = 0.0;  // getMeanAndStDev :187-202 (population sd)
    if (!scores.empty()) {
        for (double v : scores) mean += v;
        mean /= (double)scores.size();
        double dev = 0.0;
        for (double v : scores) dev += (v - mean) * (v - mean);
        sd = sqrt(dev / (double)scores.size());
    }
    return dupString(std::to_string(mean) + "," + std::to_string(sd));
}

int ub200_calibrationPairs(int seqLength, int n, unsigned seed, char** s1, char** s2) {
    std::mt19937 gen(seed);
    std::uniform_int_distribution<int> dist(0, 3);
    static const char bases[4] = {'A', 'C', 'G', 'T'};
    std::string a((size_t)std::max(0, seqLength), 'A'), b(a);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < seqLength; ++k) a[(size_t)k] = bases[dist(gen)];
        for (int k = 0; k < seqLength; ++k) b[(size_t)k] = bases[dist(gen)];
        s1[i] = dupString(a);
        s2[i] = dupString(b);
    }
    return 0;
}

// SURVEY.md 8(f)4 — Alignment.tally_up_score_and_errors (unicycler/alignment.py:142-216) walks every CIGAR base by base
// in Python under the GIL; this is the same walk on the host in C++.  readSeq is the read as aligned (already reverse-
// complemented for '-' alignments), refSeq the whole reference sequence; positions as in the result string.
// Returns "matches,mismatches,insertions,deletions,rawScore,alignmentLength,percentIdentity,scaledScore" (doubles with 17
// significant digits: they round-trip), or "" when nothing but soft clips is left (the Python returns early there).
char* ub200_alignmentTallies(const char* readSeq, const char* refSeq, int readStartPos, int refStartPos, const char* cigar,
                             int m, int mm, int go, int ge) {
    const long readLen = (long)strlen(readSeq), refLen = (long)strlen(refSeq);
    std::vector<std::pair<long, char> > parts;
    for (const char* p = cigar; *p;) {
        char* end = nullptr;
        const long n = strtol(p, &end, 10);
        if (end == p || !*end) break;
        parts.emplace_back(n, *end);
        p = end + 1;
    }
    if (!parts.empty() && parts.front().second == 'S') parts.erase(parts.begin());
    if (!parts.empty() && parts.back().second == 'S') parts.pop_back();
    if (parts.empty()) return dupString("");
    long matches = 0, mismatches = 0, insertions = 0, deletions = 0, raw = 0, alignI = 0;
    long readI = readStartPos, refI = refStartPos;
    for (const auto& part : parts) {
        const long n = part.first;
        long score = 0;
        if (part.second == 'I') { score = go + (n - 1) * ge; insertions += n; readI += n; }
        else if (part.second == 'D') { score = go + (n - 1) * ge; deletions += n; refI += n; }
        else {
            for (long k = 0; k < n; ++k) {
                if (readI >= readLen || refI >= refLen) break;
                if (readSeq[readI] == refSeq[refI]) { ++matches; score += m; }
                else { ++mismatches; score += mm; }
                ++readI; ++refI;
            }
        }
        raw += score;
        alignI += n;
    }
    const double identity = 100.0 * (double)matches / (double)alignI;
    const long perfect = (long)m * alignI, worst = (long)mm * alignI;
    const double scaled = 100.0 * (double)(raw - worst) / (double)(perfect - worst);
    char buf[256];
    snprintf(buf, sizeof(buf), "%ld,%ld,%ld,%ld,%ld,%ld,%.17g,%.17g", matches, mismatches, insertions, deletions, raw, alignI,
             identity, scaled);
    return dupString(buf);
}

void ub200_coalescerStats(int64_t* batches, int64_t* requests) {
    if (batches) *batches = g_coBatches.load();
    if (requests) *requests = g_coRequests.load();
}

static void seedsFromArray(const int64_t* seeds, int nSeeds, std::vector<ChainSeed>& chain) {
    chain.resize((size_t)nSeeds);
    for (int i = 0; i < nSeeds; ++i)
        chain[(size_t)i] = ChainSeed{(long)seeds[6 * i], (long)seeds[6 * i + 1], (long)seeds[6 * i + 2],
                                     (long)seeds[6 * i + 3], (long)seeds[6 * i + 4], (long)seeds[6 * i + 5]};
}

char* ub200_chainAlignment(const char* readSeq, const char* refSeq, const int64_t* seeds, int nSeeds, int m, int mm,
                           int go, int ge, int bandSize, const char* readName, const char* refName, int refOffset) {
    Scoring sc{m, mm, go, ge};
    std::vector<ChainSeed> chain;
    seedsFromArray(seeds, nSeeds, chain);
    ChainJob cj;
    cj.readName = readName; cj.refName = refName; cj.refOffset = refOffset;
    buildChainJob(cj, readSeq, strlen(readSeq), refSeq, strlen(refSeq), chain, sc, bandSize);
    std::vector<Job*> jobs;
    if (cj.planned) jobs.push_back(&cj.job);
    runCoalesced(jobs);
    std::string out;
    if (!finishChainJob(cj, sc, out)) out = "";
    return dupString(out);
}

static int buildChainBatch(int n, const char* const* readSeqs, const char* const* refSeqs, const int64_t* seeds,
                           const int64_t* seedOffsets, const Scoring& sc, int bandSize, const char* const* readNames,
                           const char* const* refNames, const int* refOffsets,
                           std::vector<std::unique_ptr<ChainJob> >& cjs, std::vector<Job*>& jobs) {
    cjs.resize((size_t)n);
    jobs.clear();
    for (int i = 0; i < n; ++i) {
        std::vector<ChainSeed> chain;
        seedsFromArray(seeds + 6 * seedOffsets[i], (int)(seedOffsets[i + 1] - seedOffsets[i]), chain);
        cjs[(size_t)i].reset(new ChainJob());
        ChainJob& cj = *cjs[(size_t)i];
        cj.readName = readNames ? readNames[i] : "read+";
        cj.refName = refNames ? refNames[i] : "ref";
        cj.refOffset = refOffsets ? refOffsets[i] : 0;
        buildChainJob(cj, readSeqs[i], strlen(readSeqs[i]), refSeqs[i], strlen(refSeqs[i]), chain, sc, bandSize);
        if (cj.planned) jobs.push_back(&cj.job);   // unplanned: no alignment, like the reference's swallowed exception
    }
    return 0;
}

int ub200_chainAlignmentBatch(int n, const char* const* readSeqs, const char* const* refSeqs, const int64_t* seeds,
                              const int64_t* seedOffsets, int m, int mm, int go, int ge, int bandSize,
                              const char* const* readNames, const char* const* refNames, const int* refOffsets,
                              char** results) {
    Scoring sc{m, mm, go, ge};
    std::vector<std::unique_ptr<ChainJob> > cjs;
    std::vector<Job*> jobs;
    if (buildChainBatch(n, readSeqs, refSeqs, seeds, seedOffsets, sc, bandSize, readNames, refNames, refOffsets, cjs, jobs))
        return -1;
    runCoalesced(jobs);
    parallelFor(n, [&](int i) {
        std::string out;
        if (!finishChainJob(*cjs[(size_t)i], sc, out)) out = "";
        results[i] = dupString(out);
    });
    return 0;
}

int ub200_chainBenchPrepare(int n, const char* const* readSeqs, const char* const* refSeqs, const int64_t* seeds,
                            const int64_t* seedOffsets, int m, int mm, int go, int ge, int bandSize) {
    g_bench.sc = Scoring{m, mm, go, ge};
    if (buildChainBatch(n, readSeqs, refSeqs, seeds, seedOffsets, g_bench.sc, bandSize, nullptr, nullptr, nullptr,
                        g_bench.jobs, g_bench.ptrs))
        return -1;
    engine().upload(g_bench.ptrs);
    return 0;
}

double ub200_chainBenchRun(void) {
    Engine& e = engine();
    e.launch();
    // kernel time is read back in finish(); here we only synchronise via a cheap stats fetch
    return 0.0;
}

double ub200_chainBenchRunSteps(int steps) { return engine().launchTimed(steps); }

}  // extern "C"
namespace {
void addStats(EngineStats& s, const EngineStats& t) {
    s.kernelMs += t.kernelMs; s.h2dMs += t.h2dMs; s.d2hMs += t.d2hMs; s.cells += t.cells; s.launches += t.launches;
    s.traceBytes += t.traceBytes; s.h2dBytes += t.h2dBytes; s.d2hBytes += t.d2hBytes; s.ctas = t.ctas;
}
EngineStats lastCallStats() { return g_lastCallUsedBoth ? g_batchStats : engine().lastStats(); }
}  // namespace
extern "C" {

void ub200_lastTransferBytes(int64_t* h2d, int64_t* d2h, int64_t* traceBytes, int* ctas) {
    EngineStats s = lastCallStats();
    if (h2d) *h2d = s.h2dBytes;
    if (d2h) *d2h = s.d2hBytes;
    if (traceBytes) *traceBytes = s.traceBytes;
    if (ctas) *ctas = s.ctas;
}

// Reference DP-cell count of one banded-chain alignment (planner only; needs no GPU).
int64_t ub200_chainCells(int readLen, int refLen, const int64_t* seeds, int nSeeds, int bandSize, int* nGrids) {
    std::vector<ChainSeed> chain;
    chain.resize((size_t)nSeeds);
    for (int i = 0; i < nSeeds; ++i)
        chain[(size_t)i] = ChainSeed{(long)seeds[6 * i], (long)seeds[6 * i + 1], (long)seeds[6 * i + 2],
                                     (long)seeds[6 * i + 3], (long)seeds[6 * i + 4], (long)seeds[6 * i + 5]};
    std::vector<GridDesc> grids;
    if (!planChain(chain, readLen, refLen, bandSize, grids)) return -1;
    int64_t cells = 0;
    for (const GridDesc& g : grids) cells += referenceCells(g);
    if (nGrids) *nGrids = (int)grids.size();
    return cells;
}

// Planner introspection: writes 10 int32 per sub-DP (kind, nH, nV, banded, lo, up, h0, v0, hNext, vNext); returns the grid count.
int ub200_chainPlan(int readLen, int refLen, const int64_t* seeds, int nSeeds, int bandSize, int32_t* out, int cap) {
    std::vector<ChainSeed> chain;
    seedsFromArray(seeds, nSeeds, chain);
    std::vector<GridDesc> grids;
    if (!planChain(chain, readLen, refLen, bandSize, grids)) return -1;
    for (size_t k = 0; k < grids.size() && (int)k < cap; ++k) {
        const GridDesc& g = grids[k];
        int32_t* o = out + 10 * k;
        o[0] = g.kind; o[1] = g.nH; o[2] = g.nV; o[3] = g.banded; o[4] = g.lo; o[5] = g.up; o[6] = g.h0; o[7] = g.v0;
        o[8] = g.hNext; o[9] = g.vNext;
    }
    return (int)grids.size();
}

int ub200_chainBenchFinish(char** results) {
    Engine& e = engine();
    e.fetch(g_bench.ptrs);
    for (size_t i = 0; i < g_bench.jobs.size(); ++i) {
        std::string out;
        if (!finishChainJob(*g_bench.jobs[i], g_bench.sc, out)) out = "";
        if (results) results[i] = dupString(out);
    }
    return 0;
}

char* semiGlobalAlignment(char* readName, char* readSeq, int verbosity, char* hits, void* refSeqs, int m, int mm,
                          int go, int ge, double, bool, int sensitivityLevel) {
    Scoring sc{m, mm, go, ge};
    ReadWork w;
    HostStageScope hostStage;   // lets a coalescing leader know that this call is about to submit
    prepareRead(w, readName, readSeq, verbosity, hits, (SeqMap*)refSeqs, sc, sensitivityLevel);
    parallelFor((int)w.units.size(), [&](int k) { seedUnit(w, w.units[(size_t)k], verbosity, (SeqMap*)refSeqs, sc); });
    collectUnits(w);
    std::vector<Job*> jobs;
    for (auto& cj : w.jobs)
        if (cj->planned) jobs.push_back(&cj->job);
    hostStage.leave();
    runCoalesced(jobs);
    return dupString(finishRead(w, sc));
}

int ub200_semiGlobalAlignmentBatch(int n, const char* const* readNames, const char* const* readSeqs,
                                   const char* const* hits, void* refSeqs, int m, int mm, int go, int ge,
                                   int sensitivityLevel, char** results) {
    Scoring sc{m, mm, go, ge};
    // One batch call at a time per process: a chunked call holds several engines at once (stage .. end), and two such
    // calls taking them in different order would wait for each other.  Per-read calls (request coalescer) hold one
    // engine for one launch and interleave freely.
    static std::mutex batchMu;
    std::lock_guard<std::mutex> batchLock(batchMu);
    const double t0 = nowSec();
    std::vector<std::unique_ptr<ReadWork> > works((size_t)n);
    // largest reads first: their chains have the longest spines
    std::vector<int> order((size_t)n);
    for (int i = 0; i < n; ++i) order[(size_t)i] = i;
    std::vector<size_t> len((size_t)n);
    for (int i = 0; i < n; ++i) len[(size_t)i] = strlen(readSeqs[i]);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return len[(size_t)a] > len[(size_t)b]; });

    // Host seeding + planning of the reads order[k0 .. k1): one task per read (ranges, k-mer index), then one task per
    // (read, reference range), most expensive first.  Returns the chunk's device jobs.
    // Common k-mers of the whole chunk are joined on the device (kmerjoin.cu); UNICYCLER_B200_HOST_KMERS=1 keeps the
    // host join of the per-read entry point (developer switch, same points in the same order).
    const bool hostOnly = getenv("UNICYCLER_B200_HOST_ONLY") != nullptr;
    const bool deviceJoin = !hostOnly && !getenv("UNICYCLER_B200_HOST_KMERS");
    JoinStats joinStats;
    double joinMs = 0.0;
    // Per chunk: the range units in seeding order and their common k-mer points (filled one chunk ahead, see below).
    struct ChunkSeeds {
        std::vector<std::pair<int, int> > unitList;
        std::vector<std::vector<JoinPoint> > joined;
    };
    // Stage 1 of the reads order[k0 .. k1): ranges of every read, then the k-mer join of all their units.
    auto joinChunk = [&](int k0, int k1, ChunkSeeds& cs) {
        parallelFor(k1 - k0, [&](int k) {
            const int i = order[(size_t)(k0 + k)];
            works[(size_t)i].reset(new ReadWork());
            prepareRead(*works[(size_t)i], readNames[i], readSeqs[i], 0, hits[i], (SeqMap*)refSeqs, sc, sensitivityLevel,
                        !deviceJoin);
        });
        std::vector<std::pair<int, int> >& unitList = cs.unitList;
        unitList.clear();
        cs.joined.clear();
        for (int k = k0; k < k1; ++k) {
            const int i = order[(size_t)k];
            for (size_t u = 0; u < works[(size_t)i]->units.size(); ++u) unitList.emplace_back(i, (int)u);
        }
        std::stable_sort(unitList.begin(), unitList.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
            const RangeUnit& ua = works[(size_t)a.first]->units[(size_t)a.second];
            const RangeUnit& ub = works[(size_t)b.first]->units[(size_t)b.second];
            const double ca = (double)len[(size_t)a.first] * (ua.refEnd - ua.refStart);
            const double cb = (double)len[(size_t)b.first] * (ub.refEnd - ub.refStart);
            return ca > cb;
        });
        if (deviceJoin && !unitList.empty()) {
            const double tj = nowSec();
            std::vector<JoinSeq> seqs;
            std::vector<JoinTask> tasks(unitList.size());
            std::unordered_map<const std::string*, int> seqIndex;
            for (size_t q = 0; q < unitList.size(); ++q) {
                ReadWork& w = *works[(size_t)unitList[q].first];
                const RangeUnit& u = w.units[(size_t)unitList[q].second];
                const std::string& strandSeq = (u.strand == '+') ? w.posSeq : w.negSeq;
                auto it = seqIndex.find(&strandSeq);
                if (it == seqIndex.end()) {
                    it = seqIndex.emplace(&strandSeq, (int)seqs.size()).first;
                    seqs.push_back(JoinSeq{strandSeq.data(), (int)strandSeq.size()});
                }
                const std::string& refSeq = ((SeqMap*)refSeqs)->at(u.refName);
                tasks[q] = JoinTask{it->second, refSeq.data(), refSeq.size(), u.refStart, u.refEnd - u.refStart};
            }
            joiner().run(seqs, tasks, works[(size_t)unitList[0].first]->sp.kSize, cs.joined);
            const JoinStats js = joiner().lastStats();
            joinStats.kernelMs += js.kernelMs; joinStats.launches += js.launches; joinStats.h2dBytes += js.h2dBytes;
            joinStats.d2hBytes += js.d2hBytes; joinStats.points += js.points; joinStats.refUploads += js.refUploads;
            joinMs += (nowSec() - tj) * 1e3;
        }
    };
    // Stage 2: one task per (read, reference range), most expensive first: line tracing, seeds, chain, plan.
    // Returns the chunk's device jobs.
    auto traceChunk = [&](int k0, int k1, ChunkSeeds& cs, std::vector<Job*>& jobs) {
        parallelFor((int)cs.unitList.size(), [&](int k) {
            ReadWork& w = *works[(size_t)cs.unitList[(size_t)k].first];
            seedUnit(w, w.units[(size_t)cs.unitList[(size_t)k].second], 0, (SeqMap*)refSeqs, sc,
                     deviceJoin ? &cs.joined[(size_t)k] : nullptr);
        });
        cs = ChunkSeeds();
        jobs.clear();
        for (int k = k0; k < k1; ++k) {
            ReadWork& w = *works[(size_t)order[(size_t)k]];
            collectUnits(w);
            for (auto& cj : w.jobs)
                if (cj->planned) jobs.push_back(&cj->job);
        }
    };
    // Formatting: one task per chain job (gluing + CIGAR of a long chain costs a millisecond: a read with six of them
    // would otherwise be one serial task), longest chains first; then the reads' strings are put together.
    auto finishChunk = [&](int k0, int k1) {
        std::vector<std::pair<int, int> > jobList;   // (read, job of the read)
        for (int k = k0; k < k1; ++k) {
            const int i = order[(size_t)k];
            for (size_t q = 0; q < works[(size_t)i]->jobs.size(); ++q) jobList.emplace_back(i, (int)q);
        }
        std::stable_sort(jobList.begin(), jobList.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
            return works[(size_t)a.first]->jobs[(size_t)a.second]->job.grids.size() >
                   works[(size_t)b.first]->jobs[(size_t)b.second]->job.grids.size();
        });
        std::vector<std::vector<std::string> > parts((size_t)n);
        for (int k = k0; k < k1; ++k) parts[(size_t)order[(size_t)k]].resize(works[(size_t)order[(size_t)k]]->jobs.size());
        parallelFor((int)jobList.size(), [&](int k) {
            const int i = jobList[(size_t)k].first, q = jobList[(size_t)k].second;
            std::string str;
            if (finishChainJob(*works[(size_t)i]->jobs[(size_t)q], sc, str)) parts[(size_t)i][(size_t)q] = str + ";";
        });
        parallelFor(k1 - k0, [&](int k) {
            const int i = order[(size_t)(k0 + k)];
            std::string ret;
            for (const std::string& part : parts[(size_t)i]) ret += part;
            ret += works[(size_t)i]->console;
            results[i] = dupString(ret);
            works[(size_t)i].reset();   // (thousands of small trace vectors per read: released here, in parallel, not by the caller's thread at return)
        });
    };

    if (getenv("UNICYCLER_B200_HOST_ONLY")) {  // developer aid: time the host stage without a GPU
        std::vector<Job*> jobs;
        ChunkSeeds cs;
        joinChunk(0, n, cs);
        traceChunk(0, n, cs, jobs);
        fprintf(stderr, "[ub200 host] reads=%d jobs=%zu prepare=%.1f ms (kmers %.1f, linetrace %.1f [fillCloud %.1f, densest point %.1f], seeds+chain %.1f, job build + plan %.1f thread-ms)\n", n,
                jobs.size(), (nowSec() - t0) * 1e3, g_seedProf[0] / 1e6, g_seedProf[1] / 1e6, g_seedProf[3] / 1e6, g_seedProf[4] / 1e6, g_seedProf[2] / 1e6, g_seedProf[5] / 1e6);
        for (int q = 0; q < 6; ++q) g_seedProf[q] = 0;
        fprintf(stderr, "[ub200 host] line tracer thread-ms: range tree %.1f, trace loop %.1f [searches %.1f, near set %.1f, scoring %.1f, collection %.1f], set score %.1f, used points %.1f\n",
                g_lt[7] / 1e6, g_lt[4] / 1e6, g_lt[0] / 1e6, g_lt[1] / 1e6, g_lt[2] / 1e6, g_lt[3] / 1e6, g_lt[5] / 1e6, g_lt[6] / 1e6);
        for (int q = 0; q < 10; ++q) g_lt[q] = 0;
        for (int i = 0; i < n; ++i) results[i] = dupString("");
        return 0;
    }

    // Small batches are latency bound on the device (the spines of the longest chains): one launch.  Large batches
    // are cut into chunks that alternate between the engines and run as the pipeline below.
    int chunk = n;
    if (const char* e = getenv("UNICYCLER_B200_CHUNK_READS")) chunk = std::max(1, atoi(e));
    else if (n >= 256) {
        // ~2.5 Mbp of reads per chunk (128 reads of 20 kb: measured optimum of the 512- and 2 048-read synthetic slices,
        // 64 / 96 / 128 / 171 reads per chunk: 2 311 / 2 446 / 2 368 / 2 291 reads/s; 128 / 256 / 512: 2 923 / 2 795 / 2 573),
        // at least four chunks
        size_t bases = 0;
        for (int i = 0; i < n; ++i) bases += len[(size_t)i];
        const double avg = std::max(1.0, (double)bases / n);
        chunk = (int)std::min(1024.0, std::max(64.0, 2.5e6 / avg));
        chunk = std::max(32, std::min(chunk, (n + 3) / 4));
    }
    else if (n >= 64) {   // fewer but long reads (a rank's share of a sharded long-read set): the host stages of such a
        size_t bases = 0;   // call take far longer than its kernels, which a four-chunk pipeline hides
        for (int i = 0; i < n; ++i) bases += len[(size_t)i];
        if (bases >= 1000000) chunk = std::max(32, (n + 3) / 4);
    }
    if (chunk == n && n >= 8 && n < 256 && !getenv("UNICYCLER_B200_CHUNK_READS")) {
        // Few host threads (one rank of several on a box): tracing the lighter half of the reads takes longer than the
        // kernel of the heavier half (~11 ms whatever the batch: a critical path), so two launches hide one of them.
        // With many threads the second launch would only wait for the first (profiles/r2_summary.md, section 7).
        size_t bases = 0;
        for (int i = 0; i < n; ++i) bases += len[(size_t)i];
        if (bases / (size_t)hostThreads() >= 40000) chunk = (n + 1) / 2;   // (sample_data: up to 6 threads)
    }
    const int nChunks = (n + chunk - 1) / chunk;
    std::vector<std::vector<Job*> > jobs((size_t)nChunks);
    auto lo = [&](int k) { return k * chunk; };
    auto hi = [&](int k) { return std::min(n, (k + 1) * chunk); };
    double seedMs = 0.0, finishMs = 0.0;
    EngineStats batchStats;
    const int E = engineCount();   // two per device
    // Pipeline over the chunks.  The line tracing of chunk k+1 runs on the host pool (started from a helper thread)
    // while this thread stages, launches, fetches and formats around chunk k.  The DP kernel owns every SM while it
    // runs, so the k-mer join of a chunk must not queue up behind one with the pool waiting for it: chunk k+2 is
    // joined by this thread between staging and launching chunk k (it waits there for kernel k-1 at most, while the
    // pool is busy tracing chunk k+1).
    //   pool    [trace 0]         [trace 1 ...........][trace 2 ...........]
    //   main    [join 0] [join 1] [stage 0][join 2]     [fetch+fmt 0][stage 1][join 3] ...
    //   GPU                                [kernel 0 ....]                    [kernel 1 ....]
    std::vector<ChunkSeeds> seeds(3);
    struct AsyncTrace {
        std::thread th;
        std::exception_ptr err;
        double ms = 0.0;
    } tracer[2];
    auto startTrace = [&](int k) {
        AsyncTrace& a = tracer[k & 1];
        a.th = std::thread([&, k] {
            try {
                const double ts = nowSec();
                traceChunk(lo(k), hi(k), seeds[(size_t)(k % 3)], jobs[(size_t)k]);
                a.ms = (nowSec() - ts) * 1e3;
            } catch (...) {
                a.err = std::current_exception();
            }
        });
    };
    auto waitTrace = [&](int k) {
        AsyncTrace& a = tracer[k & 1];
        if (a.th.joinable()) a.th.join();
        if (a.err) std::rethrow_exception(a.err);
        seedMs += a.ms;
        a.ms = 0.0;
    };
    auto timedJoin = [&](int k) {
        const double ts = nowSec();
        joinChunk(lo(k), hi(k), seeds[(size_t)(k % 3)]);
        seedMs += (nowSec() - ts) * 1e3;
    };
    timedJoin(0);
    startTrace(0);
    if (nChunks > 1) timedJoin(1);
    for (int k = 0; k < nChunks; ++k) {
        waitTrace(k);
        if (k + 1 < nChunks) startTrace(k + 1);
        if (k >= E) {   // the engine of chunk k is the one chunk k-E used
            engine(k % E).end(jobs[(size_t)(k - E)]);
            addStats(batchStats, engine(k % E).lastStats());
            const double tf = nowSec();
            finishChunk(lo(k - E), hi(k - E));
            finishMs += (nowSec() - tf) * 1e3;
        }
        engine(k % E).stage(jobs[(size_t)k]);
        if (k + 2 < nChunks) timedJoin(k + 2);
        engine(k % E).start();
    }
    for (int k = std::max(0, nChunks - E); k < nChunks; ++k) {
        engine(k % E).end(jobs[(size_t)k]);
        addStats(batchStats, engine(k % E).lastStats());
        const double tf = nowSec();
        finishChunk(lo(k), hi(k));
        finishMs += (nowSec() - tf) * 1e3;
    }
    g_batchStats = batchStats;
    g_batchStats.launches += joinStats.launches;
    g_batchStats.h2dBytes += joinStats.h2dBytes;
    g_batchStats.d2hBytes += joinStats.d2hBytes;
    g_lastJoinStats = joinStats;
    g_lastCallUsedBoth = true;
    if (getenv("UNICYCLER_B200_PROFILE")) {
        fprintf(stderr, "[ub200 host] seeding thread-ms: kmers %.1f, linetrace %.1f [start cloud %.1f, densest point %.1f], seeds+chain %.1f\n",
                g_seedProf[0] / 1e6, g_seedProf[1] / 1e6, g_seedProf[3] / 1e6, g_seedProf[4] / 1e6, g_seedProf[2] / 1e6);
        for (int q = 0; q < 6; ++q) g_seedProf[q] = 0;
    }
    if (getenv("UNICYCLER_B200_PROFILE"))
        fprintf(stderr, "[ub200 host] reads=%d in %d chunk(s) of %d: seeding %.1f ms (of which device k-mer join %.2f ms wall, %.3f ms kernels, %lld points), formatting %.1f ms, total %.1f ms\n",
                n, nChunks, chunk, seedMs, joinMs, joinStats.kernelMs, (long long)joinStats.points, finishMs, (nowSec() - t0) * 1e3);
    return 0;
}

char* ub200_seedChains(const char* readSeq, const char* trimmedRefSeq, int sensitivityLevel) {
    SensitivityParams sp = sensitivityParams(sensitivityLevel);
    std::string read(readSeq), ref(trimmedRefSeq);
    KmerPosMap kmers;
    buildKmerPositions(read, sp.kSize, kmers);
    RangeSeeds rs;
    seedRange(read, kmers, ref, sp, 0, "ref", 0, (int)ref.size(), rs);
    std::string out = std::to_string(rs.chains.size()) + ";";
    for (const auto& chain : rs.chains) {
        out += std::to_string(chain.size()) + ":";
        for (size_t i = 0; i < chain.size(); ++i) {
            const ChainSeed& s = chain[i];
            if (i) out += "|";
            out += std::to_string(s.beginH) + "," + std::to_string(s.beginV) + "," + std::to_string(s.endH) + "," +
                   std::to_string(s.endV) + "," + std::to_string(s.lowerDiag) + "," + std::to_string(s.upperDiag);
        }
        out += ";";
    }
    return dupString(out);
}

// Common k-mer points of one (read strand, reference window) pair: where = 0 the host join of the per-read entry point,
// where = 1 the device join of the batch path (the reference sequence is `ref`, the window [refStart, refStart+refLen)).
// Writes up to cap (x, y) pairs; returns the number of points, or -1 when the device join is not available.
int64_t ub200_commonKmers(const char* readSeq, const char* ref, int refStart, int refLen, int k, int where,
                          int32_t* xy, int64_t cap) {
    std::vector<int32_t> pts;
    if (where == 0) {
        commonKmerPoints(std::string(readSeq), std::string(ref + refStart, (size_t)refLen), k, pts);
    } else {
        std::vector<JoinSeq> seqs{JoinSeq{readSeq, (int)strlen(readSeq)}};
        std::vector<JoinTask> tasks{JoinTask{0, ref, strlen(ref), refStart, refLen}};
        std::vector<std::vector<JoinPoint> > out;
        try {
            joiner().run(seqs, tasks, k, out);
        } catch (const std::exception& e) {
            fprintf(stderr, "unicycler_b200: %s\n", e.what());
            return -1;
        }
        for (const JoinPoint& p : out[0]) { pts.push_back(p.x); pts.push_back(p.y); }
        joiner().forgetReferences();   // `ref` is caller memory that may be gone after the call
    }
    const int64_t n = (int64_t)pts.size() / 2;
    for (int64_t q = 0; q < std::min(n, cap); ++q) { xy[2 * q] = pts[(size_t)(2 * q)]; xy[2 * q + 1] = pts[(size_t)(2 * q + 1)]; }
    return n;
}

void ub200_lastJoinStats(double* kernelMs, int64_t* launches, int64_t* points, int64_t* h2dBytes, int64_t* d2hBytes) {
    if (kernelMs) *kernelMs = g_lastJoinStats.kernelMs;
    if (launches) *launches = g_lastJoinStats.launches;
    if (points) *points = g_lastJoinStats.points;
    if (h2dBytes) *h2dBytes = g_lastJoinStats.h2dBytes;
    if (d2hBytes) *d2hBytes = g_lastJoinStats.d2hBytes;
}

double ub200_intPeakOpsPerSec(void) { return measureIntPeak(engine().device()); }

void ub200_lastStats(int64_t* cells, double* kernelMs, int64_t* launches, double* h2dMs, double* d2hMs) {
    EngineStats s = lastCallStats();
    if (cells) *cells = s.cells;
    if (kernelMs) *kernelMs = s.kernelMs;
    if (launches) *launches = s.launches;
    if (h2dMs) *h2dMs = s.h2dMs;
    if (d2hMs) *d2hMs = s.d2hMs;
}

// ---- the remaining SeqAn pairwise DPs, on the same GRID_GLOBAL kernel path
char* semiGlobalAlignmentExhaustive(char* s1, char* s2, int m, int mm, int go, int ge) {
    // src/semi_global_align_exhaustive.cpp:40-67: AlignConfig<true,true,true,true>, ScoredAlignment(..., true, true, true)
    Scoring sc{m, mm, go, ge};
    HostStageScope hostStage;
    PairJob pj;
    buildFreeEndJob(pj, s1, s2, sc, true, true, true, true);
    hostStage.leave();
    if (!runFreeEndJob(pj)) return dupString("");
    AlignmentRecord rec;
    scoreAlignment(pj.job.result.gridTraces[0][0], false, pj.H.data(), (long)pj.H.size(), pj.V.data(), (long)pj.V.size(), 0,
                   true, true, true, sc, rec);
    return dupString(fullString(rec, "s1", "s2", nowMs() - pj.startMs));
}
int startAlignment(char* s1, char* s2, int m, int mm, int go, int ge) {  // src/start_end_align.cpp:19-21
    return startEndAlignmentImpl(s1, s2, true, Scoring{m, mm, go, ge});
}
int endAlignment(char* s1, char* s2, int m, int mm, int go, int ge) {    // src/start_end_align.cpp:24-26
    return startEndAlignmentImpl(s1, s2, false, Scoring{m, mm, go, ge});
}
char* overlapAlignment(char* s1, char* s2, int m, int mm, int go, int ge, int guessOverlap) {
    // src/overlap_align.cpp:17-81: suffix of s1 against prefix of s2, AlignConfig<true,false,true,false>
    Scoring sc{m, mm, go, ge};
    HostStageScope hostStage;
    std::string sequence1(s1), sequence2(s2);
    const int trimSize = int((guessOverlap + 100) * 1.5);
    if (trimSize < int(sequence1.length())) sequence1 = sequence1.substr(sequence1.length() - (size_t)trimSize, (size_t)trimSize);
    if (trimSize < int(sequence2.length())) sequence2 = sequence2.substr(0, (size_t)trimSize);
    PairJob pj;
    buildFreeEndJob(pj, sequence1, sequence2, sc, true, false, true, false);
    hostStage.leave();
    if (!runFreeEndJob(pj)) return dupString("-1,-1");
    int seq1Pos = 0, seq2Pos = 0, seq1PosAtSeq2Start = -1, seq2PosAtSeq1End = -1;
    const long cols = forEachColumn(pj.job.result.gridTraces[0][0], [&](bool b1, bool b2) {
        if (b1) seq2PosAtSeq1End = seq2Pos + 1;
        if (b2 && seq1PosAtSeq2Start == -1) seq1PosAtSeq2Start = seq1Pos;
        if (b1) ++seq1Pos;
        if (b2) ++seq2Pos;
    });
    if (cols == 0) return dupString("-1,-1");
    return dupString(std::to_string(seq1Pos - seq1PosAtSeq2Start) + "," + std::to_string(seq2PosAtSeq1End));
}

// ---- forwarders for the symbols outside the hot path (cpp_wrappers.py:180-357)
char* multipleSequenceAlignment(char** a, char** b, unsigned long c, unsigned int d, int e, int f, int g, int h) {
    typedef char* (*F)(char**, char**, unsigned long, unsigned int, int, int, int, int);
    return ((F)forwardSym("multipleSequenceAlignment"))(a, b, c, d, e, f, g, h);
}
char* minimapAlignReads(char* a, char* b, int c, int d, int e) {
    typedef char* (*F)(char*, char*, int, int, int);
    return ((F)forwardSym("minimapAlignReads"))(a, b, c, d, e);
}
char* minimapAlignReadsWithSettings(char* a, char* b, int c, bool d, int e, int f, float g, int h, int i, int j, int k) {
    typedef char* (*F)(char*, char*, int, bool, int, int, float, int, int, int, int);
    return ((F)forwardSym("minimapAlignReadsWithSettings"))(a, b, c, d, e, f, g, h, i, j, k);
}
void miniasmAssembly(char* a, char* b, char* c, int d) {
    typedef void (*F)(char*, char*, char*, int);
    ((F)forwardSym("miniasmAssembly"))(a, b, c, d);
}
char* getRandomSequenceAlignmentErrorRates(int a, int b, int c, int d, int e, int f) {
    typedef char* (*F)(int, int, int, int, int, int);
    return ((F)forwardSym("getRandomSequenceAlignmentErrorRates"))(a, b, c, d, e, f);
}
char* simulateDepths(int* a, int b, int c, int d, int e) {
    typedef char* (*F)(int*, int, int, int, int);
    return ((F)forwardSym("simulateDepths"))(a, b, c, d, e);
}

}  // extern "C"
