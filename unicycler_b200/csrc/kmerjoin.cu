// Device k-mer join (see kmerjoin.hpp).  sm_100a; byte / integer work, HBM- and latency-bound: coalesced
// grid-stride passes over flat position arrays, grids sized in multiples of the SM count.
#include "kmerjoin.hpp"
#include "engine.hpp"
#include "hostpool.hpp"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>

namespace ub200 {

namespace {

#define JOIN_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) throw std::runtime_error(std::string("unicycler_b200 k-mer join: ") + #x + ": " + cudaGetErrorString(e_)); } while (0)

struct JoinParams {
    // read strands
    const uint8_t* seq;        // concatenated bytes
    const int64_t* seqOff;     // [nSeq + 1] byte offsets
    const int64_t* posOff;     // [nSeq + 1] prefix of k-mer start positions (len - k + 1, >= 0)
    const int64_t* slotOff;    // [nSeq + 1] prefix of table sizes (powers of two, 0 for a strand without k-mers)
    int nSeq, k;
    int32_t* rep;              // per slot: a read position holding the slot's k-mer, -1 = empty
    int32_t* cnt;              // per slot: occurrences of the k-mer in the strand
    int32_t* start;            // per slot: first index of its position list (relative to the strand's list)
    int32_t* fill;             // per slot: fill cursor
    int32_t* slotOf;           // per read k-mer position: its slot (relative to the strand's table)
    int32_t* tmp;              // per read k-mer position: the lists in arrival order
    int32_t* list;             // per read k-mer position: the lists in ascending order
    // tasks
    const uint8_t* const* taskRef;  // [nTask] device pointer to the window's first base
    const int32_t* taskSeq;    // [nTask]
    const int64_t* taskPosOff; // [nTask + 1] prefix of window k-mer positions
    const int64_t* outOff;     // [nTask] first output point of the task
    int nTask;
    int32_t* refCnt;           // per window position: number of points it emits
    int32_t* refSlot;          // per window position: slot of its k-mer
    int32_t* refOff;           // per window position: exclusive prefix of refCnt inside the task
    int64_t* taskTotal;        // [nTask]
    JoinPoint* out;
    int64_t totalPos, totalSlots, totalRefPos;
};

__device__ __forceinline__ uint32_t hashBytes(const uint8_t* p, int k) {   // FNV-1a + a final mix
    uint32_t h = 2166136261u;
    for (int j = 0; j < k; ++j) { h ^= p[j]; h *= 16777619u; }
    h ^= h >> 15;
    h *= 2654435761u;
    return h ^ (h >> 13);
}
__device__ __forceinline__ bool sameBytes(const uint8_t* a, const uint8_t* b, int k) {
    for (int j = 0; j < k; ++j)
        if (a[j] != b[j]) return false;
    return true;
}
// largest s with off[s] <= g (off ascending, off[0] = 0, off[n] > g)
__device__ __forceinline__ int segmentOf(const int64_t* off, int n, int64_t g) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= g) lo = mid; else hi = mid;
    }
    return lo;
}

// every read k-mer claims / finds its slot and counts itself
__global__ void __launch_bounds__(256) insertKernel(JoinParams P) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < P.totalPos; g += (int64_t)gridDim.x * blockDim.x) {
        const int s = segmentOf(P.posOff, P.nSeq, g);
        const int i = (int)(g - P.posOff[s]);
        const uint8_t* base = P.seq + P.seqOff[s];
        const uint8_t* p = base + i;
        const int64_t tab = P.slotOff[s];
        const uint32_t mask = (uint32_t)(P.slotOff[s + 1] - tab) - 1u;
        uint32_t sl = hashBytes(p, P.k) & mask;
        for (;;) {
            const int old = atomicCAS(&P.rep[tab + sl], -1, i);
            if (old == -1 || old == i || sameBytes(base + old, p, P.k)) break;
            sl = (sl + 1) & mask;
        }
        atomicAdd(&P.cnt[tab + sl], 1);
        P.slotOf[g] = (int32_t)sl;
    }
}

// Exclusive scan of in[segOff[b] .. segOff[b+1]) into out, one block per segment; totals[b] = the segment's sum.
__global__ void __launch_bounds__(1024) segmentScanKernel(const int32_t* in, int32_t* out, const int64_t* segOff, int64_t* totals) {
    __shared__ int32_t warpSum[32];
    __shared__ int32_t carryS;
    const int64_t b0 = segOff[blockIdx.x], b1 = segOff[blockIdx.x + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carryS = 0;
    __syncthreads();
    for (int64_t t0 = b0; t0 < b1; t0 += 4096) {
        const int64_t q = t0 + (int64_t)threadIdx.x * 4;
        int32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (q + j < b1) ? in[q + j] : 0;
        const int32_t mine = v[0] + v[1] + v[2] + v[3];
        int32_t inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t o = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += o;
        }
        if (lane == 31) warpSum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int32_t w = warpSum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int32_t o = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += o;
            }
            warpSum[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const int32_t carry = carryS;
        int32_t ex = carry + (warp ? warpSum[warp - 1] : 0) + inc - mine;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (q + j < b1) out[q + j] = ex;
            ex += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carryS = carry + warpSum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0 && totals) totals[blockIdx.x] = carryS;
}

// positions drop into their k-mer's list in arrival order
__global__ void __launch_bounds__(256) fillKernel(JoinParams P) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < P.totalPos; g += (int64_t)gridDim.x * blockDim.x) {
        const int s = segmentOf(P.posOff, P.nSeq, g);
        const int64_t q = P.slotOff[s] + P.slotOf[g];
        const int r = atomicAdd(&P.fill[q], 1);
        P.tmp[P.posOff[s] + P.start[q] + r] = (int32_t)(g - P.posOff[s]);
    }
}

// ... and are put in ascending order: every position counts the smaller ones of its list (lists are short)
__global__ void __launch_bounds__(256) rankKernel(JoinParams P) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < P.totalPos; g += (int64_t)gridDim.x * blockDim.x) {
        const int s = segmentOf(P.posOff, P.nSeq, g);
        const int64_t q = P.slotOff[s] + P.slotOf[g];
        const int i = (int)(g - P.posOff[s]);
        const int c = P.cnt[q];
        const int64_t l0 = P.posOff[s] + P.start[q];
        int rank = 0;
        for (int r = 0; r < c; ++r) rank += P.tmp[l0 + r] < i;
        P.list[l0 + rank] = i;
    }
}

// every window position looks its k-mer up in its read strand's table
__global__ void __launch_bounds__(256) probeKernel(JoinParams P) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < P.totalRefPos; g += (int64_t)gridDim.x * blockDim.x) {
        const int t = segmentOf(P.taskPosOff, P.nTask, g);
        const int s = P.taskSeq[t];
        const uint8_t* p = P.taskRef[t] + (g - P.taskPosOff[t]);
        const int64_t tab = P.slotOff[s];
        const int64_t slots = P.slotOff[s + 1] - tab;
        int c = 0;
        uint32_t sl = 0;
        if (slots > 0) {
            const uint32_t mask = (uint32_t)slots - 1u;
            const uint8_t* base = P.seq + P.seqOff[s];
            sl = hashBytes(p, P.k) & mask;
            for (;;) {
                const int r = P.rep[tab + sl];
                if (r == -1) break;
                if (sameBytes(base + r, p, P.k)) { c = P.cnt[tab + sl]; break; }
                sl = (sl + 1) & mask;
            }
        }
        P.refCnt[g] = c;
        P.refSlot[g] = (int32_t)sl;
    }
}

// points in the reference's order: window position ascending, read positions of one window position ascending
__global__ void __launch_bounds__(256) emitKernel(JoinParams P) {
    for (int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; g < P.totalRefPos; g += (int64_t)gridDim.x * blockDim.x) {
        const int c = P.refCnt[g];
        if (c == 0) continue;
        const int t = segmentOf(P.taskPosOff, P.nTask, g);
        const int s = P.taskSeq[t];
        const int y = (int)(g - P.taskPosOff[t]);
        const int64_t l0 = P.posOff[s] + P.start[P.slotOff[s] + P.refSlot[g]];
        JoinPoint* o = P.out + P.outOff[t] + P.refOff[g];
        for (int r = 0; r < c; ++r) o[r] = JoinPoint{P.list[l0 + r], y};
    }
}

struct Buf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct KmerJoiner::Impl {
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
    std::mutex mu;
    JoinStats stats;
    std::unordered_map<const char*, std::pair<uint8_t*, size_t> > refs;   // host sequence -> resident copy
    Buf dSeq, dMeta, dSlots, dPos, dRefPos, dOut, hStage, hMeta, hOut;

    void growDev(Buf& b, size_t need) {
        if (need <= b.cap) return;
        if (b.p) JOIN_CUDA(cudaFree(b.p));
        b.p = nullptr; b.cap = 0;
        const size_t cap = need + need / 4 + 4096;
        Engine::noteDeviceAllocation();
        JOIN_CUDA(cudaMalloc(&b.p, cap));
        b.cap = cap;
    }
    void growHost(Buf& b, size_t need) {
        if (need <= b.cap) return;
        if (b.p) JOIN_CUDA(cudaFreeHost(b.p));
        b.p = nullptr; b.cap = 0;
        const size_t cap = need + need / 4 + 4096;
        JOIN_CUDA(cudaMallocHost(&b.p, cap));
        b.cap = cap;
    }
    const uint8_t* residentRef(const char* base, size_t len) {
        auto it = refs.find(base);
        if (it != refs.end() && it->second.second == len) return it->second.first;
        if (it != refs.end()) { cudaFree(it->second.first); refs.erase(it); }
        uint8_t* d = nullptr;
        Engine::noteDeviceAllocation();
        JOIN_CUDA(cudaMalloc((void**)&d, len + 64));
        JOIN_CUDA(cudaMemcpyAsync(d, base, len, cudaMemcpyHostToDevice, stream));
        JOIN_CUDA(cudaStreamSynchronize(stream));   // (the source is pageable caller memory)
        refs[base] = std::make_pair(d, len);
        stats.refUploads++;
        stats.h2dBytes += (int64_t)len;
        return d;
    }
    int gridFor(int64_t n) const {
        const int64_t blocks = (n + 255) / 256;
        return (int)std::max<int64_t>(1, std::min<int64_t>(blocks, (int64_t)numSMs * 8));
    }
};

KmerJoiner::KmerJoiner(int device) : impl_(new Impl) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        delete impl_;
        throw std::runtime_error("unicycler_b200 k-mer join: no CUDA device (there is no CPU fallback for the batch path)");
    }
    if (device < 0) cudaGetDevice(&device);
    impl_->device = device;
    JOIN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    JOIN_CUDA(cudaGetDeviceProperties(&prop, device));
    impl_->numSMs = prop.multiProcessorCount;
    JOIN_CUDA(cudaStreamCreateWithFlags(&impl_->stream, cudaStreamNonBlocking));
    JOIN_CUDA(cudaEventCreate(&impl_->ev0));
    JOIN_CUDA(cudaEventCreate(&impl_->ev1));
    JOIN_CUDA(cudaEventCreate(&impl_->ev2));
    JOIN_CUDA(cudaEventCreate(&impl_->ev3));
}

KmerJoiner::~KmerJoiner() {
    Impl& I = *impl_;
    cudaSetDevice(I.device);
    for (auto& r : I.refs) cudaFree(r.second.first);
    for (Buf* b : {&I.dSeq, &I.dMeta, &I.dSlots, &I.dPos, &I.dRefPos, &I.dOut})
        if (b->p) cudaFree(b->p);
    for (Buf* b : {&I.hStage, &I.hMeta, &I.hOut})
        if (b->p) cudaFreeHost(b->p);
    if (I.stream) cudaStreamDestroy(I.stream);
    for (cudaEvent_t ev : {I.ev0, I.ev1, I.ev2, I.ev3})
        if (ev) cudaEventDestroy(ev);
    delete impl_;
}

void KmerJoiner::forgetReferences() {
    Impl& I = *impl_;
    std::lock_guard<std::mutex> lock(I.mu);
    cudaSetDevice(I.device);
    for (auto& r : I.refs) cudaFree(r.second.first);
    I.refs.clear();
}

JoinStats KmerJoiner::lastStats() const {
    std::lock_guard<std::mutex> lock(impl_->mu);
    return impl_->stats;
}

static size_t roundUp(size_t x, size_t a) { return (x + a - 1) / a * a; }

void KmerJoiner::run(const std::vector<JoinSeq>& seqs, const std::vector<JoinTask>& tasks, int k,
                     std::vector<std::vector<JoinPoint> >& out) {
    Impl& I = *impl_;
    std::lock_guard<std::mutex> lock(I.mu);
    const int nSeq = (int)seqs.size(), nTask = (int)tasks.size();
    out.assign((size_t)nTask, std::vector<JoinPoint>());
    I.stats.kernelMs = 0.0; I.stats.launches = 0; I.stats.h2dBytes = 0; I.stats.d2hBytes = 0; I.stats.points = 0;
    if (nTask == 0 || nSeq == 0 || k <= 0) return;
    JOIN_CUDA(cudaSetDevice(I.device));

    // ---- host metadata: offsets of strands, tables, window positions
    // layout of the metadata block (int64 unless noted):
    //   seqOff[nSeq+1] posOff[nSeq+1] slotOff[nSeq+1] taskPosOff[nTask+1] outOff[nTask] taskRef[nTask] (pointers)
    //   taskSeq[nTask] (int32)
    const size_t metaI64 = 3 * (size_t)(nSeq + 1) + (size_t)(nTask + 1) + 2 * (size_t)nTask;
    const size_t metaBytes = metaI64 * 8 + roundUp((size_t)nTask * 4, 8);
    I.growHost(I.hMeta, metaBytes);
    I.growDev(I.dMeta, metaBytes);
    int64_t* hSeqOff = (int64_t*)I.hMeta.p;
    int64_t* hPosOff = hSeqOff + (nSeq + 1);
    int64_t* hSlotOff = hPosOff + (nSeq + 1);
    int64_t* hTaskPosOff = hSlotOff + (nSeq + 1);
    int64_t* hOutOff = hTaskPosOff + (nTask + 1);
    const uint8_t** hTaskRef = (const uint8_t**)(hOutOff + nTask);
    int32_t* hTaskSeq = (int32_t*)(hTaskRef + nTask);
    hSeqOff[0] = hPosOff[0] = hSlotOff[0] = 0;
    for (int s = 0; s < nSeq; ++s) {
        const int64_t kc = std::max(0, seqs[(size_t)s].len - k + 1);
        int64_t slots = 0;
        if (kc > 0) { slots = 16; while (slots < 2 * kc) slots <<= 1; }
        hSeqOff[s + 1] = hSeqOff[s] + seqs[(size_t)s].len;
        hPosOff[s + 1] = hPosOff[s] + kc;
        hSlotOff[s + 1] = hSlotOff[s] + slots;
    }
    hTaskPosOff[0] = 0;
    for (int t = 0; t < nTask; ++t) {
        const JoinTask& T = tasks[(size_t)t];
        if (T.seq < 0 || T.seq >= nSeq || T.refStart < 0 || (size_t)T.refStart + (size_t)T.refLen > T.refBaseLen)
            throw std::runtime_error("unicycler_b200 k-mer join: task outside its sequences");
        hTaskPosOff[t + 1] = hTaskPosOff[t] + std::max(0, T.refLen - k + 1);
        hTaskRef[t] = I.residentRef(T.refBase, T.refBaseLen) + T.refStart;
        hTaskSeq[t] = T.seq;
        hOutOff[t] = 0;
    }
    const int64_t totalBytes = hSeqOff[nSeq], totalPos = hPosOff[nSeq], totalSlots = hSlotOff[nSeq];
    const int64_t totalRefPos = hTaskPosOff[nTask];
    if (totalPos == 0 || totalRefPos == 0) return;

    // ---- read strands: pinned staging -> device
    I.growHost(I.hStage, (size_t)totalBytes + 64);
    I.growDev(I.dSeq, (size_t)totalBytes + 64);
    parallelFor(nSeq, [&](int s) { memcpy((char*)I.hStage.p + hSeqOff[s], seqs[(size_t)s].s, (size_t)seqs[(size_t)s].len); }, 4);
    I.growDev(I.dSlots, (size_t)totalSlots * 16);
    I.growDev(I.dPos, (size_t)totalPos * 12);
    I.growDev(I.dRefPos, (size_t)totalRefPos * 12 + (size_t)nTask * 8);

    JoinParams P;
    P.seq = (const uint8_t*)I.dSeq.p;
    P.seqOff = (const int64_t*)I.dMeta.p;
    P.posOff = P.seqOff + (nSeq + 1);
    P.slotOff = P.posOff + (nSeq + 1);
    P.taskPosOff = P.slotOff + (nSeq + 1);
    P.outOff = P.taskPosOff + (nTask + 1);
    P.taskRef = (const uint8_t* const*)(P.outOff + nTask);
    P.taskSeq = (const int32_t*)(P.taskRef + nTask);
    P.nSeq = nSeq; P.k = k; P.nTask = nTask;
    P.rep = (int32_t*)I.dSlots.p;
    P.cnt = P.rep + totalSlots;
    P.start = P.cnt + totalSlots;
    P.fill = P.start + totalSlots;
    P.slotOf = (int32_t*)I.dPos.p;
    P.tmp = P.slotOf + totalPos;
    P.list = P.tmp + totalPos;
    P.refCnt = (int32_t*)I.dRefPos.p;
    P.refSlot = P.refCnt + totalRefPos;
    P.refOff = P.refSlot + totalRefPos;
    P.taskTotal = (int64_t*)(P.refOff + totalRefPos + (totalRefPos & 1));
    P.out = nullptr;
    P.totalPos = totalPos; P.totalSlots = totalSlots; P.totalRefPos = totalRefPos;

    cudaStream_t st = I.stream;
    JOIN_CUDA(cudaMemcpyAsync(I.dMeta.p, I.hMeta.p, metaBytes, cudaMemcpyHostToDevice, st));
    JOIN_CUDA(cudaMemcpyAsync(I.dSeq.p, I.hStage.p, (size_t)totalBytes, cudaMemcpyHostToDevice, st));
    I.stats.h2dBytes += (int64_t)metaBytes + totalBytes;
    JOIN_CUDA(cudaEventRecord(I.ev0, st));
    JOIN_CUDA(cudaMemsetAsync(P.rep, 0xFF, (size_t)totalSlots * 4, st));
    JOIN_CUDA(cudaMemsetAsync(P.cnt, 0, (size_t)totalSlots * 4, st));
    JOIN_CUDA(cudaMemsetAsync(P.fill, 0, (size_t)totalSlots * 4, st));
    insertKernel<<<I.gridFor(totalPos), 256, 0, st>>>(P);
    segmentScanKernel<<<nSeq, 1024, 0, st>>>(P.cnt, P.start, P.slotOff, nullptr);
    fillKernel<<<I.gridFor(totalPos), 256, 0, st>>>(P);
    rankKernel<<<I.gridFor(totalPos), 256, 0, st>>>(P);
    probeKernel<<<I.gridFor(totalRefPos), 256, 0, st>>>(P);
    segmentScanKernel<<<nTask, 1024, 0, st>>>(P.refCnt, P.refOff, P.taskPosOff, P.taskTotal);
    JOIN_CUDA(cudaEventRecord(I.ev1, st));
    JOIN_CUDA(cudaGetLastError());
    // per-task point counts -> output offsets (the only round trip)
    I.growHost(I.hOut, (size_t)nTask * 8);
    JOIN_CUDA(cudaMemcpyAsync(I.hOut.p, P.taskTotal, (size_t)nTask * 8, cudaMemcpyDeviceToHost, st));
    JOIN_CUDA(cudaStreamSynchronize(st));
    std::vector<int64_t> total((size_t)nTask);
    int64_t nPoints = 0;
    for (int t = 0; t < nTask; ++t) {
        total[(size_t)t] = ((const int64_t*)I.hOut.p)[t];
        hOutOff[t] = nPoints;
        nPoints += total[(size_t)t];
    }
    I.stats.launches += 6;
    I.stats.d2hBytes += (int64_t)nTask * 8;
    I.stats.points = nPoints;
    float ms = 0.f;
    JOIN_CUDA(cudaEventElapsedTime(&ms, I.ev0, I.ev1));
    I.stats.kernelMs += ms;
    if (nPoints == 0) return;
    I.growDev(I.dOut, (size_t)nPoints * sizeof(JoinPoint));
    I.growHost(I.hOut, (size_t)nPoints * sizeof(JoinPoint));
    P.out = (JoinPoint*)I.dOut.p;
    JOIN_CUDA(cudaMemcpyAsync((void*)P.outOff, hOutOff, (size_t)nTask * 8, cudaMemcpyHostToDevice, st));
    JOIN_CUDA(cudaEventRecord(I.ev2, st));
    emitKernel<<<I.gridFor(totalRefPos), 256, 0, st>>>(P);
    JOIN_CUDA(cudaEventRecord(I.ev3, st));
    JOIN_CUDA(cudaGetLastError());
    JOIN_CUDA(cudaMemcpyAsync(I.hOut.p, I.dOut.p, (size_t)nPoints * sizeof(JoinPoint), cudaMemcpyDeviceToHost, st));
    JOIN_CUDA(cudaStreamSynchronize(st));
    JOIN_CUDA(cudaEventElapsedTime(&ms, I.ev2, I.ev3));
    I.stats.kernelMs += ms;
    I.stats.launches += 1;
    I.stats.h2dBytes += (int64_t)nTask * 8;
    I.stats.d2hBytes += nPoints * (int64_t)sizeof(JoinPoint);
    const JoinPoint* hp = (const JoinPoint*)I.hOut.p;
    parallelFor(nTask, [&](int t) {
        out[(size_t)t].assign(hp + hOutOff[t], hp + hOutOff[t] + total[(size_t)t]);
    }, 4);
}

}  // namespace ub200
