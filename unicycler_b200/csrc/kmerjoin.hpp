// Common k-mer join on the device (first half: row a4).
//
// The reference finds, for one (read strand, reference window), every pair (read position, window position) whose
// k-mers are equal as STRINGS (src/semi_global_align.cpp:197-207 over KmerPositions::addPositions,
// src/kmers.cpp:51-65), window positions ascending and, for one window position, read positions ascending.  That
// list ("common k-mer points") feeds the line tracer.  Here the whole batch is joined in one pass on the GPU:
//
//   build   one open-addressing table per read strand, keyed by the k literal bytes of a k-mer (any character:
//           N, lower case, ... compare as bytes, like the reference's std::string keys); per distinct k-mer the
//           ascending list of its read positions (count -> exclusive scan -> fill -> per-list sort)
//   probe   every window position of every task looks its k-mer up (count), a per-task exclusive scan gives
//           each position its place in the task's output, and a second pass writes the points in the
//           reference's order
//
// The reference sequences stay resident in HBM between calls (uploaded once per sequence, dropped by
// deleteRefSeqs); read strands are uploaded per call.  Integer / byte work only: results are identical to the
// host join (seeding.cpp) by construction and are checked against it in tests/.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace ub200 {

struct JoinSeq {             // one read strand
    const char* s;
    int len;
};
struct JoinTask {            // one reference window of one read strand
    int seq;                 // index into the call's JoinSeq list
    const char* refBase;     // the whole reference sequence (host memory; resident copy keyed by this pointer)
    size_t refBaseLen;
    int refStart, refLen;    // the window [refStart, refStart + refLen) of it
};
struct JoinPoint {           // (read position, window position)
    int32_t x, y;
};
struct JoinStats {
    double kernelMs = 0.0;   // CUDA-event time of the join kernels of the last call
    int64_t launches = 0, h2dBytes = 0, d2hBytes = 0, points = 0, refUploads = 0;
};

class KmerJoiner {
public:
    explicit KmerJoiner(int device);   // device < 0: current device; throws when CUDA is unusable
    ~KmerJoiner();
    // out[t] = the common k-mer points of task t in the reference's order.  Thread-safe (serialised).
    void run(const std::vector<JoinSeq>& seqs, const std::vector<JoinTask>& tasks, int k,
             std::vector<std::vector<JoinPoint> >& out);
    void forgetReferences();           // drops the resident reference copies (deleteRefSeqs)
    JoinStats lastStats() const;

private:
    struct Impl;
    Impl* impl_;
};

}  // namespace ub200
