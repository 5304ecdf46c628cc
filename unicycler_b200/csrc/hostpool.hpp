// Process-wide host thread pool shared by every host stage of the library (seeding, staging, result parsing,
// formatting).  The reference gets its host parallelism from Python threads calling the library concurrently
// (unicycler_align.py:203-225); here the library owns ONE lazily created pool, sized by
// UNICYCLER_B200_HOST_THREADS or, by default, the machine's cores divided by LOCAL_WORLD_SIZE (one process per GPU
// must not oversubscribe the box N-fold).  parallelFor() may be called concurrently from many threads and may be
// nested: the caller always takes part in its own loop, pool workers help with whatever loops are open — the NEWEST
// loop first, and a worker leaves an older loop between two of its items when a newer one has opened: the long loop of
// a pipeline (line tracing of the next chunk, milliseconds per item) lends its workers to the short loops of the
// thread that drives the pipeline (result parsing, formatting, staging) instead of leaving that thread to run them alone.
#pragma once
#include <atomic>
#include <condition_variable>
#include <exception>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace ub200 {

// Host threads this process may use (>= 1).
int hostThreads();

class HostPool {
public:
    static HostPool& instance();
    // f(i) for every i in [0, n); grain = indices claimed at a time.  maxHelpers < 0: the whole pool.
    void run(int n, const std::function<void(int)>& f, int grain = 1, int maxHelpers = -1);
    int workers() const { return (int)threads_.size(); }
    ~HostPool();

private:
    struct Loop {
        const std::function<void(int)>* f;
        int n, grain;
        unsigned long long seq = 0;   // opening order
        std::atomic<int> next{0};
        std::atomic<int> active{0};   // threads currently inside the loop body
        std::atomic<int> helpersLeft{0};
        std::exception_ptr err;
        std::mutex errMu;
    };
    HostPool();
    void workerMain();
    void drain(Loop& L, bool helper);
    bool newerLoopWantsHelp(unsigned long long seq);
    unsigned long long seqCounter_ = 0;   // under mu_
    std::atomic<unsigned long long> newestSeq_{0};
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_;
    std::vector<std::shared_ptr<Loop> > open_;
    bool stop_ = false;
};

template <typename F>
inline void parallelFor(int n, F f, int grain = 1, int maxHelpers = -1) {
    if (n <= 0) return;
    if (n <= grain || hostThreads() == 1) { for (int i = 0; i < n; ++i) f(i); return; }
    const std::function<void(int)> fn(f);
    HostPool::instance().run(n, fn, grain, maxHelpers);
}

}  // namespace ub200
