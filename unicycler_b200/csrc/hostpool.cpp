// See hostpool.hpp.
#include "hostpool.hpp"

#include <algorithm>
#include <cstdlib>

namespace ub200 {

int hostThreads() {
    static const int n = [] {
        int v = 0;
        if (const char* e = getenv("UNICYCLER_B200_HOST_THREADS")) v = atoi(e);
        if (v < 1) {
            v = (int)std::thread::hardware_concurrency();
            // one process per GPU (torchrun / mpirun): share the box's cores between the local ranks
            int local = 0;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) local = atoi(e);
            if (local > 1) v = (v + local - 1) / local;
        }
        return std::max(1, v);
    }();
    return n;
}

HostPool& HostPool::instance() {
    static HostPool* pool = new HostPool();   // intentionally leaked: worker threads must outlive static destructors
    return *pool;
}

HostPool::HostPool() {
    const int n = hostThreads() - 1;   // the calling thread is the n-th worker of its own loop
    for (int t = 0; t < n; ++t) threads_.emplace_back([this] { workerMain(); });
}

HostPool::~HostPool() {
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
    }
    cv_.notify_all();
    for (auto& th : threads_) th.join();
}

void HostPool::drain(Loop& L, bool helper) {
    for (;;) {
        // a helper gives way to loops that opened after this one (it comes back when they are done)
        if (helper && newestSeq_.load(std::memory_order_relaxed) > L.seq && newerLoopWantsHelp(L.seq)) break;
        const int b = L.next.fetch_add(L.grain);
        if (b >= L.n) break;
        const int e = std::min(L.n, b + L.grain);
        try {
            for (int i = b; i < e; ++i) (*L.f)(i);
        } catch (...) {
            std::lock_guard<std::mutex> lk(L.errMu);
            if (!L.err) L.err = std::current_exception();
            L.next.store(L.n);
        }
    }
}

bool HostPool::newerLoopWantsHelp(unsigned long long seq) {
    std::lock_guard<std::mutex> lk(mu_);
    for (size_t q = open_.size(); q > 0; --q) {
        const auto& o = open_[q - 1];
        if (o->seq <= seq) break;   // (open_ is in opening order)
        if (o->next.load(std::memory_order_relaxed) < o->n && o->helpersLeft.load(std::memory_order_relaxed) > 0) return true;
    }
    return false;
}

void HostPool::workerMain() {
    for (;;) {
        std::shared_ptr<Loop> L;
        {
            std::unique_lock<std::mutex> lk(mu_);
            for (;;) {
                if (stop_) return;
                for (size_t q = open_.size(); q > 0; --q) {   // newest first
                    auto& o = open_[q - 1];
                    if (o->next.load(std::memory_order_relaxed) >= o->n) continue;
                    if (o->helpersLeft.fetch_sub(1) <= 0) { o->helpersLeft.fetch_add(1); continue; }
                    L = o;
                    break;
                }
                if (L) break;
                cv_.wait(lk);
            }
            L->active.fetch_add(1);
        }
        drain(*L, true);
        L->helpersLeft.fetch_add(1);
        if (L->active.fetch_sub(1) == 1) {
            std::lock_guard<std::mutex> lk(mu_);   // pairs with the owner's wait below
            cv_.notify_all();
        }
    }
}

void HostPool::run(int n, const std::function<void(int)>& f, int grain, int maxHelpers) {
    std::shared_ptr<Loop> L = std::make_shared<Loop>();
    L->f = &f; L->n = n; L->grain = std::max(1, grain);
    const int pool = (int)threads_.size();
    L->helpersLeft.store(maxHelpers < 0 ? pool : std::min(pool, maxHelpers));
    L->active.store(1);
    {
        std::lock_guard<std::mutex> lk(mu_);
        L->seq = ++seqCounter_;
        open_.push_back(L);
        newestSeq_.store(L->seq, std::memory_order_relaxed);
    }
    cv_.notify_all();
    drain(*L, false);
    {
        std::unique_lock<std::mutex> lk(mu_);
        open_.erase(std::find(open_.begin(), open_.end(), L));
        L->active.fetch_sub(1);
        // helpers that are still inside the body finish their indices before the caller's captures go away
        while (L->active.load() > 0) cv_.wait(lk);
    }
    if (L->err) std::rethrow_exception(L->err);
}

}  // namespace ub200
