// Host pool (unicycler_b200/csrc/hostpool.cpp): every index exactly once, nested loops, concurrent callers, exceptions.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <stdexcept>
#include <thread>
#include <vector>

#include "../../unicycler_b200/csrc/hostpool.hpp"

using namespace ub200;

int main() {
    int bad = 0;
    {   // every index exactly once, various grains
        for (int grain : {1, 7, 256}) {
            std::vector<std::atomic<int> > hit(10000);
            parallelFor(10000, [&](int i) { hit[(size_t)i].fetch_add(1); }, grain);
            for (auto& h : hit) if (h.load() != 1) ++bad;
        }
    }
    {   // nested loops: the inner loop runs inside tasks of the outer one
        std::atomic<long> sum(0);
        parallelFor(64, [&](int i) { parallelFor(100, [&](int j) { sum.fetch_add((long)i * 100 + j); }); });
        if (sum.load() != 6399L * 6400 / 2) ++bad;
    }
    {   // concurrent callers (the reference calls the library from a pool of Python threads)
        std::atomic<long> total(0);
        std::vector<std::thread> callers;
        for (int t = 0; t < 8; ++t)
            callers.emplace_back([&] { for (int rep = 0; rep < 50; ++rep) parallelFor(200, [&](int) { total.fetch_add(1); }); });
        for (auto& c : callers) c.join();
        if (total.load() != 8L * 50 * 200) ++bad;
    }
    {   // an exception in a task reaches the caller, the pool stays usable
        bool caught = false;
        try { parallelFor(1000, [&](int i) { if (i == 500) throw std::runtime_error("x"); }); } catch (const std::runtime_error&) { caught = true; }
        if (!caught) ++bad;
        std::atomic<int> n(0);
        parallelFor(100, [&](int) { n.fetch_add(1); });
        if (n.load() != 100) ++bad;
    }
    if (hostThreads() >= 4) {
        // a long loop (the pipeline's line tracing: milliseconds per item) lends its workers to a short loop that a second
        // thread opens later (the driving thread's parsing / formatting): the short loop must not be run by its caller alone
        std::atomic<int> longDone(0), shortThreads(0);
        std::atomic<bool> shortOpen(false);
        std::vector<std::atomic<int> > seen(64);
        std::thread longCaller([&] {
            parallelFor(400, [&](int) { std::this_thread::sleep_for(std::chrono::milliseconds(2)); longDone.fetch_add(1); });
        });
        std::this_thread::sleep_for(std::chrono::milliseconds(20));   // every worker is inside the long loop by now
        shortOpen = true;
        std::atomic<int> shortDone(0);
        thread_local int tid = -1;
        std::atomic<int> nextTid(0);
        parallelFor(64, [&](int) {
            if (tid < 0) tid = nextTid.fetch_add(1);
            seen[(size_t)tid % 64].fetch_add(1);
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            shortDone.fetch_add(1);
        });
        const int stillLong = 400 - longDone.load();
        for (auto& x : seen) if (x.load() > 0) shortThreads.fetch_add(1);
        longCaller.join();
        if (shortDone.load() != 64 || longDone.load() != 400) ++bad;
        if (stillLong <= 0) ++bad;               // the short loop ran while the long one was still open
        if (shortThreads.load() < 2) { ++bad; fprintf(stderr, "the short loop got no helper from the long loop's workers\n"); }
    }
    printf("hostpool threads %d bad %d\n", hostThreads(), bad);
    return bad ? 1 : 0;
}
