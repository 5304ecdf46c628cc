// Host pool (unicycler_b200/csrc/hostpool.cpp): every index exactly once, nested loops, concurrent callers, exceptions.
#include <atomic>
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
    printf("hostpool threads %d bad %d\n", hostThreads(), bad);
    return bad ? 1 : 0;
}
