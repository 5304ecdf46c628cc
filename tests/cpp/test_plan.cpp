// CPU unit test: the product's chain planner (host_align.cpp) against the sequence of sub-DPs the
// oracle's literal restatement performs, on seed chains read from stdin:
//   lines "JOB lenH lenV band nSeeds" followed by nSeeds lines "bH bV eH eV lo up".
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../oracle/dp_oracle.hpp"
#include "../../unicycler_b200/csrc/host_align.hpp"

struct G { int kind; long nH, nV; int banded; long lo, up, hNext, vNext; };
static std::vector<G> g_grids;
static void hook(int kind, long nH, long nV, int banded, long lo, long up, long hNext, long vNext) {
    g_grids.push_back(G{kind, nH, nV, banded, lo, up, hNext, vNext});
}

int main() {
    orc::g_gridHook = hook;
    long jobs = 0, bad = 0, grids = 0;
    long lenH, lenV, band, nSeeds;
    char tag[16];
    std::mt19937 rng(7);
    while (scanf("%15s %ld %ld %ld %ld", tag, &lenH, &lenV, &band, &nSeeds) == 5) {
        std::vector<orc::Seed> oseeds((size_t)nSeeds);
        std::vector<ub200::ChainSeed> pseeds((size_t)nSeeds);
        for (long i = 0; i < nSeeds; ++i) {
            long a, b, c, d, e, f;
            if (scanf("%ld %ld %ld %ld %ld %ld", &a, &b, &c, &d, &e, &f) != 6) return 2;
            oseeds[(size_t)i] = orc::Seed{a, b, c, d, e, f};
            pseeds[(size_t)i] = ub200::ChainSeed{a, b, c, d, e, f};
        }
        // random sequences: the geometry does not depend on the sequence content
        std::vector<uint8_t> H((size_t)lenH), V((size_t)lenV);
        for (auto& x : H) x = rng() % 4;
        for (auto& x : V) x = rng() % 4;
        // make the seeds real matches so that the oracle does not bail out with a bad score
        for (const auto& s : oseeds)
            for (long k = 0; k < s.endH - s.beginH && s.beginV + k < lenV; ++k) V[(size_t)(s.beginV + k)] = H[(size_t)(s.beginH + k)];
        g_grids.clear();
        orc::Trace tr; bool empty; int score;
        orc::Score sc{3, -6, -2, -5};
        orc::FreeEnds fe{true, true, true, true};
        bool ok = orc::bandedChainAlignmentTrace(H, V, oseeds, sc, fe, (unsigned)band, tr, empty, score);
        std::vector<ub200::GridDesc> plan;
        bool pok = ub200::planChain(pseeds, lenH, lenV, band, plan);
        ++jobs;
        bool mismatch = !pok;
        // if the oracle threw part-way (ok == false) only the prefix it reached is comparable
        size_t n = ok ? plan.size() : g_grids.size();
        if (ok && plan.size() != g_grids.size()) mismatch = true;
        for (size_t k = 0; k < n && k < plan.size() && k < g_grids.size(); ++k) {
            const ub200::GridDesc& p = plan[k];
            const G& o = g_grids[k];
            bool fullBand = o.banded && o.lo <= -o.nV && o.up >= o.nH;  // planner runs these unbanded
            if (p.kind != o.kind || p.nH != o.nH || p.nV != o.nV || p.hNext != o.hNext || p.vNext != o.vNext) mismatch = true;
            if (!fullBand && (p.banded != o.banded || (o.banded && (p.lo != o.lo || p.up != o.up)))) mismatch = true;
            if (fullBand && p.banded) mismatch = true;
            ++grids;
        }
        if (mismatch) {
            ++bad;
            if (bad <= 5) fprintf(stderr, "MISMATCH job %ld (plan %zu grids, oracle %zu grids, ok=%d pok=%d)\n", jobs - 1, plan.size(), g_grids.size(), ok, pok);
        }
    }
    printf("jobs %ld grids %ld bad %ld\n", jobs, grids, bad);
    return bad ? 1 : 0;
}
