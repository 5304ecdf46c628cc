// CPU unit test: PointSet (unicycler_b200/csrc/pointset.hpp) hands out its elements in exactly the order
// std::unordered_set<Point, PointHash> does, after every kind of operation the line tracer performs on it
// (range construction, range insertion with duplicates, single insertions, copies), on random and on
// diagonal-run point patterns.
#include <cstdio>
#include <random>
#include <unordered_set>
#include <vector>

#include <cstring>

#include "../../unicycler_b200/csrc/linegeom.hpp"
#include "../../unicycler_b200/csrc/pointset.hpp"

using ub200::seed::Point;
using ub200::seed::PointHash;
using ub200::seed::PointSet;
typedef std::unordered_set<Point, PointHash> RefSet;

static long g_checks = 0, g_bad = 0;

static void compare(const RefSet& r, const PointSet& s, const char* what, int it) {
    ++g_checks;
    bool ok = r.size() == s.size();
    auto a = r.begin();
    auto b = s.begin();
    for (; ok && a != r.end() && b != s.end(); ++a, ++b) ok = (*a == *b);
    ok = ok && a == r.end() && !(b != s.end());
    if (!ok) {
        ++g_bad;
        if (g_bad <= 5) fprintf(stderr, "ORDER MISMATCH after %s (iteration %d): sizes %zu / %zu\n", what, it, r.size(), s.size());
    }
}

int main(int argc, char** argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 300;
    std::mt19937 rng(20240607);
    for (int it = 0; it < iters; ++it) {
        ub200::seed::Arena::mine().rewind();
        const int span = 50 + (int)(rng() % 40000);
        auto point = [&]() {
            if (it % 3 == 0) return Point((int)(rng() % span), (int)(rng() % span));           // scattered
            const int x = (int)(rng() % span), d = (int)(rng() % 41) - 20;                     // near a diagonal
            return Point(x, std::max(0, x + 137 + d));
        };
        auto batch = [&](size_t n) {
            std::vector<Point> v(n);
            for (auto& p : v) p = point();
            if (n > 4) for (size_t k = 0; k < n / 5; ++k) v[rng() % n] = v[rng() % n];         // duplicates inside the batch
            return v;
        };
        // range construction (lineTracing: the points around the start point)
        std::vector<Point> first = batch(rng() % 400);
        RefSet r(first.begin(), first.end());
        PointSet s(first.begin(), first.end());
        compare(r, s, "range construction", it);
        // rounds of range insertions (mutateLineToBestFitPoints) and single insertions (addPointsNearLine, used points)
        const int rounds = 1 + (int)(rng() % 12);
        for (int q = 0; q < rounds; ++q) {
            std::vector<Point> more = batch(rng() % (q % 4 == 3 ? 6000 : 500));
            if (!first.empty()) for (size_t k = 0; k < more.size() / 4; ++k) more[rng() % more.size()] = first[rng() % first.size()];   // already present
            if (q & 1) {
                r.insert(more.begin(), more.end());
                s.insert(more.begin(), more.end());
            } else {
                for (const Point& p : more) {
                    const bool a = r.insert(p).second, b = s.insert(p);
                    if (a != b) { ++g_bad; if (g_bad <= 5) fprintf(stderr, "insert() result differs\n"); }
                }
            }
            compare(r, s, "insertions", it);
            for (int k = 0; k < 50; ++k) {
                const Point p = (k & 1) && !more.empty() ? more[rng() % more.size()] : point();
                if ((r.find(p) != r.end()) != s.contains(p)) { ++g_bad; if (g_bad <= 5) fprintf(stderr, "contains() differs\n"); }
            }
            first.insert(first.end(), more.begin(), more.end());
        }
        // copies keep the order, and keep growing like the original
        RefSet r2(r);
        PointSet s2(s);
        compare(r2, s2, "copy", it);
        std::vector<Point> more = batch(300);
        r2.insert(more.begin(), more.end());
        s2.insert(more.begin(), more.end());
        compare(r2, s2, "insertions into a copy", it);
        compare(r, s, "original after the copy grew", it);
    }
    // distancesToLineSegment (two points per step) against distanceToLineSegment, bit for bit
    long distChecks = 0;
    for (int it = 0; it < 4000; ++it) {
        const int span = (it % 5 == 0) ? 30 : 60000;
        Point l1((int)(rng() % span), (int)(rng() % span)), l2((int)(rng() % span), (int)(rng() % span));
        if (it % 7 == 0) l2 = l1;                                   // degenerate segment
        if (it % 11 == 0) l2 = Point(l1.x + 500, l1.y + 500);         // the tracer's step
        if (it % 13 == 0) { l1 = Point(l1.x - 40000, l1.y - 40000); }   // segments may leave the rectangle (negative ends)
        std::vector<Point> pts(rng() % 70);
        for (auto& p : pts) p = (rng() % 4 == 0) ? (rng() % 2 ? l1 : l2) : Point((int)(rng() % span), (int)(rng() % span));
        std::vector<double> got(pts.size() + 1, -1.0);
        ub200::seed::distancesToLineSegment(pts.data(), pts.size(), l1, l2, got.data());
        for (size_t i = 0; i < pts.size(); ++i) {
            const double want = ub200::seed::distanceToLineSegment(pts[i], l1, l2);
            ++distChecks;
            if (memcmp(&want, &got[i], sizeof(double)) != 0) { ++g_bad; if (g_bad <= 5) fprintf(stderr, "DISTANCE differs: %.17g vs %.17g\n", want, got[i]); }
        }
    }
    printf("checks %ld distances %ld bad %ld\n", g_checks, distChecks, g_bad);
    return g_bad ? 1 : 0;
}
