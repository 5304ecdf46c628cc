// Host seeding stage (see seeding.hpp).  Restates unicycler/src/semi_global_align.cpp:197-291,350-605,739-803,
// the L1 kd-tree radius search of nanoflann 1.2.3 (include/nanoflann.hpp:1043-1256) and SeqAn's seed merge
// and sparse global chaining (seqan/seeds/seeds_seed_set_unordered.h:248-361, seeds_combination.h:102-197,
// seeds_global_chaining.h:102-280).  The container types, hash and iteration orders are the reference's
// on purpose: several decisions sum doubles in container order.
#include "seeding.hpp"
#include "hostpool.hpp"
#include "pointset.hpp"
#include "linegeom.hpp"

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

#include <algorithm>
#include <cmath>
#include <limits>
#include <map>
#include <numeric>
#include <set>
#include <unordered_map>
#include <unordered_set>

namespace ub200 {

// developer timers of the host seeding stages (seconds, summed over threads)
std::atomic<long long> g_seedProf[6];
// ... and of the line tracer's parts: 0 radius searches, 1 near-point set, 2 segment scoring, 3 collection, 4 whole trace loop,
// 5 point-set score, 6 used-point bookkeeping, 7 the range's kd-tree
std::atomic<long long> g_lt[10];
// UNICYCLER_B200_MARGINS=1: per range, the smallest relative distance between the two sides of every floating-point
// comparison that decides a result (0 density winner vs runner-up, 1 segment mutation, 2 good/bad line, 3 best line) —
// the evidence behind DESIGN.md 4b (which decisions would survive another summation order).
static const bool g_margins = getenv("UNICYCLER_B200_MARGINS") != nullptr;
static thread_local double t_minMargin[4] = {1e300, 1e300, 1e300, 1e300};
static inline void noteMargin(int k, double a, double b) {
    const double m = std::fabs(a - b) / (std::fabs(a) + std::fabs(b) + 1e-300);
    if (m < t_minMargin[k]) t_minMargin[k] = m;
}
static inline long long nowNs() {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}


// include/settings.h
static const int LINE_TRACING_START_POINT_SEARCH_RADIUS = 100;
static const double TRACE_LINE_COLLECTION_DISTANCE = 20.0;
static const int TRACE_LINE_STEP_DISTANCE = 500;
static const int TRACE_LINE_MUTATION_SIZE = 5;
static const double MAX_POINTS_SCORE = 10000.0;
static const double MAX_SLOPE_SCORE = 1.0;
static const double MIN_ACCEPTABLE_LINE_SEGMENT_SLOPE = 0.5;
static const long long MAX_BANDED_ALIGNMENT_GAP_AREA = 100000000LL;
static const double SCORE_DISTANCE_FROM_DIAGONAL = 5.0;

SensitivityParams sensitivityParams(int level) {
    SensitivityParams p{10, 25, 2, 4};
    if (level == 1) p = SensitivityParams{10, 50, 2, 8};
    else if (level == 2) p = SensitivityParams{9, 75, 3, 12};
    else if (level == 3) p = SensitivityParams{8, 100, 4, 16};
    return p;
}

namespace {
inline int baseCode(char c) {  // upper-case ACGT -> 0..3, anything else -> -1
    switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return -1; }
}
inline uint32_t kmerHash(uint32_t code) { return (code * 2654435761u) >> 7; }

// Calls f(i, clean, code) for every k-mer start i of seq: clean k-mers carry their 2-bit code.
template <typename F>
inline void forEachKmer(const std::string& seq, int k, F f) {
    const int n = (int)seq.size() - k + 1;
    if (n <= 0) return;
    const uint32_t codeMask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    uint32_t code = 0;
    int run = 0;  // number of consecutive clean characters ending at the current one
    for (int p = 0; p < (int)seq.size(); ++p) {
        const int b = baseCode(seq[(size_t)p]);
        if (b < 0) { run = 0; code = 0; }
        else { ++run; code = ((code << 2) | (uint32_t)b) & codeMask; }
        const int i = p - k + 1;
        if (i >= 0) f(i, run >= k, code);
    }
}
}  // namespace

void buildKmerPositions(const std::string& sequence, int kSize, KmerPosMap& out) {
    out = KmerPosMap();
    out.k = kSize;
    const int kCount = (int)sequence.size() - kSize + 1;
    if (kCount <= 0) return;
    if (kSize > 16) {  // codes would not fit 32 bits: literal keys only (never the case for sensitivity 0..3)
        for (int i = 0; i < kCount; ++i) out.other[sequence.substr((size_t)i, (size_t)kSize)].push_back(i);
        return;
    }
    uint32_t slots = 16;
    while (slots < (uint32_t)kCount * 2u) slots <<= 1;
    out.mask = slots - 1;
    out.head.assign(slots, -1);
    out.key.assign(slots, 0);
    out.next.assign((size_t)kCount, -1);
    // lists are built back to front so that they read in ascending position order
    std::vector<uint32_t> codes((size_t)kCount);
    std::vector<uint8_t> clean((size_t)kCount);
    forEachKmer(sequence, kSize, [&](int i, bool ok, uint32_t code) { codes[(size_t)i] = code; clean[(size_t)i] = ok; });
    for (int i = kCount - 1; i >= 0; --i) {
        if (!clean[(size_t)i]) continue;
        const uint32_t code = codes[(size_t)i];
        uint32_t sl = kmerHash(code) & out.mask;
        while (out.head[sl] != -1 && out.key[sl] != code) sl = (sl + 1) & out.mask;
        out.key[sl] = code;
        out.next[(size_t)i] = out.head[sl];
        out.head[sl] = i;
    }
    for (int i = 0; i < kCount; ++i)
        if (!clean[(size_t)i]) out.other[sequence.substr((size_t)i, (size_t)kSize)].push_back(i);
}

std::string reverseComplement(const std::string& s) {
    std::string rc;
    rc.reserve(s.size());
    for (size_t k = s.size(); k > 0; --k) {
        switch (s[k - 1]) {
        case 'A': rc.push_back('T'); break;
        case 'T': rc.push_back('A'); break;
        case 'G': rc.push_back('C'); break;
        case 'C': rc.push_back('G'); break;
        case 'R': rc.push_back('Y'); break;
        case 'Y': rc.push_back('R'); break;
        case 'S': rc.push_back('S'); break;
        case 'W': rc.push_back('W'); break;
        case 'K': rc.push_back('M'); break;
        case 'M': rc.push_back('K'); break;
        case 'B': rc.push_back('V'); break;
        case 'D': rc.push_back('H'); break;
        case 'H': rc.push_back('D'); break;
        case 'V': rc.push_back('B'); break;
        case 'N': rc.push_back('N'); break;
        case '.': rc.push_back('.'); break;
        case '-': rc.push_back('-'); break;
        case '?': rc.push_back('?'); break;
        case '*': rc.push_back('*'); break;
        default: break;  // any other character is dropped, like the reference
        }
    }
    return rc;
}

namespace {

using seed::Point;
using seed::PointHash;
using seed::PointSet;      // the reference's std::unordered_set<Point, PointHash>, iteration order included (pointset.hpp)
using seed::PointVector;

// ------------------------------------------------------------------ kd-tree (nanoflann algorithm, int L1 metric)
// Same tree, same leaf order and same search order as nanoflann's KDTreeSingleIndexAdaptor<L1_Adaptor<int,...>, ..., 2>
// (include/nanoflann.hpp:1043-1256) built with leaf size 10 (semi_global_align.cpp:217): the ORDER in which a radius
// search reports its matches is observable (doubles are summed in that order), so it is kept.  What is exploited:
// with integer element / distance types the EPS of middleSplit is static_cast<int>(0.00001) == 0, so the test
// "span > (1 - EPS) * max_span" never holds and EVERY split is on x (nanoflann.hpp:1113-1131).  The build therefore
// only ever looks at x: it permutes one contiguous (x, index) array instead of chasing indices into the point array,
// keeps x-intervals only, and the points are then copied in leaf order so that leaf scans are contiguous.
class KdTree {
public:
    explicit KdTree(const PointVector& pts) : pts_(pts), root_(-1) {}

    void build() {
        const size_t n = pts_.size();
        nodes_.clear();
        root_ = -1;
        if (n == 0) return;
        ent_.resize(n);
        for (int d = 0; d < 2; ++d) rootBox_[d].low = rootBox_[d].high = (d == 0 ? pts_[0].x : pts_[0].y);
        for (size_t i = 0; i < n; ++i) {   // computeBoundingBox (the boxes merged back up the tree are the same extremes)
            ent_[i] = Ent{pts_[i].x, (uint32_t)i};
            rootBox_[0].low = std::min(rootBox_[0].low, pts_[i].x); rootBox_[0].high = std::max(rootBox_[0].high, pts_[i].x);
            rootBox_[1].low = std::min(rootBox_[1].low, pts_[i].y); rootBox_[1].high = std::max(rootBox_[1].high, pts_[i].y);
        }
        nodes_.reserve(n / 4 + 8);
        Interval box = rootBox_[0];
        root_ = divide(0, n, box);
        leafPts_.resize(n);
        for (size_t i = 0; i < n; ++i) leafPts_[i] = pts_[ent_[i].idx];
    }

    // radiusSearch with SearchParams() defaults: eps = 0, sorted by distance (unstable std::sort on the matches in
    // search order: the permutation depends on the comparison results only); the matched points in that order
    void radiusSearch(Point q, int radius, PointVector& out) const {
        out.clear();
        if (pts_.empty() || root_ < 0) return;
        static thread_local std::vector<Match> matches;
        matches.clear();
        int dists[2] = {0, 0};
        int distsq = 0;
        const int qc[2] = {q.x, q.y};
        for (int d = 0; d < 2; ++d) {
            if (qc[d] < rootBox_[d].low) { dists[d] = std::abs(qc[d] - rootBox_[d].low); distsq += dists[d]; }
            if (qc[d] > rootBox_[d].high) { dists[d] = std::abs(qc[d] - rootBox_[d].high); distsq += dists[d]; }
        }
        search(matches, q, radius, root_, distsq, dists[0]);
        std::sort(matches.begin(), matches.end(), [](const Match& a, const Match& b) { return a.dist < b.dist; });
        out.reserve(matches.size());
        for (const Match& m : matches) out.push_back(m.p);
    }

private:
    struct Interval { int low, high; };
    struct Ent { int x; uint32_t idx; };
    struct Match { int dist; Point p; };
    struct Node {
        int32_t left, right;      // leaf: range of leafPts_; inner: left = -1
        int32_t divlow, divhigh;  // inner: largest x of the left subtree, smallest x of the right subtree
        int32_t child1, child2;
    };
    const PointVector& pts_;
    std::vector<Ent> ent_;
    PointVector leafPts_;
    std::vector<Node> nodes_;
    int root_;
    Interval rootBox_[2];
    static const size_t kLeafMax = 10;  // KDTreeSingleIndexAdaptorParams(10), semi_global_align.cpp:217

    // divideTree (nanoflann.hpp:1043-1094) on x-intervals; box comes in as the parent's clipped interval and goes back
    // as the exact interval of the subtree
    int divide(size_t left, size_t right, Interval& box) {
        const int id = (int)nodes_.size();
        nodes_.push_back(Node());
        if (right - left <= kLeafMax) {
            int mn = ent_[left].x, mx = mn;
            for (size_t k = left + 1; k < right; ++k) { mn = std::min(mn, ent_[k].x); mx = std::max(mx, ent_[k].x); }
            box.low = mn; box.high = mx;
            nodes_[(size_t)id] = Node{(int32_t)left, (int32_t)right, 0, 0, -1, -1};
            return id;
        }
        // middleSplit (nanoflann.hpp:1113-1160) with cutfeat == 0
        Ent* ind = &ent_[left];
        const size_t count = right - left;
        const int splitVal = (box.low + box.high) / 2;
        int mn = ind[0].x, mx = mn;
        for (size_t i = 1; i < count; ++i) { mn = std::min(mn, ind[i].x); mx = std::max(mx, ind[i].x); }
        const int cutval = splitVal < mn ? mn : (splitVal > mx ? mx : splitVal);
        size_t lim1, lim2;
        planeSplit(ind, count, cutval, lim1, lim2);
        size_t index;
        if (lim1 > count / 2) index = lim1;
        else if (lim2 < count / 2) index = lim2;
        else index = count / 2;
        Interval lbox = box, rbox = box;
        lbox.high = cutval;
        const int c1 = divide(left, left + index, lbox);
        rbox.low = cutval;
        const int c2 = divide(left + index, right, rbox);
        nodes_[(size_t)id] = Node{-1, 0, lbox.high, rbox.low, c1, c2};
        box.low = std::min(lbox.low, rbox.low);
        box.high = std::max(lbox.high, rbox.high);
        return id;
    }

    // planeSplit (nanoflann.hpp:1170-1202): the same two exchange passes, on the (x, index) entries
    static void planeSplit(Ent* ind, size_t count, int cutval, size_t& lim1, size_t& lim2) {
        size_t left = 0, right = count - 1;
        for (;;) {
            while (left <= right && ind[left].x < cutval) ++left;
            while (right && left <= right && ind[right].x >= cutval) --right;
            if (left > right || !right) break;
            std::swap(ind[left], ind[right]);
            ++left;
            --right;
        }
        lim1 = left;
        right = count - 1;
        for (;;) {
            while (left <= right && ind[left].x <= cutval) ++left;
            while (right && left <= right && ind[right].x > cutval) --right;
            if (left > right || !right) break;
            std::swap(ind[left], ind[right]);
            ++left;
            --right;
        }
        lim2 = left;
    }

    // searchLevel (nanoflann.hpp:1212-1256); every split is on x, so only the x entry of `dists` ever changes
    void search(std::vector<Match>& out, Point q, int radius, int nodeId, int mindistsq, int distX) const {
        const Node& nd = nodes_[(size_t)nodeId];
        if (nd.left >= 0) {
            for (int32_t i = nd.left; i < nd.right; ++i) {
                const Point& p = leafPts_[(size_t)i];
                const int dist = std::abs(q.x - p.x) + std::abs(q.y - p.y);
                if (dist < radius) out.push_back(Match{dist, p});
            }
            return;
        }
        const int val = q.x;
        const int diff1 = val - nd.divlow, diff2 = val - nd.divhigh;
        int best, other, cutDist;
        if (diff1 + diff2 < 0) { best = nd.child1; other = nd.child2; cutDist = std::abs(val - nd.divhigh); }
        else { best = nd.child2; other = nd.child1; cutDist = std::abs(val - nd.divlow); }
        search(out, q, radius, best, mindistsq, distX);
        mindistsq = mindistsq + cutDist - distX;
        if (mindistsq <= radius) search(out, q, radius, other, mindistsq, cutDist);
    }
};

// ------------------------------------------------------------------ line tracing (semi_global_align.cpp:350-605,739-803)
struct Cloud {
    PointVector pts;
    std::vector<uint32_t> orig;   // index of every point in the range's common points (ascending)
    KdTree tree;
    Cloud() : tree(pts) {}
};

void fillCloud(Cloud& cloud, const PointVector& common, const PointSet& used) {
    cloud.pts.clear();
    cloud.orig.clear();
    for (size_t i = 0; i < common.size(); ++i)
        if (used.empty() || !used.contains(common[i])) { cloud.pts.push_back(common[i]); cloud.orig.push_back((uint32_t)i); }
    cloud.tree.build();
}

// The range's common points ordered by diagonal (x - y), then along the diagonal (x + y): sorted once per range, every
// start cloud (a subset in the same relative order) filters it.
struct DiagOrder {
    std::vector<uint32_t> order;
    std::vector<int32_t> pos;     // scratch: common index -> index in the current cloud, -1 = not in it
    void build(const PointVector& common) {
        struct Key { uint64_t key; uint32_t idx; };
        std::vector<Key> keys(common.size());
        for (size_t i = 0; i < common.size(); ++i) {
            const uint32_t d = (uint32_t)(common[i].x - common[i].y + 0x40000000), sum = (uint32_t)(common[i].x + common[i].y);
            keys[i] = Key{((uint64_t)d << 32) | sum, (uint32_t)i};
        }
        std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.key < b.key; });
        order.resize(common.size());
        for (size_t i = 0; i < common.size(); ++i) order[i] = keys[i].idx;
    }
};

PointVector radiusSearchAroundPoint(Point point, int radius, const Cloud& cloud) {
    PointVector out;
    cloud.tree.radiusSearch(point, radius, out);
    return out;
}

double getSlope(const Point& p1, const Point& p2) {
    int xDiff = p1.x - p2.x, yDiff = p1.y - p2.y;
    double slope = 0.0;
    if (xDiff != 0) slope = double(yDiff) / double(xDiff);
    if (xDiff == 0 && yDiff == 0) slope = 1.0;
    return slope;
}

using seed::distanceToLineSegment;   // linegeom.hpp (one point / two points at a time, bit-identical)

// the near-point set of the current step in its iteration order, and scratch for the distances (per thread)
static thread_local PointVector t_flat;
static thread_local std::vector<double> t_dist;

// pointsNearLine: the step's near-point set copied out once in its iteration order (the order of the sum)
double scoreLineSegment(Point p1, Point p2, const PointVector& pointsNearLine) {
    double slope = getSlope(p1, p2);
    if (slope > 1.0) slope = 1.0 / slope;
    double slopeScore = (MAX_SLOPE_SCORE / (1.0 - MIN_ACCEPTABLE_LINE_SEGMENT_SLOPE)) * (slope - MIN_ACCEPTABLE_LINE_SEGMENT_SLOPE);
    double maxScorePerPoint = MAX_POINTS_SCORE / TRACE_LINE_STEP_DISTANCE;
    double pointDistanceScore = 0.0;
    const size_t n = pointsNearLine.size();
    if (t_dist.size() < n) t_dist.resize(n);
    seed::distancesToLineSegment(pointsNearLine.data(), n, p1, p2, t_dist.data());
    for (size_t i = 0; i < n; ++i) pointDistanceScore += maxScorePerPoint / (t_dist[i] + 1.0);   // (summed in set order)
    return slopeScore + pointDistanceScore;
}

Point shiftUp(Point p, int steps) { p.x -= steps; p.y += steps; return p; }
Point shiftDown(Point p, int steps) { p.x += steps; p.y -= steps; return p; }

// the last radius search of the line tracer on this thread (valid for one range's cloud: seedRange resets it)
struct LastSearch { const Cloud* cloud = nullptr; Point q; PointVector pts; };
static thread_local LastSearch g_lastSearch;

Point mutateLineToBestFitPoints(Point p1, Point p2, const Cloud& cloud, PointSet& pointsNearLine, bool leftRectangle) {
    int radius = int(TRACE_LINE_STEP_DISTANCE * 1.1);
    // p1 is the previous step's end point: when that step kept its unmutated end, the search around it has just been
    // done (as that step's p2).  Same cloud, same query, same radius => same list in the same order.
    LastSearch& last = g_lastSearch;
    const long long tm0 = nowNs();
    PointVector nearP1;
    if (last.cloud == &cloud && last.q == p1) nearP1.swap(last.pts);
    else nearP1 = radiusSearchAroundPoint(p1, radius, cloud);
    PointVector nearP2 = radiusSearchAroundPoint(p2, radius, cloud);
    struct Keep {   // remember the search around the unmutated p2 for the next step
        LastSearch& l; const Cloud* c; Point q; PointVector& v;
        ~Keep() { l.cloud = c; l.q = q; l.pts.swap(v); }
    } keep{last, &cloud, p2, nearP2};
    const long long tm1 = nowNs();
    pointsNearLine.insert(nearP1.begin(), nearP1.end());
    pointsNearLine.insert(nearP2.begin(), nearP2.end());
    const long long tm2 = nowNs();
    g_lt[0] += tm1 - tm0; g_lt[1] += tm2 - tm1;
    struct Sc { long long t; ~Sc() { g_lt[2] += nowNs() - t; } } scTimer{tm2};
    PointVector& flat = t_flat;
    flat.assign(pointsNearLine.begin(), pointsNearLine.end());
    if (leftRectangle) return p2;
    Point p2Up = shiftUp(p2, TRACE_LINE_MUTATION_SIZE), p2Down = shiftDown(p2, TRACE_LINE_MUTATION_SIZE);
    double unmutated = scoreLineSegment(p1, p2, flat);
    double up = scoreLineSegment(p1, p2Up, flat);
    double down = scoreLineSegment(p1, p2Down, flat);
    while (true) {
        if (g_margins) { noteMargin(1, unmutated, up); if (unmutated >= up) noteMargin(1, unmutated, down); }
        if (unmutated >= up && unmutated >= down) break;
        else if (up > unmutated) {
            p2Down = p2; down = unmutated;
            p2 = p2Up; unmutated = up;
            p2Up = shiftUp(p2, TRACE_LINE_MUTATION_SIZE);
            up = scoreLineSegment(p1, p2Up, flat);
        } else if (down > unmutated) {
            p2Up = p2; up = unmutated;
            p2 = p2Down; unmutated = down;
            p2Down = shiftDown(p2, TRACE_LINE_MUTATION_SIZE);
            down = scoreLineSegment(p1, p2Down, flat);
        }
    }
    return p2;
}

double getPointDensityScore(int radius, Point p, const Cloud& cloud) {
    PointVector neighbours = radiusSearchAroundPoint(p, radius, cloud);
    double a = 1.0 / SCORE_DISTANCE_FROM_DIAGONAL;
    double score = 0.0;
    for (const Point& n : neighbours) {
        int xDiff = n.x - p.x, yDiff = n.y - p.y;
        score += ((1.0 + a) / (abs(xDiff - yDiff) + 1.0)) - a;
    }
    return score;
}

// Exact density scores of the listed points (independent of each other: spare host cores help).
static void densityScores(int radius, const Cloud& cloud, const std::vector<uint32_t>& which, std::vector<double>& score) {
    const int n = (int)which.size();
    // idle pool workers help (nested loop of the per-read task that called us)
    parallelFor(n, [&](int q) { score[which[(size_t)q]] = getPointDensityScore(radius, cloud.pts[which[(size_t)q]], cloud); },
                n >= 2048 ? 256 : n, 7);
}

Point getHighestDensityPoint(int radius, const Cloud& cloud, const PointVector& common, DiagOrder& diag) {
    // The winner is the first point in cloud order whose score is strictly greater than all before it
    // (semi_global_align.cpp:577-589).  The score of a point sums ((1+a)/(k+1) - a) over its neighbours, k = distance
    // between the diagonals of the two points: only neighbours within 5 diagonals contribute a positive term, and the
    // L1 ball is a square in (x+y, x-y), so an UPPER BOUND of every score comes from counting, per nearby diagonal,
    // the points whose x+y lies within the radius.  Exact scores (the reference's neighbour order and summation) are
    // then only computed for the points whose bound reaches the best exact score found among the top bounds.
    const size_t n = cloud.pts.size();
    std::vector<double> score(n, -1.0);
    std::vector<uint32_t> all;
    if (n < 48) {
        all.resize(n);
        for (size_t i = 0; i < n; ++i) all[i] = (uint32_t)i;
        densityScores(radius, cloud, all, score);
    } else {
        const double a = 1.0 / SCORE_DISTANCE_FROM_DIAGONAL;
        const int KMAX = 5;
        double term[KMAX + 1];
        for (int k = 0; k <= KMAX; ++k) term[k] = std::max(0.0, ((1.0 + a) / (k + 1.0)) - a);
        // (diagonal, anti-diagonal, point) sorted by diagonal, then along the diagonal
        struct DS { int d, s; uint32_t idx; };
        std::vector<DS> ds;
        ds.reserve(n);
        if (diag.order.size() != common.size()) diag.build(common);
        diag.pos.assign(common.size(), -1);
        for (size_t i = 0; i < n; ++i) diag.pos[cloud.orig[i]] = (int32_t)i;
        for (uint32_t o : diag.order) {
            const int32_t i = diag.pos[o];
            if (i >= 0) ds.push_back(DS{common[o].x - common[o].y, common[o].x + common[o].y, (uint32_t)i});
        }
        struct Grp { int d; uint32_t b, e; };
        std::vector<Grp> grp;
        for (size_t i = 0; i < n;) {
            size_t e = i;
            while (e < n && ds[e].d == ds[i].d) ++e;
            grp.push_back(Grp{ds[i].d, (uint32_t)i, (uint32_t)e});
            i = e;
        }
        // per pair (diagonal, diagonal + k): the window [s - radius, s + radius] slides monotonically
        std::vector<double> bound(n, 0.0);
        for (size_t g = 0; g < grp.size(); ++g) {
            size_t h0 = g;
            while (h0 > 0 && grp[h0 - 1].d >= grp[g].d - KMAX) --h0;
            for (size_t h = h0; h < grp.size() && grp[h].d <= grp[g].d + KMAX; ++h) {
                const int k = grp[h].d - grp[g].d;
                const double t = term[k < 0 ? -k : k];
                if (t <= 0.0) continue;
                uint32_t lo = grp[h].b, hi = grp[h].b;
                for (uint32_t q = grp[g].b; q < grp[g].e; ++q) {
                    const int sv = ds[q].s;
                    while (lo < grp[h].e && ds[lo].s < sv - radius) ++lo;
                    while (hi < grp[h].e && ds[hi].s <= sv + radius) ++hi;
                    bound[ds[q].idx] += (double)(hi - lo) * t;
                }
            }
        }
        for (size_t i = 0; i < n; ++i) bound[i] = bound[i] * (1.0 + 1e-9) + 1e-9;
        // exact scores of the 3 largest bounds give the threshold
        std::vector<uint32_t> byBound(n);
        for (size_t i = 0; i < n; ++i) byBound[i] = (uint32_t)i;
        const size_t top = std::min<size_t>(3, n);
        std::partial_sort(byBound.begin(), byBound.begin() + (long)top, byBound.end(),
                          [&](uint32_t x, uint32_t y) { return bound[x] > bound[y]; });
        std::vector<uint32_t> first(byBound.begin(), byBound.begin() + (long)top);
        densityScores(radius, cloud, first, score);
        double threshold = 0.0;   // (scores that are not positive never win: the start value of the scan is 0)
        for (uint32_t i : first) threshold = std::max(threshold, score[i]);
        for (size_t i = 0; i < n; ++i)
            if (score[i] < 0.0 && bound[i] >= threshold) all.push_back((uint32_t)i);
        densityScores(radius, cloud, all, score);
    }
    Point best = cloud.pts[0];
    double bestScore = 0.0;
    for (size_t i = 0; i < n; ++i)
        if (score[i] > bestScore) { bestScore = score[i]; best = cloud.pts[i]; }
    if (g_margins) {
        double second = 0.0;
        bool seen = false;
        for (size_t i = 0; i < n; ++i) {
            if (score[i] == bestScore && !seen) { seen = true; continue; }
            if (score[i] > second) second = score[i];
        }
        noteMargin(0, bestScore, second);
    }
    return best;
}

double variance(std::vector<double>& v) {
    double sum = std::accumulate(v.begin(), v.end(), 0.0);
    double mean = sum / v.size();
    std::vector<double> diff(v.size());
    std::transform(v.begin(), v.end(), diff.begin(), [mean](double x) { return x - mean; });
    double sqSum = std::inner_product(diff.begin(), diff.end(), diff.begin(), 0.0);
    return sqSum / v.size();
}

double getWorstSlope(PointVector traceDots) {
    double worst = 1.0;
    std::sort(traceDots.begin(), traceDots.end());
    for (size_t i = 0; i + 1 < traceDots.size(); ++i) {
        double slope = getSlope(traceDots[i], traceDots[i + 1]);
        if (slope > 1) slope = 1.0 / slope;
        if (slope < worst) worst = slope;
    }
    return worst;
}

double scorePointSet(const PointSet& pointSet, const PointVector& traceDots, bool& failedLine) {
    if (pointSet.size() == 1) return 0.0;
    double pointCount = double(pointSet.size());
    double worstSlopeScore = getWorstSlope(traceDots) * 0.9 + 0.1;
    std::vector<double> xPlusY;
    double mn = std::numeric_limits<int>::max(), mx = std::numeric_limits<int>::min();
    xPlusY.reserve(pointSet.size());
    for (const Point& p : pointSet) {
        double sum = p.x + p.y;
        xPlusY.push_back(sum);
        if (sum > mx) mx = sum;
        if (sum < mn) mn = sum;
    }
    double uniformVariance = (mx - mn) * (mx - mn) / 12.0;
    double varianceScore = variance(xPlusY) / uniformVariance;
    if (varianceScore > 1.0) varianceScore = 1.0 / varianceScore;
    failedLine = (worstSlopeScore * varianceScore) < 0.8;
    if (g_margins) noteMargin(2, worstSlopeScore * varianceScore, 0.8);
    return pointCount * worstSlopeScore * varianceScore;
}

PointSet lineTracing(const PointVector& common, PointSet& usedPoints, const Cloud& cloud, DiagOrder& diag, int readLen, int refLen,
                     int lineNum, int verbosity, std::string& console, bool& failedLine, double& pointSetScore) {
    // The start cloud holds the points no earlier line used: for the first line that is the range's cloud itself
    // (same points in the same order, hence the same tree), so it is only rebuilt from the second line on.
    Cloud rebuilt;
    const long long tA = nowNs();
    if (!usedPoints.empty()) fillCloud(rebuilt, common, usedPoints);
    const Cloud& startCloud = usedPoints.empty() ? cloud : rebuilt;
    const long long tB = nowNs();
    Point startPoint = getHighestDensityPoint(LINE_TRACING_START_POINT_SEARCH_RADIUS, startCloud, common, diag);
    g_seedProf[3] += tB - tA;
    g_seedProf[4] += nowNs() - tB;
    Point p = startPoint;
    PointVector traceDots;
    traceDots.push_back(p);
    const long long tl0 = nowNs();
    PointVector nearby = radiusSearchAroundPoint(p, (int)TRACE_LINE_COLLECTION_DISTANCE, cloud);
    PointSet pointSet(nearby.begin(), nearby.end());
    const int directions[2] = {1, -1};
    for (int direction : directions) {
        p = startPoint;
        const int maxX = readLen, maxY = refLen;
        while (true) {
            int step = direction * TRACE_LINE_STEP_DISTANCE;
            Point previousP = p;
            Point newP(p.x + step, p.y + step);
            bool left = false;
            if (direction == 1 && (newP.x > maxX || newP.y > maxY)) left = true;
            if (direction == -1 && (newP.x < 0 || newP.y < 0)) left = true;
            PointSet pointsNearLine;
            p = mutateLineToBestFitPoints(previousP, newP, cloud, pointsNearLine, left);
            traceDots.push_back(p);
            const long long ta0 = nowNs();
            {   // addPointsNearLine :506-512, over the step's near points in their set order (t_flat, filled by the call above)
                const size_t nNear = t_flat.size();
                if (t_dist.size() < nNear) t_dist.resize(nNear);
                seed::distancesToLineSegment(t_flat.data(), nNear, previousP, p, t_dist.data());
                for (size_t q = 0; q < nNear; ++q)
                    if (t_dist[q] <= TRACE_LINE_COLLECTION_DISTANCE) pointSet.insert(t_flat[q]);
            }
            g_lt[3] += nowNs() - ta0;
            if (left) break;
        }
    }
    const long long tl1 = nowNs();
    pointSetScore = scorePointSet(pointSet, traceDots, failedLine);
    const long long tl2 = nowNs();
    g_lt[4] += tl1 - tl0; g_lt[5] += tl2 - tl1;
    if (verbosity > 2) {
        console += "    line " + std::to_string(lineNum + 1) + ": " + std::to_string(pointSet.size()) + " points, " +
                   "score=" + std::to_string(pointSetScore) + " (" + (failedLine ? "bad" : "good") + ")\n";
    }
    const long long tl3 = nowNs();
    for (const Point& q : pointSet) usedPoints.insert(q);
    g_lt[6] += nowNs() - tl3;
    return pointSet;
}

// ------------------------------------------------------------------ seeds (SeqAn Seed<Simple>, Unordered SeedSet)
// seeds_combination.h:102-124 (Merge): b right of / overlapping a, diagonals at most maxDiag apart
bool combineable(const ChainSeed& a, const ChainSeed& b, unsigned maxDiag) {
    if (b.beginH < a.beginH || b.beginV < a.beginV) return false;
    if (b.beginH > a.endH || b.beginV > a.endV) return false;
    long d = (a.endH - a.endV) - (b.beginH - b.beginV);
    if ((unsigned)std::labs(d) > maxDiag) return false;
    return true;
}

void mergeInto(ChainSeed& seed, const ChainSeed& other) {  // seeds_combination.h:186-197
    seed.beginH = std::min(seed.beginH, other.beginH);
    seed.beginV = std::min(seed.beginV, other.beginV);
    seed.endH = std::max(seed.endH, other.endH);
    seed.endV = std::max(seed.endV, other.endV);
    seed.lowerDiag = std::min(seed.lowerDiag, other.lowerDiag);
    seed.upperDiag = std::max(seed.upperDiag, other.upperDiag);
}

// SeqAn's SeedSet<Simple, Unordered> is a std::multiset ordered by begin diagonal (equal keys keep insertion order),
// and addSeed(set, seed, maxDiag, Merge()) walks it from the start and merges the new seed into the FIRST element it
// can be combined with, in either direction (seeds_seed_set_unordered.h:248-343) — a linear scan per seed, quadratic
// per point set (15 ms of the 32 ms a 20 kb read costs on the host).  Same elements, same order, same choice here,
// without the scan: an element that can combine with the new seed has its END diagonal within maxDiag of the seed's
// begin diagonal (element left of the seed) or its BEGIN diagonal within maxDiag of the seed's end diagonal (seed left
// of the element), so only the elements filed under those 2 * (2 * maxDiag + 1) diagonals are tested, and the one that
// comes first in the multiset's order (begin diagonal, then insertion sequence) wins.  The files are arrays over the
// range's diagonals, the elements of one diagonal linked through the nodes (no hashing).
class SeedSet {
public:
    // the diagonals (H - V) of every seed of the range lie in [-maxV, maxH]
    SeedSet(long maxH, long maxV)
        : diag0_(-maxV), headBegin_((size_t)(maxH + maxV + 1), -1), headEnd_((size_t)(maxH + maxV + 1), -1) {
        nodes_.reserve(1024);
    }
    bool addMerge(const ChainSeed& seed, unsigned maxDiag) {
        const long bd = seed.beginH - seed.beginV, ed = seed.endH - seed.endV;
        int best = -1;
        auto consider = [&](int id) {
            if (best < 0 || before(id, best)) best = id;
        };
        // Seeds arrive with ascending beginH (the caller sorts the points): an element that ends left of this seed's
        // begin can never be the left partner again, one that begins left of it never the right partner.  Such
        // elements leave the files here (they stay in the set), which keeps the lists at a handful of elements.
        const long last = (long)headBegin_.size() - 1;
        for (long d = std::max(0L, bd - (long)maxDiag - diag0_); d <= std::min(last, bd + (long)maxDiag - diag0_); ++d) {
            int32_t* link = &headEnd_[(size_t)d];
            while (*link >= 0) {
                Node& nd = nodes_[(size_t)*link];
                if (nd.s.endH < seed.beginH) { *link = nd.nextEnd; continue; }
                if (combineable(nd.s, seed, maxDiag)) consider(*link);
                link = &nd.nextEnd;
            }
        }
        for (long d = std::max(0L, ed - (long)maxDiag - diag0_); d <= std::min(last, ed + (long)maxDiag - diag0_); ++d) {
            int32_t* link = &headBegin_[(size_t)d];
            while (*link >= 0) {
                Node& nd = nodes_[(size_t)*link];
                if (nd.s.beginH < seed.beginH) { *link = nd.nextBegin; continue; }
                if (combineable(seed, nd.s, maxDiag)) consider(*link);
                link = &nd.nextBegin;
            }
        }
        if (best < 0) return false;
        ChainSeed merged = nodes_[(size_t)best].s;   // (the merge is symmetric: min of the begins, max of the ends)
        mergeInto(merged, seed);
        erase(best);
        insert(merged);
        return true;
    }
    void insert(const ChainSeed& s) {
        const int id = (int)nodes_.size();
        int32_t& hb = headBegin_[slot(s.beginH - s.beginV)];
        int32_t& he = headEnd_[slot(s.endH - s.endV)];
        nodes_.push_back(Node{s, true, hb, he});
        hb = id;
        he = id;
    }
    // the elements in the multiset's iteration order
    std::vector<ChainSeed> ordered() const {
        std::vector<int> ids;
        for (size_t i = 0; i < nodes_.size(); ++i)
            if (nodes_[i].alive) ids.push_back((int)i);
        std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return beginDiag(a) < beginDiag(b); });
        std::vector<ChainSeed> out;
        out.reserve(ids.size());
        for (int id : ids) out.push_back(nodes_[(size_t)id].s);
        return out;
    }

private:
    // alive elements are filed under their begin and their end diagonal (lists of a handful of elements; which one
    // wins does not depend on their order in the list)
    struct Node { ChainSeed s; bool alive; int32_t nextBegin, nextEnd; };
    size_t slot(long d) const { return (size_t)(d - diag0_); }
    long beginDiag(int id) const { return nodes_[(size_t)id].s.beginH - nodes_[(size_t)id].s.beginV; }
    // ids grow with the insertion sequence (a merged seed is erased and re-inserted: it moves behind its equals)
    bool before(int a, int b) const {
        const long da = beginDiag(a), db = beginDiag(b);
        return da < db || (da == db && a < b);
    }
    void erase(int id) {
        Node& n = nodes_[(size_t)id];
        n.alive = false;
        unlink(headBegin_[slot(n.s.beginH - n.s.beginV)], id, &Node::nextBegin);
        unlink(headEnd_[slot(n.s.endH - n.s.endV)], id, &Node::nextEnd);
    }
    void unlink(int32_t& head, int id, int32_t Node::*next) {
        for (int32_t* link = &head; *link >= 0; link = &(nodes_[(size_t)*link].*next))   // (it may have left the file already)
            if (*link == id) { *link = nodes_[(size_t)id].*next; return; }
    }
    long diag0_;
    std::vector<Node> nodes_;
    std::vector<int32_t> headBegin_, headEnd_;
};

// seeds_global_chaining.h:102-280 (Gusfield sparse chaining, maximising the summed seed sizes)
void chainSeedsGlobally(std::vector<ChainSeed>& target, const SeedSet& seedSet) {
    typedef unsigned long TPos;
    const std::vector<ChainSeed> seeds = seedSet.ordered();
    struct IntervalPoint {
        TPos pos; bool isBegin; unsigned idx;
        bool operator<(const IntervalPoint& o) const {
            if (pos < o.pos) return true;
            if (pos == o.pos && isBegin < o.isBegin) return true;
            if (pos == o.pos && isBegin == o.isBegin && idx < o.idx) return true;
            return false;
        }
    };
    struct Solution {
        TPos endV; TPos quality; unsigned idx;
        bool operator<(const Solution& o) const {
            if (endV < o.endV) return true;
            if (endV == o.endV && quality < o.quality) return true;
            if (endV == o.endV && quality == o.quality && idx < o.idx) return true;
            return false;
        }
    };
    const unsigned NONE = std::numeric_limits<unsigned>::max();
    std::vector<IntervalPoint> points;
    points.reserve(2 * seeds.size());
    std::vector<TPos> quality(seeds.size());          // (the reference keeps these two in std::map keyed by seed index)
    std::vector<unsigned> predecessor(seeds.size());
    for (unsigned i = 0; i < seeds.size(); ++i) {
        quality[i] = (TPos)std::max(seeds[i].endH - seeds[i].beginH, seeds[i].endV - seeds[i].beginV);
        predecessor[i] = NONE;
        points.push_back(IntervalPoint{(TPos)seeds[i].beginH, true, i});
        points.push_back(IntervalPoint{(TPos)seeds[i].endH, false, i});
    }
    std::sort(points.begin(), points.end());
    std::multiset<Solution> sols;
    for (const IntervalPoint& pt : points) {
        const ChainSeed& seedK = seeds[pt.idx];
        if (pt.isBegin) {
            Solution ref{(TPos)seedK.beginV, std::numeric_limits<TPos>::max(), NONE};
            std::multiset<Solution>::iterator itJ = sols.upper_bound(ref);
            if (itJ == sols.begin()) {
                if (sols.size() > 0 && sols.rbegin()->endV <= (TPos)seedK.beginV) { itJ = sols.end(); --itJ; }
                else continue;
            } else
                --itJ;
            quality[pt.idx] += itJ->quality;
            predecessor[pt.idx] = itJ->idx;
        } else {
            Solution ref{(TPos)seedK.endV, 0, NONE};
            std::multiset<Solution>::iterator itSol = sols.upper_bound(ref);
            if (itSol == sols.end()) {
                sols.insert(Solution{(TPos)seedK.endV, quality[pt.idx], pt.idx});
            } else {
                const ChainSeed& seedJ = seeds[itSol->idx];
                if (seedJ.endV > seedK.endV || (seedJ.endV == seedK.endV && quality[pt.idx] > itSol->quality))
                    sols.insert(Solution{(TPos)seedK.endV, quality[pt.idx], pt.idx});
            }
            std::multiset<Solution>::iterator itDel = sols.upper_bound(ref);
            while (itDel != sols.end()) {
                std::multiset<Solution>::iterator ptr = itDel;
                ++itDel;
                if (quality[pt.idx] > ptr->quality) sols.erase(ptr);
            }
        }
    }
    target.clear();
    if (sols.empty()) return;
    unsigned next = sols.rbegin()->idx;
    while (next != NONE) {
        target.push_back(seeds[next]);
        next = predecessor[next];
    }
    std::reverse(target.begin(), target.end());
}

// std::sort(points) of the reference (by x, then y; the points of a set are distinct, so every correct sort gives the
// same sequence): least-significant-digit radix sort on the packed (x, y) key, passes over constant digits skipped.
void sortPoints(PointVector& pts) {
    const size_t n = pts.size();
    bool packable = n >= 64;
    for (size_t i = 0; i < n && packable; ++i) packable = pts[i].x >= 0 && pts[i].y >= 0;
    if (!packable) { std::sort(pts.begin(), pts.end()); return; }
    static thread_local std::vector<uint64_t> a, b;
    a.resize(n); b.resize(n);
    uint64_t all1 = ~0ull, any1 = 0;
    for (size_t i = 0; i < n; ++i) {
        a[i] = ((uint64_t)(uint32_t)pts[i].x << 32) | (uint32_t)pts[i].y;
        all1 &= a[i]; any1 |= a[i];
    }
    const uint64_t varying = all1 ^ any1;   // bits that differ between some two keys
    uint64_t* src = a.data();
    uint64_t* dst = b.data();
    for (int shift = 0; shift < 64; shift += 11) {
        if (((varying >> shift) & 0x7ff) == 0) continue;
        size_t count[2049] = {0};
        for (size_t i = 0; i < n; ++i) ++count[((src[i] >> shift) & 0x7ff) + 1];
        for (int d = 0; d < 2048; ++d) count[d + 1] += count[d];
        for (size_t i = 0; i < n; ++i) dst[count[(src[i] >> shift) & 0x7ff]++] = src[i];
        std::swap(src, dst);
    }
    for (size_t i = 0; i < n; ++i) pts[i] = Point((int)(src[i] >> 32), (int)(uint32_t)src[i]);
}

long long maxSeedChainGapArea(const std::vector<ChainSeed>& chain, int readLen, int refLen) {  // :321-347
    int prevH = 0, prevV = 0;
    long long maxArea = 0;
    const int n = (int)chain.size();
    for (int i = 0; i <= n; ++i) {
        int hPos = (i == n) ? readLen : (int)chain[(size_t)i].beginH;
        int vPos = (i == n) ? refLen : (int)chain[(size_t)i].beginV;
        long long area = (long long)(hPos - prevH) * (long long)(vPos - prevV);
        if (area > maxArea) maxArea = area;
        if (i < n) { prevH = (int)chain[(size_t)i].endH; prevV = (int)chain[(size_t)i].endV; }
    }
    return maxArea;
}

}  // namespace

// common k-mers on the host (semi_global_align.cpp:197-207): window positions ascending, read positions ascending
static void hostCommonKmers(const KmerPosMap& readKmers, const std::string& trimmedRefSeq, int kSize, PointVector& common) {
    const int refLen = (int)trimmedRefSeq.size();
    if (kSize > 16) {
        const int maxI = refLen - kSize + 1;
        std::string kmer;
        for (int i = 0; i < maxI; ++i) {
            kmer.assign(trimmedRefSeq, (size_t)i, (size_t)kSize);
            auto it = readKmers.other.find(kmer);
            if (it != readKmers.other.end())
                for (int pos : it->second) common.push_back(Point(pos, i));
        }
    } else {
        std::string kmer;
        forEachKmer(trimmedRefSeq, kSize, [&](int i, bool ok, uint32_t code) {
            if (ok) {
                if (readKmers.head.empty()) return;
                uint32_t sl = kmerHash(code) & readKmers.mask;
                while (readKmers.head[sl] != -1 && readKmers.key[sl] != code) sl = (sl + 1) & readKmers.mask;
                for (int pos = readKmers.head[sl]; pos != -1; pos = readKmers.next[(size_t)pos]) common.push_back(Point(pos, i));
            } else if (!readKmers.other.empty()) {
                kmer.assign(trimmedRefSeq, (size_t)i, (size_t)kSize);
                auto it = readKmers.other.find(kmer);
                if (it != readKmers.other.end())
                    for (int pos : it->second) common.push_back(Point(pos, i));
            }
        });
    }
}

void commonKmerPoints(const std::string& readSeq, const std::string& trimmedRefSeq, int kSize, std::vector<int32_t>& xy) {
    KmerPosMap kmers;
    buildKmerPositions(readSeq, kSize, kmers);
    PointVector common;
    hostCommonKmers(kmers, trimmedRefSeq, kSize, common);
    xy.clear();
    xy.reserve(2 * common.size());
    for (const Point& p : common) { xy.push_back(p.x); xy.push_back(p.y); }
}

void seedRange(const std::string& readSeq, const KmerPosMap& readKmers, const std::string& trimmedRefSeq,
               const SensitivityParams& sp, int verbosity, const std::string& refName, int refStart, int refEnd,
               RangeSeeds& out, const int32_t* joinedXY, size_t nJoined) {
    out.chains.clear();
    seed::Arena::mine().rewind();   // (every point set of the previous range on this thread is gone)
    const int kSize = sp.kSize;
    const int readLen = (int)readSeq.size(), refLen = (int)trimmedRefSeq.size();
    if (verbosity > 2)
        out.console += "Range: " + refName + ": " + std::to_string(refStart) + " - " + std::to_string(refEnd) + "\n";
    // common k-mers :197-207
    const long long t0 = nowNs();
    PointVector common;
    if (joinedXY) {   // the batch path joined the k-mers on the device (kmerjoin.cu): same points, same order
        common.resize(nJoined);
        for (size_t q = 0; q < nJoined; ++q) common[q] = Point(joinedXY[2 * q], joinedXY[2 * q + 1]);
    } else if (readKmers.k != kSize) return;
    else hostCommonKmers(readKmers, trimmedRefSeq, kSize, common);
    if (verbosity > 2)
        out.console += "    common " + std::to_string(kSize) + "-mers: " + std::to_string(common.size()) + "\n";
    if (common.empty()) return;  // the reference dereferences an empty vector here (undefined); no alignment
    const long long t1 = nowNs();
    g_seedProf[0] += t1 - t0;
    PointSet usedPoints;
    Cloud cloud;
    fillCloud(cloud, common, usedPoints);
    g_lt[7] += nowNs() - t1;
    g_lastSearch.cloud = nullptr;   // (a new cloud may live at the old one's address)
    std::vector<PointSet> goodPointSets;
    DiagOrder diag;
    double bestPointScore = 0.0;
    for (int lineNum = 0; lineNum < sp.maxLineTraceCount; ++lineNum) {
        bool failedLine = false;
        double pointSetScore = 0.0;
        PointSet pointSet = lineTracing(common, usedPoints, cloud, diag, readLen, refLen, lineNum, verbosity, out.console,
                                        failedLine, pointSetScore);
        if (g_margins && lineNum > 0) noteMargin(3, pointSetScore, bestPointScore);
        if (pointSetScore > bestPointScore) bestPointScore = pointSetScore;
        if (!failedLine) goodPointSets.push_back(pointSet);
        else if (pointSetScore == bestPointScore) goodPointSets.push_back(pointSet);
        if (!failedLine && lineNum >= sp.minLineTraceCount - 1) break;
        if (usedPoints.size() >= common.size()) break;
    }
    const long long t2 = nowNs();
    g_seedProf[1] += t2 - t1;
    if (g_margins) {
        fprintf(stderr, "[ub200 margins] points=%zu density %.3e mutation %.3e good/bad %.3e best-line %.3e\n", common.size(),
                t_minMargin[0], t_minMargin[1], t_minMargin[2], t_minMargin[3]);
        for (int q = 0; q < 4; ++q) t_minMargin[q] = 1e300;
    }
    struct Fin { long long t; ~Fin() { g_seedProf[2] += nowNs() - t; } } fin{t2};
    for (const PointSet& good : goodPointSets) {
        PointVector pts;
        pts.reserve(good.size());
        for (const Point& p : good) pts.push_back(p);
        sortPoints(pts);
        SeedSet seedSet((long)readLen + kSize, (long)refLen + kSize);
        for (const Point& p : pts) {
            ChainSeed s{(long)p.x, (long)p.y, (long)p.x + kSize, (long)p.y + kSize, (long)p.x - (long)p.y, (long)p.x - (long)p.y};
            if (!seedSet.addMerge(s, 2)) seedSet.insert(s);
        }
        std::vector<ChainSeed> chain;
        chainSeedsGlobally(chain, seedSet);
        if (chain.empty()) return;
        if (maxSeedChainGapArea(chain, readLen, refLen) > MAX_BANDED_ALIGNMENT_GAP_AREA) return;
        out.chains.push_back(chain);
    }
}

}  // namespace ub200
