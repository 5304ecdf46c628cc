// Distance from points to a line segment (unicycler/src/semi_global_align.cpp:756-785, getDistanceToLineSegment), one
// point at a time and two at a time.  The pair version runs the SAME IEEE-754 double operations in the same order per
// lane (SSE2 packed add / mul / div / sqrt round exactly like the scalar SSE2 instructions the compiler emits for the
// scalar version; there is no fused multiply-add without -mfma), so both give bit-identical distances — checked on
// random and degenerate segments by tests/cpp/test_pointset.cpp.  The caller adds the terms up in the original order.
#pragma once
#include <cmath>
#include <cstddef>

#include "pointset.hpp"

#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace ub200 {
namespace seed {

inline double distanceToLineSegment(Point p, Point l1, Point l2) {
    double A = p.x - l1.x, B = p.y - l1.y, C = l2.x - l1.x, D = l2.y - l1.y;
    double dot = A * C + B * D;
    double lenSq = C * C + D * D;
    double param = -1;
    if (lenSq != 0) param = dot / lenSq;
    double xx, yy;
    if (param < 0) { xx = l1.x; yy = l1.y; }
    else if (param > 1) { xx = l2.x; yy = l2.y; }
    else { xx = l1.x + param * C; yy = l1.y + param * D; }
    double dx = p.x - xx, dy = p.y - yy;
    return std::sqrt(dx * dx + dy * dy);
}

// out[i] = distanceToLineSegment(pts[i], l1, l2) for i in [0, n)
inline void distancesToLineSegment(const Point* pts, size_t n, Point l1, Point l2, double* out) {
    size_t i = 0;
#if defined(__SSE2__)
    const double Cs = l2.x - l1.x, Ds = l2.y - l1.y;
    const double lenSqS = Cs * Cs + Ds * Ds;
    const __m128d C = _mm_set1_pd(Cs), D = _mm_set1_pd(Ds), lenSq = _mm_set1_pd(lenSqS);
    const __m128d l1x = _mm_set1_pd((double)l1.x), l1y = _mm_set1_pd((double)l1.y);
    const __m128d l2x = _mm_set1_pd((double)l2.x), l2y = _mm_set1_pd((double)l2.y);
    const __m128d zero = _mm_setzero_pd(), one = _mm_set1_pd(1.0), minusOne = _mm_set1_pd(-1.0);
    const __m128i l1xi = _mm_set1_epi32(l1.x), l1yi = _mm_set1_epi32(l1.y);
    for (; i + 2 <= n; i += 2) {
        // two points = four consecutive ints (x0, y0, x1, y1)
        const __m128i xy = _mm_loadu_si128((const __m128i*)(pts + i));
        const __m128i xi = _mm_shuffle_epi32(xy, _MM_SHUFFLE(3, 3, 2, 0));   // lanes 0, 1 = x0, x1
        const __m128i yi = _mm_shuffle_epi32(xy, _MM_SHUFFLE(3, 3, 3, 1));   // lanes 0, 1 = y0, y1
        const __m128d A = _mm_cvtepi32_pd(_mm_sub_epi32(xi, l1xi));          // (double)(p.x - l1.x): int subtraction first
        const __m128d B = _mm_cvtepi32_pd(_mm_sub_epi32(yi, l1yi));
        const __m128d px = _mm_cvtepi32_pd(xi), py = _mm_cvtepi32_pd(yi);
        const __m128d dot = _mm_add_pd(_mm_mul_pd(A, C), _mm_mul_pd(B, D));
        const __m128d param = (lenSqS != 0) ? _mm_div_pd(dot, lenSq) : minusOne;
        const __m128d below = _mm_cmplt_pd(param, zero);                      // param < 0
        const __m128d above = _mm_andnot_pd(below, _mm_cmpgt_pd(param, one)); // else param > 1
        const __m128d midX = _mm_add_pd(l1x, _mm_mul_pd(param, C)), midY = _mm_add_pd(l1y, _mm_mul_pd(param, D));
        __m128d xx = _mm_or_pd(_mm_and_pd(below, l1x), _mm_andnot_pd(below, _mm_or_pd(_mm_and_pd(above, l2x), _mm_andnot_pd(above, midX))));
        __m128d yy = _mm_or_pd(_mm_and_pd(below, l1y), _mm_andnot_pd(below, _mm_or_pd(_mm_and_pd(above, l2y), _mm_andnot_pd(above, midY))));
        const __m128d dx = _mm_sub_pd(px, xx), dy = _mm_sub_pd(py, yy);
        _mm_storeu_pd(out + i, _mm_sqrt_pd(_mm_add_pd(_mm_mul_pd(dx, dx), _mm_mul_pd(dy, dy))));
    }
#endif
    for (; i < n; ++i) out[i] = distanceToLineSegment(pts[i], l1, l2);
}

}  // namespace seed
}  // namespace ub200
