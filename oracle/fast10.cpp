// oracle/fast10.cpp -- FAST-10 segment test, bisection score, 3x3 non-maximum suppression.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// The reference calls the uzh-rpg `fast` library (github.com/uzh-rpg/fast, unpinned: "git clone",
// Readme.md:7-8) at src/Algorithm/FeatureDetector.cpp:365-381:
//     fast::fast_corner_detect_10_sse2(img, cols, rows, cols, threshold, corners)
//     fast::fast_corner_score_10(img, cols, corners, threshold, scores)
//     fast::fast_nonmax_3x3(corners, scores, nm_corners)
// Its sources are NOT in /root/reference, so this file restates Rosten's published algorithm
// (SURVEY.md appendix A.1).  PARITY UNPINNED: no reference test asserts a FAST output; the
// restatement is pinned structurally (ring geometry vs. cv2's FAST-9 with n switched to 9,
// closed-form score vs. bisection, nonmax vs. a dense 8-neighbour check; tests/test_oracle_fast.py).
#include "oracle.h"

#include <cstdlib>
#include <vector>
#if defined(__AVX2__)
#include <immintrin.h>   // the timed CPU-baseline build (-march=x86-64-v3) gets a 32-pixel prefilter like the SSE2 path of the library
#endif

namespace {

// Bresenham circle of radius 3, circular order (dx,dy)
const int kRingDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
const int kRingDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

inline bool has_run(unsigned m16, int n) {
    // >= n contiguous set bits on the 16-bit circle
    unsigned x = m16 | (m16 << 16);
    unsigned run = x;
    for (int k = 1; k < n; ++k) run &= (x >> k);
    return (run & 0xFFFFu) != 0;
}

// segment test at pixel p with barrier b: n contiguous ring pixels all > p+b or all < p-b (strict)
inline bool is_corner(const uint8_t* p, const int* off, int b, int n) {
    const int cb = *p + b, c_b = *p - b;
    // quick rejection: a 10-arc (also a 9-arc) always contains one pixel of each opposite pair
    const int v0 = p[off[0]], v8 = p[off[8]];
    if (!(v0 > cb || v8 > cb || v0 < c_b || v8 < c_b)) return false;
    unsigned bright = 0, dark = 0;
    for (int i = 0; i < 16; ++i) {
        const int v = p[off[i]];
        bright |= (unsigned)(v > cb) << i;
        dark |= (unsigned)(v < c_b) << i;
    }
    return has_run(bright, n) || has_run(dark, n);
}

inline void make_offsets(int* off, int stride) {
    for (int i = 0; i < 16; ++i) off[i] = kRingDy[i] * stride + kRingDx[i];
}

}  // namespace

// generic arc length so that tests can pin the ring/arc logic against cv2's TYPE_9_16
extern "C" int ora_fastN_detect(const uint8_t* img, int w, int h, int stride, int barrier, int arc, int16_t* xy,
                                int cap) {
    int off[16];
    make_offsets(off, stride);
    int n = 0;
    // fast_corner_detect_10: for y in [3,h-3) for x in [3,w-3), raster order (the SSE2 variant walks
    // 16-pixel blocks plus a scalar tail and emits the same set in the same order)
    auto emit = [&](int x, int y) {
        if (n < cap) {
            xy[2 * n] = (int16_t)x;
            xy[2 * n + 1] = (int16_t)y;
        }
        ++n;
    };
    for (int y = 3; y < h - 3; ++y) {
        int x = 3;
#if defined(__AVX2__)
        // 32 pixels at a time (the reference calls fast_corner_detect_10_sse2, FeatureDetector.cpp:365-368, which filters 16
        // pixels per step the same way): an arc of >= 9 ring pixels contains at least two of the four compass pixels, so a
        // pixel with fewer than two brighter AND fewer than two darker compass pixels cannot be a corner; the survivors take
        // the exact scalar test, in raster order -- same set, same order as the scalar loop
        if (arc >= 9 && barrier >= 0 && barrier <= 255) {
            const __m256i vb = _mm256_set1_epi8((char)barrier), zero = _mm256_setzero_si256(), one = _mm256_set1_epi8(1);
            const int comp[4] = {off[0], off[4], off[8], off[12]};
            for (; x + 32 <= w - 3; x += 32) {
                const uint8_t* p = img + (size_t)y * stride + x;
                const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
                const __m256i hi = _mm256_adds_epu8(c, vb), lo = _mm256_subs_epu8(c, vb);
                __m256i nb = zero, nd = zero;   // number of brighter / darker compass pixels per lane
                for (int k = 0; k < 4; ++k) {
                    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + comp[k]));
                    // v > hi  <=>  subs(v, hi) != 0 ; v < lo  <=>  subs(lo, v) != 0  (the saturation keeps p + b > 255 / p - b < 0 exact)
                    nb = _mm256_add_epi8(nb, _mm256_andnot_si256(_mm256_cmpeq_epi8(_mm256_subs_epu8(v, hi), zero), one));
                    nd = _mm256_add_epi8(nd, _mm256_andnot_si256(_mm256_cmpeq_epi8(_mm256_subs_epu8(lo, v), zero), one));
                }
                unsigned m = (unsigned)_mm256_movemask_epi8(_mm256_or_si256(_mm256_cmpgt_epi8(nb, one), _mm256_cmpgt_epi8(nd, one)));
                while (m) {
                    const int k = __builtin_ctz(m);
                    m &= m - 1;
                    if (is_corner(p + k, off, barrier, arc)) emit(x + k, y);
                }
            }
        }
#endif
        for (; x < w - 3; ++x)
            if (is_corner(img + (size_t)y * stride + x, off, barrier, arc)) emit(x, y);
    }
    return n;
}

extern "C" int ora_fast10_detect(const uint8_t* img, int w, int h, int stride, int barrier, int16_t* xy, int cap) {
    return ora_fastN_detect(img, w, h, stride, barrier, 10, xy, cap);
}

// fast_corner_score_10: largest barrier for which the pixel is still a corner, by bisection on
// [barrier, 255]:  bmin=b0, bmax=255, b=(bmin+bmax)/2; loop { corner(b) ? bmin=b : bmax=b;
//                  if (bmin==bmax-1 || bmin==bmax) return bmin; b=(bmin+bmax)/2; }
extern "C" void ora_fast10_score(const uint8_t* img, int stride, const int16_t* xy, int n, int barrier,
                                 int32_t* scores) {
    int off[16];
    make_offsets(off, stride);
    for (int i = 0; i < n; ++i) {
        const uint8_t* p = img + (size_t)xy[2 * i + 1] * stride + xy[2 * i];
        int bmin = barrier, bmax = 255, b = (bmax + bmin) / 2;
        for (;;) {
            if (is_corner(p, off, b, 10)) bmin = b;
            else bmax = b;
            if (bmin == bmax - 1 || bmin == bmax) break;
            b = (bmin + bmax) / 2;
        }
        scores[i] = bmin;
    }
}

// fast_nonmax_3x3: corners are in raster order; corner i survives iff none of its 8-neighbour
// corners has score >= score[i] (non-strict: two equal neighbours kill each other).  Output =
// indices into the corner list, ascending.  Walks the list with row-start pointers like the
// library does (left / right neighbours are i-1 / i+1, rows above and below are scanned with
// monotone cursors).
// The neighbour rule of uzh-rpg/fast is restated from memory (its source is not in the reference tree): ">=" (a corner dies when
// a neighbour scores at least as much) is the default; ora_set_fast_nonmax_strict(1) or YGZ_ORACLE_NONMAX_STRICT=1 switches the
// oracle to ">" (only a strictly better neighbour kills), so that the choice is one flag once the library source can be read.
static int g_nonmax_strict = -1;
extern "C" void ora_set_fast_nonmax_strict(int strict) { g_nonmax_strict = strict ? 1 : 0; }
static inline bool nonmax_beats(int other, int s) {
    if (g_nonmax_strict < 0) {
        const char* e = getenv("YGZ_ORACLE_NONMAX_STRICT");
        g_nonmax_strict = (e && e[0] == '1') ? 1 : 0;
    }
    return g_nonmax_strict ? other > s : other >= s;
}

extern "C" int ora_fast_nonmax_3x3(const int16_t* xy, const int32_t* scores, int n, int32_t* keep_idx) {
    if (n < 1) return 0;
    const int last_row = xy[2 * (n - 1) + 1];
    std::vector<int> row_start(last_row + 2, -1);
    {
        int prev_row = -1;
        for (int i = 0; i < n; ++i)
            if (xy[2 * i + 1] != prev_row) {
                row_start[xy[2 * i + 1]] = i;
                prev_row = xy[2 * i + 1];
            }
    }
    int n_keep = 0;
    int above = 0, below = 0;
    for (int i = 0; i < n; ++i) {
        const int s = scores[i], x = xy[2 * i], y = xy[2 * i + 1];
        bool dead = false;
        if (i > 0 && xy[2 * (i - 1) + 1] == y && xy[2 * (i - 1)] == x - 1 && nonmax_beats(scores[i - 1], s)) dead = true;
        if (!dead && i < n - 1 && xy[2 * (i + 1) + 1] == y && xy[2 * (i + 1)] == x + 1 && nonmax_beats(scores[i + 1], s)) dead = true;
        if (!dead && y > 0 && row_start[y - 1] != -1) {
            if (xy[2 * above + 1] < y - 1) above = row_start[y - 1];
            while (xy[2 * above + 1] < y && xy[2 * above] < x - 1) ++above;
            for (int j = above; xy[2 * j + 1] < y && xy[2 * j] <= x + 1; ++j)
                if (nonmax_beats(scores[j], s)) {  // x in {x-1, x, x+1} guaranteed by the two loop bounds
                    dead = true;
                    break;
                }
        }
        if (!dead && y != last_row && row_start[y + 1] != -1 && below < n) {
            if (xy[2 * below + 1] < y + 1) below = row_start[y + 1];
            while (below < n && xy[2 * below + 1] == y + 1 && xy[2 * below] < x - 1) ++below;
            for (int j = below; j < n && xy[2 * j + 1] == y + 1 && xy[2 * j] <= x + 1; ++j)
                if (nonmax_beats(scores[j], s)) {
                    dead = true;
                    break;
                }
        }
        if (!dead) keep_idx[n_keep++] = i;
    }
    return n_keep;
}
