// oracle/image.cpp -- Frame::InitFrame / CreateImagePyramid (reference src/Basic/Frame.cpp:22-40).
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// The arithmetic lives in OpenCV (not in the reference tree; "3.1 or higher", Readme.md:13-14):
//   cv::cvtColor(BGR2GRAY) 8u : fixed point, 15-bit coefficients in OpenCV 4.x
//                               gray = (B*3735 + G*19235 + R*9798 + 16384) >> 15
//   cv::pyrDown 8u            : separable [1 4 6 4 1], (sum + 128) >> 8, BORDER_REFLECT_101,
//                               dst = ((w+1)/2, (h+1)/2)
// Both are pinned bit-exactly against cv2 4.13 in tests/test_oracle_cv2.py.
#include "oracle.h"

#include <vector>

extern "C" void ora_bgr2gray(const uint8_t* bgr, int w, int h, uint8_t* gray) {
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i) {
        const int b = bgr[3 * i], g = bgr[3 * i + 1], r = bgr[3 * i + 2];
        gray[i] = (uint8_t)((b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15);
    }
}

static inline int reflect101(int i, int n) {
    // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba ; n == 1 collapses to 0
    if (n == 1) return 0;
    while (i < 0 || i >= n) {
        if (i < 0) i = -i;
        else i = 2 * (n - 1) - i;
    }
    return i;
}

extern "C" void ora_pyrdown(const uint8_t* src, int w, int h, uint8_t* dst) {
    const int dw = (w + 1) / 2, dh = (h + 1) / 2;
    // horizontal pass on every source row: row_h[y][dx] = sum_k w_k * src[y][reflect(2dx + k - 2)]
    std::vector<int> hbuf((size_t)h * dw);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * w;
        for (int dx = 0; dx < dw; ++dx) {
            const int c = 2 * dx;
            hbuf[(size_t)y * dw + dx] = s[reflect101(c - 2, w)] + 4 * s[reflect101(c - 1, w)] + 6 * s[reflect101(c, w)] +
                                        4 * s[reflect101(c + 1, w)] + s[reflect101(c + 2, w)];
        }
    }
    for (int dy = 0; dy < dh; ++dy) {
        const int c = 2 * dy;
        const int* r0 = &hbuf[(size_t)reflect101(c - 2, h) * dw];
        const int* r1 = &hbuf[(size_t)reflect101(c - 1, h) * dw];
        const int* r2 = &hbuf[(size_t)reflect101(c, h) * dw];
        const int* r3 = &hbuf[(size_t)reflect101(c + 1, h) * dw];
        const int* r4 = &hbuf[(size_t)reflect101(c + 2, h) * dw];
        for (int dx = 0; dx < dw; ++dx) {
            const int v = r0[dx] + 4 * r1[dx] + 6 * r2[dx] + 4 * r3[dx] + r4[dx];
            dst[(size_t)dy * dw + dx] = (uint8_t)((v + 128) >> 8);
        }
    }
}

extern "C" size_t ora_pyramid_layout(int w, int h, int n_levels, int* lw, int* lh, size_t* off) {
    size_t total = 0;
    for (int L = 0; L < n_levels; ++L) {
        if (lw) lw[L] = w;
        if (lh) lh[L] = h;
        if (off) off[L] = total;
        total += (size_t)w * h;
        w = (w + 1) / 2;
        h = (h + 1) / 2;
    }
    return total;
}

// Frame::CreateImagePyramid (Frame.cpp:31-40): level i = pyrDown(level i-1), level 0 = gray
extern "C" void ora_build_pyramid(const uint8_t* gray, int w, int h, int n_levels, uint8_t* pyr) {
    int lw[ORA_MAX_LEVELS], lh[ORA_MAX_LEVELS];
    size_t off[ORA_MAX_LEVELS];
    ora_pyramid_layout(w, h, n_levels, lw, lh, off);
    for (size_t i = 0; i < (size_t)w * h; ++i) pyr[i] = gray[i];
    for (int L = 1; L < n_levels; ++L) ora_pyrdown(pyr + off[L - 1], lw[L - 1], lh[L - 1], pyr + off[L]);
}
