// oracle/detector.cpp -- FeatureDetector::{Detect,ShiTomasiScore,IC_Angle,ComputeOrbDescriptor,
// ComputeAngleAndDescriptor} (reference src/Algorithm/FeatureDetector.cpp:299-596).
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Parity hazards pinned here (SURVEY.md 8a "parity hazards"):
//  1. _umax: the reference's init loop (:304-322) reads an un-filled vector (UB / heap garbage); the
//     oracle uses the canonical ORB table {15,15,15,15,14,14,14,13,13,12,11,10,9,8,6,3}.
//  3. Frame::InFrame(px, 20, L) (include/ygz/Basic/Frame.h:67-71) divides LEVEL coordinates by 2^L
//     again and compares against the FULL-RES size: kept faithfully (it decides which corners exist).
//  4. IC_Angle / descriptor taps are plain pointer arithmetic on a continuous cv::Mat (:514-531,
//     :547-552): taps that leave the row wrap into the neighbouring row (deterministic, kept); taps
//     that leave the level buffer are UB in the reference and are DEFINED here as reading 0.
//  9. all f32 maths evaluated exactly as written, no FMA contraction (-ffp-contract=off).
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace {

const int8_t kOrbPattern[1024] = {
#include "../ygz_slam_b200/csrc/orb_pattern.inc"
};

const int kHalfPatch = 15;  // FeatureDetector.h:50
const int kUmax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

struct Level {
    const uint8_t* data;
    int w, h;
};

// linear-address tap (hazard 4)
inline int tap(const Level& im, long lin) {
    return (lin >= 0 && lin < (long)im.w * im.h) ? im.data[lin] : 0;
}

}  // namespace

// cvRound: round half to even (SSE cvtss2si / lrint under the default rounding mode)
extern "C" int ora_cv_round_f(float v) { return (int)lrintf(v); }
extern "C" int ora_cv_round_d(double v) { return (int)lrint(v); }

// cv::fastAtan2 (OpenCV core mathfuncs_core, scalar path), degrees, f32 polynomial
extern "C" float ora_fast_atan2(float y, float x) {
    const float p1 = 0.9997878412794807f * (float)(180 / 3.14159265358979323846);
    const float p3 = -0.3258083974640975f * (float)(180 / 3.14159265358979323846);
    const float p5 = 0.1555786518463281f * (float)(180 / 3.14159265358979323846);
    const float p7 = -0.04432655554792128f * (float)(180 / 3.14159265358979323846);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = std::fabs(x), ay = std::fabs(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// FeatureDetector::ShiTomasiScore (:467-507): 8x8 box [u-4,u+4) x [v-4,v+4), central differences,
// f32 sums (exact, < 2^24), smaller eigenvalue of the 2x2 structure tensor.
extern "C" float ora_shi_tomasi(const uint8_t* img, int w, int h, int u, int v) {
    float dXX = 0.0f, dYY = 0.0f, dXY = 0.0f;
    const int half = 4, box = 8, area = 64;
    const int x_min = u - half, x_max = u + half, y_min = v - half, y_max = v + half;
    if (x_min < 1 || x_max >= w - 1 || y_min < 1 || y_max >= h - 1) return 0.0f;
    for (int y = y_min; y < y_max; ++y) {
        const uint8_t* r = img + (size_t)y * w;
        for (int x = 0; x < box; ++x) {
            const float dx = (float)(r[x_min + x + 1] - r[x_min + x - 1]);
            const float dy = (float)(r[x_min + x + w] - r[x_min + x - w]);
            dXX += dx * dx;
            dYY += dy * dy;
            dXY += dx * dy;
        }
    }
    dXX = (float)(dXX / (2.0 * area));
    dYY = (float)(dYY / (2.0 * area));
    dXY = (float)(dXY / (2.0 * area));
    // `sqrt` on a float under `using namespace std` (Common.h:16) resolves to the f32 overload
    const float s = dXX + dYY;
    const float disc = s * s - 4 * (dXX * dYY - dXY * dXY);
    return (float)(0.5 * (s - std::sqrt(disc)));
}

namespace {

// FeatureDetector::IC_Angle (:509-537)
float ic_angle(const Level& im, double ptx, double pty) {
    int m_01 = 0, m_10 = 0;
    const long c = (long)ora_cv_round_d(pty) * im.w + ora_cv_round_d(ptx);
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m_10 += u * tap(im, c + u);
    const int step = im.w;
    for (int v = 1; v <= kHalfPatch; ++v) {
        int v_sum = 0;
        const int d = kUmax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = tap(im, c + u + (long)v * step), val_minus = tap(im, c + u - (long)v * step);
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return ora_fast_atan2((float)m_01, (float)m_10);
}

// FeatureDetector::ComputeOrbDescriptor (:539-578)
void orb_descriptor(const Level& im, double px, double py, int level, float angle_deg, uint8_t* desc) {
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angle_deg * factorPI;
    // Hazard 11: the reference calls libm cosf/sinf (unqualified cos/sin on a float under `using
    // namespace std`), whose last bit depends on the glibc version (glibc 2.39 cosf differs from the
    // correctly rounded value for 1.3 % of arguments, measured).  The oracle pins the CORRECTLY ROUNDED
    // f32 cosine/sine (f64 libm result rounded once), which is what every libm approximates.
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const int scale = 1 << level;
    const long c = (long)ora_cv_round_d(py / scale) * im.w + ora_cv_round_d(px / scale);
    const int step = im.w;
    const int8_t* pat = kOrbPattern;
    for (int i = 0; i < 32; ++i, pat += 32) {  // 16 cv::Point = 8 tests per byte
        int val = 0;
        for (int k = 0; k < 8; ++k) {
            const int x0 = pat[4 * k], y0 = pat[4 * k + 1], x1 = pat[4 * k + 2], y1 = pat[4 * k + 3];
            const int t0 = tap(im, c + (long)ora_cv_round_f(x0 * b + y0 * a) * step + ora_cv_round_f(x0 * a - y0 * b));
            const int t1 = tap(im, c + (long)ora_cv_round_f(x1 * b + y1 * a) * step + ora_cv_round_f(x1 * a - y1 * b));
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

}  // namespace

extern "C" void ora_describe(const uint8_t* pyr, int w, int h, int n_levels, int n, const double* px, const double* py,
                             const int32_t* level, float* angle, uint8_t* desc) {
    int lw[ORA_MAX_LEVELS], lh[ORA_MAX_LEVELS];
    size_t off[ORA_MAX_LEVELS];
    ora_pyramid_layout(w, h, n_levels, lw, lh, off);
    for (int i = 0; i < n; ++i) {
        const int L = level[i];
        Level im{pyr + off[L], lw[L], lh[L]};
        // ComputeAngleAndDescriptor (:580-588): IC_Angle(pyramid[L], pixel/(1<<L)) then the descriptor
        angle[i] = ic_angle(im, px[i] / (1 << L), py[i] / (1 << L));
        orb_descriptor(im, px[i], py[i], L, angle[i], desc + 32 * (size_t)i);
    }
}

// FeatureDetector::Detect (:345-444) with LoadParams() applied (:331-340)
extern "C" int ora_detect(const uint8_t* pyr, const ora_detect_params* p, const uint8_t* occupied, ora_features* out) {
    const int W = p->image_width, H = p->image_height;
    const int grid_cols = (int)std::ceil((double)W / p->cell_size);
    const int grid_rows = (int)std::ceil((double)H / p->cell_size);
    const int n_cells = grid_cols * grid_rows;
    int lw[ORA_MAX_LEVELS], lh[ORA_MAX_LEVELS];
    size_t off[ORA_MAX_LEVELS];
    ora_pyramid_layout(W, H, p->n_levels, lw, lh, off);

    struct Cand {
        bool set;
        int x, y, level;
        float score;
    };
    std::vector<Cand> cells(n_cells, Cand{false, 0, 0, 0, 0.f});
    std::vector<int16_t> xy;
    std::vector<int32_t> scores, keep;

    for (int L = 0; L < p->n_levels; ++L) {
        const int scale = 1 << L;
        const uint8_t* img = pyr + off[L];
        const int w = lw[L], h = lh[L];
        const int cap = w * h;
        xy.resize(2 * (size_t)cap);
        const int nc = ora_fast10_detect(img, w, h, w, p->threshold, xy.data(), cap);
        scores.resize(nc);
        keep.resize(nc);
        ora_fast10_score(img, w, xy.data(), nc, p->threshold, scores.data());
        const int nk = ora_fast_nonmax_3x3(xy.data(), scores.data(), nc, keep.data());
        for (int t = 0; t < nk; ++t) {
            const int x = xy[2 * keep[t]], y = xy[2 * keep[t] + 1];
            // frame->InFrame(Vector2d(x,y), 20, L): (Frame.h:67-71) double-scaled, vs full-res size
            const double xs = (double)x / scale, ys = (double)y / scale;
            if (!(xs >= 20 && xs < W - 20 && ys >= 20 && ys < H - 20)) continue;
            const int gy = (y * scale) / p->cell_size, gx = (x * scale) / p->cell_size;
            const size_t k = (size_t)gy * grid_cols + gx;
            if (k > (size_t)n_cells) continue;  // sic (:394)
            if (k == (size_t)n_cells) continue; // the reference would index out of range here; unreachable for 640x480
            if (occupied && occupied[k]) continue;
            const float s = ora_shi_tomasi(img, w, h, x, y);
            if (cells[k].set) {
                if (s > cells[k].score) cells[k] = Cand{true, x, y, L, s};
            } else {
                cells[k] = Cand{true, x, y, L, s};
            }
        }
    }
    int n = 0;
    for (int k = 0; k < n_cells; ++k) {
        if (!cells[k].set) continue;
        const Cand& c = cells[k];
        const int scale = 1 << c.level;
        out->px[n] = (double)(c.x * scale);
        out->py[n] = (double)(c.y * scale);
        out->level[n] = c.level;
        out->score[n] = c.score;
        if (out->cell) out->cell[n] = k;
        Level im{pyr + off[c.level], lw[c.level], lh[c.level]};
        out->angle[n] = ic_angle(im, out->px[n] / scale, out->py[n] / scale);
        orb_descriptor(im, out->px[n], out->py[n], c.level, out->angle[n], out->desc + 32 * (size_t)n);
        ++n;
    }
    out->n = n;
    return n;
}
