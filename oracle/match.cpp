// oracle/match.cpp -- descriptor distance and brute-force matching.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
//   Matcher::DescriptorDistance      reference src/Algorithm/Matcher.cpp:30-43 (8 x 32-bit SWAR popcount)
//   Matcher::CheckFrameDescriptors   reference src/Algorithm/Matcher.cpp:45-84
//   cv::BFMatcher(NORM_HAMMING, crossCheck=true).match + "good match" filter
//                                    reference test/test_orb_match.cpp:86-105  (OpenCV-owned; pinned
//                                    against cv2.BFMatcher in tests/test_oracle_cv2.py)
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <vector>

extern "C" int ora_descriptor_distance(const uint8_t* a, const uint8_t* b) {
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t wa, wb;
        std::memcpy(&wa, a + 4 * i, 4);
        std::memcpy(&wb, b + 4 * i, 4);
        uint32_t v = wa ^ wb;
        v = v - ((v >> 1) & 0x55555555u);
        v = (v & 0x33333333u) + ((v >> 2) & 0x33333333u);
        dist += (int)((((v + (v >> 4)) & 0xF0F0F0Fu) * 0x1010101u) >> 24);
    }
    return dist;
}

// BFMatcher::match -> knnMatch(k=1): per query the FIRST minimum over the train rows.
// crossCheck: (i,j) kept iff i is also the first minimum of train row j over the query rows.
namespace {
// Hamming distance of two 256-bit descriptors for the brute-force matcher: cv::BFMatcher uses OpenCV's normHamming (hardware
// POPCNT where available), so the timed CPU-baseline build does too; the value is identical to DescriptorDistance
inline int bf_distance(const uint8_t* a, const uint8_t* b) {
#if defined(__POPCNT__)
    int d = 0;
    for (int i = 0; i < 4; ++i) {
        uint64_t wa, wb;
        std::memcpy(&wa, a + 8 * i, 8);
        std::memcpy(&wb, b + 8 * i, 8);
        d += __builtin_popcountll(wa ^ wb);
    }
    return d;
#else
    return ora_descriptor_distance(a, b);
#endif
}
}  // namespace

extern "C" void ora_match_bf(const uint8_t* A, int nA, const uint8_t* B, int nB, int cross_check, int32_t* train_idx,
                             int32_t* dist) {
    std::vector<int32_t> best_q(nB, -1), best_qd(nB, 1 << 30);
    for (int i = 0; i < nA; ++i) {
        int bj = -1, bd = 1 << 30;
        for (int j = 0; j < nB; ++j) {
            const int d = bf_distance(A + 32 * (size_t)i, B + 32 * (size_t)j);
            if (d < bd) {
                bd = d;
                bj = j;
            }
            if (d < best_qd[j]) {
                best_qd[j] = d;
                best_q[j] = i;
            }
        }
        train_idx[i] = bj;
        dist[i] = bj >= 0 ? bd : -1;
    }
    if (cross_check)
        for (int i = 0; i < nA; ++i)
            if (train_idx[i] >= 0 && best_q[train_idx[i]] != i) {
                train_idx[i] = -1;
                dist[i] = -1;
            }
}

// test_orb_match.cpp:97-105
extern "C" int ora_good_matches(const int32_t* train_idx, const int32_t* dist, int nA, uint8_t* keep) {
    double min_dis = 1e30;
    for (int i = 0; i < nA; ++i)
        if (train_idx[i] >= 0 && dist[i] < min_dis) min_dis = dist[i];
    min_dis = min_dis < 20 ? 20 : min_dis;
    min_dis = min_dis > 50 ? 50 : min_dis;
    int n = 0;
    for (int i = 0; i < nA; ++i) {
        keep[i] = (train_idx[i] >= 0 && dist[i] < 3 * min_dis) ? 1 : 0;
        n += keep[i];
    }
    return n;
}

// Matcher::CheckFrameDescriptors (:45-84): initMatchRatio = 3.0 (Matcher.h Option)
extern "C" int ora_check_descriptors(const uint8_t* A, const uint8_t* B, const int32_t* ia, const int32_t* ib, int n,
                                     int init_low, int init_high, int32_t* dist, uint8_t* keep) {
    if (n <= 0) return 0;
    int best = 1 << 30;
    for (int k = 0; k < n; ++k) {
        dist[k] = ora_descriptor_distance(A + 32 * (size_t)ia[k], B + 32 * (size_t)ib[k]);
        if (dist[k] < best) best = dist[k];
    }
    best = best > init_low ? best : init_low;
    best = best < init_high ? best : init_high;
    int good = 0;
    for (int k = 0; k < n; ++k) {
        keep[k] = (dist[k] < 3.0 * best) ? 1 : 0;
        good += keep[k];
    }
    return good;
}

// ---- matching for triangulation -------------------------------------------------------------
//   Matcher::SearchForTriangulation   reference src/Algorithm/Matcher.cpp:86-193
//   Matcher::CheckDistEpipolarLine    reference src/Algorithm/Matcher.cpp:338-354
// The DBoW3 feature vectors arrive as one node id per feature (-1 = in no node); index lists of a node are ascending.
namespace {
inline bool check_dist_epipolar_line(double x1, double y1, double x2, double y2, const double* E, float dsqr_th) {
    const float a = (float)(x1 * E[0] + y1 * E[3] + E[6]);
    const float b = (float)(x1 * E[1] + y1 * E[4] + E[7]);
    const float c = (float)(x1 * E[2] + y1 * E[5] + E[8]);
    const float num = (float)(a * x2 + b * y2 + c);   // float * double -> double, one rounding on assignment
    const float den = a * a + b * b;
    if (den < 1e-6) return false;
    const float dsqr = num * num / den;
    return std::fabs(dsqr) < dsqr_th;
}
}  // namespace

extern "C" void ora_search_for_triangulation(const ora_camera* cam, int n1, const uint8_t* desc1, const double* px1, const int32_t* node1,
                                             int n2, const uint8_t* desc2, const double* px2, const int32_t* node2, const double* E12,
                                             int th_low, double epipolar_dsqr, int32_t* match12) {
    const float fx = cam->fx, fy = cam->fy, cx = cam->cx, cy = cam->cy;
    for (int i = 0; i < n1; ++i) {
        match12[i] = -1;
        if (node1[i] < 0) continue;
        const double x1 = (px1[2 * i] - cx) * 1.0 / fx, y1 = (px1[2 * i + 1] - cy) * 1.0 / fy;   // PinholeCamera::Pixel2Camera
        int bestDist = 256, bestIdx2 = -1;
        for (int j = 0; j < n2; ++j) {
            if (node2[j] != node1[i]) continue;
            const int dist = ora_descriptor_distance(desc1 + 32 * (size_t)i, desc2 + 32 * (size_t)j);
            if (dist > th_low || dist > bestDist) continue;
            const double x2 = (px2[2 * j] - cx) * 1.0 / fx, y2 = (px2[2 * j + 1] - cy) * 1.0 / fy;
            if (check_dist_epipolar_line(x1, y1, x2, y2, E12, (float)epipolar_dsqr)) {
                bestIdx2 = j;
                bestDist = dist;
            }
        }
        match12[i] = bestIdx2;
    }
}

// cvutils::DepthFromTriangulation (include/ygz/Algorithm/CVUtils.h:18-38); T = 3x4 [R|t] of T_search_ref
extern "C" int ora_depth_from_triangulation(const double* T, const double* f_ref, const double* f_cur, double determinant_th,
                                            double* depth1, double* depth2) {
    double a0[3], a1[3];
    for (int r = 0; r < 3; ++r) {
        a0[r] = T[4 * r] * f_ref[0] + T[4 * r + 1] * f_ref[1] + T[4 * r + 2] * f_ref[2];
        a1[r] = -f_cur[r];
    }
    const double m00 = a0[0] * a0[0] + a0[1] * a0[1] + a0[2] * a0[2], m01 = a0[0] * a1[0] + a0[1] * a1[1] + a0[2] * a1[2],
                 m11 = a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2];
    const double det = m00 * m11 - m01 * m01;
    if (det < determinant_th) return 0;
    const double b0 = a0[0] * T[3] + a0[1] * T[7] + a0[2] * T[11], b1 = a1[0] * T[3] + a1[1] * T[7] + a1[2] * T[11];
    const double id = 1.0 / det;
    *depth1 = std::fabs(-(m11 * id * b0 + -m01 * id * b1));
    *depth2 = std::fabs(-(-m01 * id * b0 + m00 * id * b1));
    return 1;
}
