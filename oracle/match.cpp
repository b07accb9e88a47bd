// oracle/match.cpp -- descriptor distance and brute-force matching.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
//   Matcher::DescriptorDistance      reference src/Algorithm/Matcher.cpp:30-43 (8 x 32-bit SWAR popcount)
//   Matcher::CheckFrameDescriptors   reference src/Algorithm/Matcher.cpp:45-84
//   cv::BFMatcher(NORM_HAMMING, crossCheck=true).match + "good match" filter
//                                    reference test/test_orb_match.cpp:86-105  (OpenCV-owned; pinned
//                                    against cv2.BFMatcher in tests/test_oracle_cv2.py)
#include "oracle.h"

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
