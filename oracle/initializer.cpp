// oracle/initializer.cpp -- CPU restatement of the RANSAC half of the monocular initialiser (TEST INFRASTRUCTURE ONLY, see
// oracle.h):
//   Initializer::TryInitialize, RANSAC sets   reference src/Algorithm/Initializer.cpp:9-87   (cv::RNG minimal sets :33-49, model choice :66-78)
//   Initializer::FindHomography                reference src/Algorithm/Initializer.cpp:89-138
//   Initializer::Normalize                     reference src/Algorithm/Initializer.cpp:140-175
//   Initializer::ComputeH21                    reference src/Algorithm/Initializer.cpp:196-239
//   Initializer::CheckHomography               reference src/Algorithm/Initializer.cpp:251-318
//   Initializer::FindFundamental               reference src/Algorithm/Initializer.cpp:670-717
//   Initializer::ComputeF21                    reference src/Algorithm/Initializer.cpp:730-762
//   Initializer::CheckFundamental              reference src/Algorithm/Initializer.cpp:772-853
// Third-party arithmetic that is not in the reference tree, restated:
//   * cv::RNG (OpenCV core/operations.hpp): multiply-with-carry, state = (uint64)(unsigned)state * 4164903690 + (state >> 32),
//     default state 0xffffffff, uniform(a, b) = a + next() % (b - a).  Not reachable through cv2's Python API: unpinned.
//   * Eigen::JacobiSVD (ComputeH21 / ComputeF21 take the right singular vector of the smallest singular value; ComputeF21 zeroes
//     the smallest singular value of the 3 x 3 matrix): restated as a one-sided (Hestenes) Jacobi SVD -- same singular
//     subspaces to rounding.  A singular vector is defined up to sign; H and F are homogeneous and every use below is
//     sign-invariant.  The null vector is canonicalised to a non-negative largest-magnitude component.
//   * Eigen 3 x 3 inverse: cofactors / determinant.
// Reference defects kept or resolved: Normalize() accumulates into uninitialised `Vector2d mean, meanDev` (:146, :153) --
// zero-initialised here; CheckHomography scores only the transfer of image 2 into image 1 (:284-303) -- kept; a RANSAC run
// whose best score stays 0 leaves H21 / F21 untouched in the reference -- reported as best = -1 and a zero matrix here.
// The CUDA path (ygz_slam_b200/csrc/initializer.cu) performs the same operations in the same order without FMA contraction,
// so the comparison in tests/test_initializer.py is bit for bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

namespace {

// one-sided Jacobi SVD of the M x N matrix a (row major, overwritten: columns become U * sigma); v = N x N right vectors
template <int M, int N>
void jacobi_svd(double* a, double* v) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) v[i * N + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 40; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                double alpha = 0, beta = 0, gamma = 0;
                for (int r = 0; r < M; ++r) {
                    const double x = a[r * N + p], y = a[r * N + q];
                    alpha += x * x;
                    beta += y * y;
                    gamma += x * y;
                }
                if (gamma == 0.0 || std::fabs(gamma) <= 1e-15 * std::sqrt(alpha * beta)) continue;
                rotated = true;
                const double zeta = (beta - alpha) / (2.0 * gamma);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < M; ++r) {
                    const double x = a[r * N + p], y = a[r * N + q];
                    a[r * N + p] = c * x - s * y;
                    a[r * N + q] = s * x + c * y;
                }
                for (int r = 0; r < N; ++r) {
                    const double x = v[r * N + p], y = v[r * N + q];
                    v[r * N + p] = c * x - s * y;
                    v[r * N + q] = s * x + c * y;
                }
            }
        if (!rotated) break;
    }
}

template <int M, int N>
int smallest_column(const double* a, double* norms2) {
    int best = 0;
    for (int c = 0; c < N; ++c) {
        double n2 = 0;
        for (int r = 0; r < M; ++r) n2 += a[r * N + c] * a[r * N + c];
        norms2[c] = n2;
        if (n2 < norms2[best]) best = c;
    }
    return best;
}

// right singular vector of the smallest singular value, largest-magnitude component made non-negative
template <int M>
void null_vector(double* a, double* out9) {
    double v[81], n2[9];
    jacobi_svd<M, 9>(a, v);
    const int c = smallest_column<M, 9>(a, n2);
    int big = 0;
    for (int r = 1; r < 9; ++r)
        if (std::fabs(v[r * 9 + c]) > std::fabs(v[big * 9 + c])) big = r;
    const double sg = v[big * 9 + c] < 0 ? -1.0 : 1.0;
    for (int r = 0; r < 9; ++r) out9[r] = sg * v[r * 9 + c];
}

void mul3(const double* A, const double* B, double* C) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C[r * 3 + c] = A[r * 3] * B[c] + A[r * 3 + 1] * B[3 + c] + A[r * 3 + 2] * B[6 + c];
}

void inverse3(const double* m, double* inv) {
    const double c00 = m[4] * m[8] - m[5] * m[7], c01 = m[5] * m[6] - m[3] * m[8], c02 = m[3] * m[7] - m[4] * m[6];
    const double det = m[0] * c00 + m[1] * c01 + m[2] * c02;
    const double id = 1.0 / det;
    inv[0] = c00 * id;
    inv[1] = (m[2] * m[7] - m[1] * m[8]) * id;
    inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    inv[3] = c01 * id;
    inv[4] = (m[0] * m[8] - m[2] * m[6]) * id;
    inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    inv[6] = c02 * id;
    inv[7] = (m[1] * m[6] - m[0] * m[7]) * id;
    inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

// Initializer::Normalize (:140-175); T row major
void normalize(int n, const double* px, std::vector<double>& pn, double* T) {
    double mean[2] = {0, 0}, dev[2] = {0, 0};
    for (int i = 0; i < n; ++i) {
        mean[0] += px[2 * i];
        mean[1] += px[2 * i + 1];
    }
    mean[0] = mean[0] / n;
    mean[1] = mean[1] / n;
    pn.resize(2 * (size_t)n);
    for (int i = 0; i < n; ++i) {
        pn[2 * i] = px[2 * i] - mean[0];
        pn[2 * i + 1] = px[2 * i + 1] - mean[1];
        dev[0] += std::fabs(pn[2 * i]);
        dev[1] += std::fabs(pn[2 * i + 1]);
    }
    dev[0] /= n;
    dev[1] /= n;
    const float sX = (float)(1.0 / dev[0]), sY = (float)(1.0 / dev[1]);
    for (int i = 0; i < n; ++i) {
        pn[2 * i] *= sX;
        pn[2 * i + 1] *= sY;
    }
    const double t[9] = {sX, 0, -mean[0] * sX, 0, sY, -mean[1] * sY, 0, 0, 1};
    std::memcpy(T, t, sizeof t);
}

// CheckHomography (:251-318): score, optionally the inlier flags
float check_homography(int n, const double* px1, const double* px2, const double* H12, float sigma, uint8_t* inl) {
    float score = 0;
    const float th = 5.991f;
    const float invSigmaSquare = (float)(1.0 / (sigma * sigma));
    for (int i = 0; i < n; ++i) {
        const double u1 = px1[2 * i], v1 = px1[2 * i + 1], u2 = px2[2 * i], v2 = px2[2 * i + 1];
        const float w2in1inv = (float)(1.0 / (H12[6] * u2 + H12[7] * v2 + H12[8]));
        const float u2in1 = (float)((H12[0] * u2 + H12[1] * v2 + H12[2]) * w2in1inv);
        const float v2in1 = (float)((H12[3] * u2 + H12[4] * v2 + H12[5]) * w2in1inv);
        const float squareDist1 = (float)((u1 - u2in1) * (u1 - u2in1) + (v1 - v2in1) * (v1 - v2in1));
        const float chiSquare1 = squareDist1 * invSigmaSquare;
        const bool in = !(chiSquare1 > th);
        if (in) score += th - chiSquare1;
        if (inl) inl[i] = in ? 1 : 0;
    }
    return score;
}

// CheckFundamental (:772-853)
float check_fundamental(int n, const double* px1, const double* px2, const double* F21, float sigma, uint8_t* inl) {
    const float f11 = (float)F21[0], f12 = (float)F21[1], f13 = (float)F21[2], f21 = (float)F21[3], f22 = (float)F21[4], f23 = (float)F21[5],
                f31 = (float)F21[6], f32 = (float)F21[7], f33 = (float)F21[8];
    float score = 0;
    const float th = 3.841f, thScore = 5.991f;
    const float invSigmaSquare = (float)(1.0 / (sigma * sigma));
    for (int i = 0; i < n; ++i) {
        bool bIn = true;
        const float u1 = (float)px1[2 * i], v1 = (float)px1[2 * i + 1], u2 = (float)px2[2 * i], v2 = (float)px2[2 * i + 1];
        const float a2 = f11 * u1 + f12 * v1 + f13, b2 = f21 * u1 + f22 * v1 + f23, c2 = f31 * u1 + f32 * v1 + f33;
        const float num2 = a2 * u2 + b2 * v2 + c2;
        const float squareDist1 = num2 * num2 / (a2 * a2 + b2 * b2);
        const float chiSquare1 = squareDist1 * invSigmaSquare;
        if (chiSquare1 > th) bIn = false;
        else score += thScore - chiSquare1;
        const float a1 = f11 * u2 + f21 * v2 + f31, b1 = f12 * u2 + f22 * v2 + f32, c1 = f13 * u2 + f23 * v2 + f33;
        const float num1 = a1 * u1 + b1 * v1 + c1;
        const float squareDist2 = num1 * num1 / (a1 * a1 + b1 * b1);
        const float chiSquare2 = squareDist2 * invSigmaSquare;
        if (chiSquare2 > th) bIn = false;
        else score += thScore - chiSquare2;
        if (inl) inl[i] = bIn ? 1 : 0;
    }
    return score;
}

}  // namespace

// the RANSAC minimal sets of TryInitialize (:25-49): sets[it * 8 + j]; a fresh cv::RNG per call, like the reference
extern "C" void ora_initializer_sets(int n_points, int max_iter, int32_t* sets) {
    uint64_t state = 0xffffffffULL;
    auto next = [&]() {
        state = (uint64_t)(uint32_t)state * 4164903690ULL + (uint32_t)(state >> 32);
        return (uint32_t)state;
    };
    std::vector<int32_t> avail;
    for (int it = 0; it < max_iter; ++it) {
        avail.resize(n_points);
        for (int i = 0; i < n_points; ++i) avail[i] = i;
        for (int j = 0; j < 8; ++j) {
            const int b = (int)avail.size();
            const int r = b == 0 ? 0 : (int)(next() % (uint32_t)b);
            sets[it * 8 + j] = avail[r];
            avail[r] = avail.back();
            avail.pop_back();
        }
    }
}

// FindHomography + FindFundamental for one pair of point lists (n >= 8).  H21 / F21: 9 doubles row major, score_*: the
// float scores (sh, sf of TryInitialize), best_*: the winning iteration or -1, inl_*: n flags, models (may be NULL):
// max_iter x 18 doubles, every iteration's H21i then F21i (for the tests)
extern "C" void ora_initializer_ransac(int n, const double* px1, const double* px2, int max_iter, const int32_t* sets, float sigma, double* H21,
                                       float* score_H, int32_t* best_H, uint8_t* inl_H, double* F21, float* score_F, int32_t* best_F,
                                       uint8_t* inl_F, double* models) {
    std::vector<double> pn1, pn2;
    double T1[9], T2[9], T2inv[9], T2t[9];
    normalize(n, px1, pn1, T1);
    normalize(n, px2, pn2, T2);
    inverse3(T2, T2inv);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T2t[r * 3 + c] = T2[c * 3 + r];
    *score_H = 0;
    *score_F = 0;
    *best_H = -1;
    *best_F = -1;
    std::memset(H21, 0, 72);
    std::memset(F21, 0, 72);
    std::memset(inl_H, 0, n);
    std::memset(inl_F, 0, n);
    for (int it = 0; it < max_iter; ++it) {
        // ---- homography: ComputeH21 (:196-239)
        double A[16 * 9];
        for (int j = 0; j < 8; ++j) {
            const int idx = sets[it * 8 + j];
            const double u1 = pn1[2 * idx], v1 = pn1[2 * idx + 1], u2 = pn2[2 * idx], v2 = pn2[2 * idx + 1];
            double* r0 = A + (2 * j) * 9;
            double* r1 = r0 + 9;
            r0[0] = 0.0; r0[1] = 0.0; r0[2] = 0.0; r0[3] = -u1; r0[4] = -v1; r0[5] = -1; r0[6] = v2 * u1; r0[7] = v2 * v1; r0[8] = v2;
            r1[0] = u1; r1[1] = v1; r1[2] = 1; r1[3] = 0.0; r1[4] = 0.0; r1[5] = 0.0; r1[6] = -u2 * u1; r1[7] = -u2 * v1; r1[8] = -u2;
        }
        double Hn[9], tmp[9], H21i[9], H12i[9];
        null_vector<16>(A, Hn);
        mul3(T2inv, Hn, tmp);
        mul3(tmp, T1, H21i);
        inverse3(H21i, H12i);
        const float sh = check_homography(n, px1, px2, H12i, sigma, nullptr);
        if (sh > *score_H) {
            *score_H = sh;
            *best_H = it;
            std::memcpy(H21, H21i, 72);
        }
        // ---- fundamental matrix: ComputeF21 (:730-762)
        double B[8 * 9];
        for (int j = 0; j < 8; ++j) {
            const int idx = sets[it * 8 + j];
            const double u1 = pn1[2 * idx], v1 = pn1[2 * idx + 1], u2 = pn2[2 * idx], v2 = pn2[2 * idx + 1];
            double* r = B + j * 9;
            r[0] = u2 * u1; r[1] = u2 * v1; r[2] = u2; r[3] = v2 * u1; r[4] = v2 * v1; r[5] = v2; r[6] = u1; r[7] = v1; r[8] = 1;
        }
        double Fpre[9], Fa[9], Fv[9], n2[3], Fn[9], F21i[9];
        null_vector<8>(B, Fpre);
        std::memcpy(Fa, Fpre, 72);
        jacobi_svd<3, 3>(Fa, Fv);
        const int cz = smallest_column<3, 3>(Fa, n2);
        // U diag(s0, s1, 0) V^T = Fpre - (column cz of U sigma) (column cz of V)^T
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Fn[r * 3 + c] = Fpre[r * 3 + c] - Fa[r * 3 + cz] * Fv[c * 3 + cz];
        mul3(T2t, Fn, tmp);
        mul3(tmp, T1, F21i);
        const float sf = check_fundamental(n, px1, px2, F21i, sigma, nullptr);
        if (sf > *score_F) {
            *score_F = sf;
            *best_F = it;
            std::memcpy(F21, F21i, 72);
        }
        if (models) {
            std::memcpy(models + 18 * (size_t)it, H21i, 72);
            std::memcpy(models + 18 * (size_t)it + 9, F21i, 72);
        }
    }
    if (*best_H >= 0) {
        double H12[9];
        inverse3(H21, H12);
        check_homography(n, px1, px2, H12, sigma, inl_H);
    }
    if (*best_F >= 0) check_fundamental(n, px1, px2, F21, sigma, inl_F);
}

// =====================================================================================================================
// Second half: pose and structure from the chosen model.
//   Initializer::ReconstructH   reference src/Algorithm/Initializer.cpp:330-513  (Faugeras' eight hypotheses)
//   Initializer::CheckRT        reference src/Algorithm/Initializer.cpp:515-630
//   Initializer::Triangulate    reference src/Algorithm/Initializer.cpp:661-675
//   Initializer::ReconstructF   reference src/Algorithm/Initializer.cpp:855-941
//   Initializer::DecomposeE     reference src/Algorithm/Initializer.cpp:943-963
// Eigen::JacobiSVD of the 3 x 3 / 4 x 4 matrices is again the one-sided Jacobi SVD above, singular values sorted descending,
// U = A V / sigma (the left vector of a singular value below 1e-12 sigma_max -- the essential matrix' third -- is the cross
// product of the other two, as an orthogonal U requires).  A singular pair is defined up to a common sign; the eight (H) / four
// (F) pose hypotheses are closed under those sign changes, so the selected pose does not depend on them, only the order in which
// equal candidates would be met does.  acos() of CheckRT runs in float like the reference's (std::acos(float) through
// `using namespace std`); the CUDA path's acosf may differ in the last bit: parallax is compared with a tolerance.
namespace {

struct Svd3 {
    double U[9], V[9], s[3];   // row major; columns sorted by descending singular value
};

Svd3 svd3_sorted(const double* A) {
    double a[9], v[9], n2[3];
    std::memcpy(a, A, 72);
    jacobi_svd<3, 3>(a, v);
    for (int c = 0; c < 3; ++c) n2[c] = a[c] * a[c] + a[3 + c] * a[3 + c] + a[6 + c] * a[6 + c];
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (n2[ord[j]] > n2[ord[i]]) {
                const int t = ord[i];
                ord[i] = ord[j];
                ord[j] = t;
            }
    Svd3 r;
    for (int c = 0; c < 3; ++c) {
        const int o = ord[c];
        r.s[c] = std::sqrt(n2[o]);
        for (int k = 0; k < 3; ++k) r.V[k * 3 + c] = v[k * 3 + o];
    }
    for (int c = 0; c < 3; ++c) {
        const int o = ord[c];
        if (c < 2 || r.s[2] > 1e-12 * r.s[0]) {
            for (int k = 0; k < 3; ++k) r.U[k * 3 + c] = a[k * 3 + o] / r.s[c];
        } else {
            r.U[2] = r.U[3] * r.U[7] - r.U[6] * r.U[4];
            r.U[5] = r.U[6] * r.U[1] - r.U[0] * r.U[7];
            r.U[8] = r.U[0] * r.U[4] - r.U[3] * r.U[1];
        }
    }
    return r;
}

double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

void transpose3(const double* A, double* T) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) T[r * 3 + c] = A[c * 3 + r];
}

// Initializer::Triangulate: the right singular vector of the smallest singular value of the 4 x 4 DLT matrix, dehomogenised
void triangulate(const double* kp1, const double* kp2, const double* P1, const double* P2, double* x3D) {
    double A[16], v[16], n2[4];
    for (int c = 0; c < 4; ++c) {
        A[c] = kp1[0] * P1[8 + c] - P1[c];
        A[4 + c] = kp1[1] * P1[8 + c] - P1[4 + c];
        A[8 + c] = kp2[0] * P2[8 + c] - P2[c];
        A[12 + c] = kp2[1] * P2[8 + c] - P2[4 + c];
    }
    jacobi_svd<4, 4>(A, v);
    const int c = smallest_column<4, 4>(A, n2);
    for (int k = 0; k < 3; ++k) x3D[k] = v[k * 4 + c] / v[12 + c];
}

// Initializer::CheckRT: number of points in front of both cameras (and, optionally, with a small reprojection error);
// good[i] additionally needs parallax; p3d of the counted points; parallax = the 51st smallest angle in degrees
int check_rt(int n, const double* px1, const double* px2, const double* R, const double* t, const double* K /* fx fy cx cy */, float th2,
             bool check_reprojection, uint8_t* good, double* p3d, double* parallax) {
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    const double P1[12] = {fx, 0, cx, 0, 0, fy, cy, 0, 0, 0, 1, 0};
    const double Rt[12] = {R[0], R[1], R[2], t[0], R[3], R[4], R[5], t[1], R[6], R[7], R[8], t[2]};
    double P2[12];
    for (int c = 0; c < 4; ++c) {   // K * [R | t]
        P2[c] = fx * Rt[c] + 0.0 * Rt[4 + c] + cx * Rt[8 + c];
        P2[4 + c] = 0.0 * Rt[c] + fy * Rt[4 + c] + cy * Rt[8 + c];
        P2[8 + c] = 0.0 * Rt[c] + 0.0 * Rt[4 + c] + 1.0 * Rt[8 + c];
    }
    const double O2[3] = {-(R[0] * t[0] + R[3] * t[1] + R[6] * t[2]), -(R[1] * t[0] + R[4] * t[1] + R[7] * t[2]),
                          -(R[2] * t[0] + R[5] * t[1] + R[8] * t[2])};
    std::vector<float> cosv;
    cosv.reserve(n);
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        good[i] = 0;
        p3d[3 * i] = p3d[3 * i + 1] = p3d[3 * i + 2] = 0;
        double X[3];
        triangulate(px1 + 2 * i, px2 + 2 * i, P1, P2, X);
        if (!std::isfinite(X[0])) continue;
        const double dist1 = std::sqrt(X[0] * X[0] + X[1] * X[1] + X[2] * X[2]);
        const double n2v[3] = {X[0] - O2[0], X[1] - O2[1], X[2] - O2[2]};
        const double dist2 = std::sqrt(n2v[0] * n2v[0] + n2v[1] * n2v[1] + n2v[2] * n2v[2]);
        const double cosParallax = (X[0] * n2v[0] + X[1] * n2v[1] + X[2] * n2v[2]) / (dist1 * dist2);
        if (X[2] < 0 && cosParallax < 0.99998) continue;
        const double Y[3] = {R[0] * X[0] + R[1] * X[1] + R[2] * X[2] + t[0], R[3] * X[0] + R[4] * X[1] + R[5] * X[2] + t[1],
                             R[6] * X[0] + R[7] * X[1] + R[8] * X[2] + t[2]};
        if (Y[2] < 0 && cosParallax < 0.99998) continue;
        if (check_reprojection) {
            const double invZ1 = 1.0 / X[2];
            const double im1x = fx * X[0] * invZ1 + cx, im1y = fy * X[1] * invZ1 + cy;
            const double e1 = (im1x - px1[2 * i]) * (im1x - px1[2 * i]) + (im1y - px1[2 * i + 1]) * (im1y - px1[2 * i + 1]);
            if (e1 > th2) continue;
            const double invZ2 = 1.0 / Y[2];
            const double im2x = fx * Y[0] * invZ2 + cx, im2y = fy * Y[1] * invZ2 + cy;
            const double e2 = (im2x - px2[2 * i]) * (im2x - px2[2 * i]) + (im2y - px2[2 * i + 1]) * (im2y - px2[2 * i + 1]);
            if (e2 > th2) continue;
        }
        cosv.push_back((float)cosParallax);
        p3d[3 * i] = X[0];
        p3d[3 * i + 1] = X[1];
        p3d[3 * i + 2] = X[2];
        ++cnt;
        if (cosParallax < 0.99998) good[i] = 1;
    }
    if (cnt > 0) {
        std::sort(cosv.begin(), cosv.end());
        const size_t idx = (size_t)std::min(50, (int)cosv.size() - 1);
        *parallax = std::acos(cosv[idx]) * 180 / M_PI;   // float acos, like std::acos(float) in the reference
    } else {
        *parallax = 0;
    }
    return cnt;
}

void mat_R(const double* U, double s, const double* Rp, const double* V, double* R) {   // (s U) Rp V^T
    double sU[9], t1[9], Vt[9];
    for (int k = 0; k < 9; ++k) sU[k] = s * U[k];
    mul3(sU, Rp, t1);
    transpose3(V, Vt);
    mul3(t1, Vt, R);
}

}  // namespace

// ReconstructH (use_h != 0) or ReconstructF on the model chosen by TryInitialize.  K = {fx, fy, cx, cy}; inliers = the model's
// flags (only their count enters, ReconstructF :861-864, 882).  Out: R21 (row major), t21, p3d (3 n, zero where not
// triangulated), triangulated flags, info[8] = {candidates' good counts ...}; parallax of the selected candidate.  Returns the
// function's bool.
extern "C" int ora_initializer_reconstruct(int n, const double* px1, const double* px2, int use_h, const double* model, const uint8_t* inliers,
                                           const double* K, float sigma2, float min_parallax, int min_triangulated, double ratio_h,
                                           double* R21, double* t21, double* p3d, uint8_t* triangulated, int32_t* n_good, double* parallax_out,
                                           double* candidates /* may be NULL: 8 x 12, every candidate's R (9) and t (3) */) {
    int N = 0;
    for (int i = 0; i < n; ++i) N += inliers[i] ? 1 : 0;
    const double Km[9] = {K[0], 0, K[2], 0, K[1], K[3], 0, 0, 1};
    std::memset(R21, 0, 72);
    std::memset(t21, 0, 24);
    std::memset(p3d, 0, 24 * (size_t)n);
    std::memset(triangulated, 0, n);
    for (int k = 0; k < 8; ++k) n_good[k] = 0;
    *parallax_out = 0;
    if (candidates) std::memset(candidates, 0, 96 * sizeof(double));
    std::vector<uint8_t> good(n);
    std::vector<double> pts(3 * (size_t)n);
    if (use_h) {
        double invK[9], t1[9], A[9];
        inverse3(Km, invK);
        mul3(invK, model, t1);
        mul3(t1, Km, A);
        const Svd3 sv = svd3_sorted(A);
        const double d1 = sv.s[0], d2 = sv.s[1], d3 = sv.s[2];
        const double s = det3(sv.U) * det3(sv.V);
        if (d1 / d2 < 1.00001 || d2 / d3 < 1.00001) return 0;
        double Rs[8][9], ts[8][3];
        const float aux1 = (float)std::sqrt((d1 * d1 - d2 * d2) / (d1 * d1 - d3 * d3));
        const float aux3 = (float)std::sqrt((d2 * d2 - d3 * d3) / (d1 * d1 - d3 * d3));
        const float x1[4] = {aux1, aux1, -aux1, -aux1}, x3[4] = {aux3, -aux3, aux3, -aux3};
        const float aux_stheta = (float)(std::sqrt((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 + d3) * d2));
        const float ctheta = (float)((d2 * d2 + d1 * d3) / ((d1 + d3) * d2));
        const float stheta[4] = {aux_stheta, -aux_stheta, -aux_stheta, aux_stheta};
        for (int i = 0; i < 4; ++i) {
            const double Rp[9] = {ctheta, 0, -stheta[i], 0, 1, 0, stheta[i], 0, ctheta};
            mat_R(sv.U, s, Rp, sv.V, Rs[i]);
            const double tp[3] = {x1[i] * (d1 - d3), 0.0 * (d1 - d3), -x3[i] * (d1 - d3)};
            double tt[3];
            for (int r = 0; r < 3; ++r) tt[r] = sv.U[r * 3] * tp[0] + sv.U[r * 3 + 1] * tp[1] + sv.U[r * 3 + 2] * tp[2];
            const double nn = std::sqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
            for (int r = 0; r < 3; ++r) ts[i][r] = tt[r] / nn;
        }
        const float aux_sphi = (float)(std::sqrt((d1 * d1 - d2 * d2) * (d2 * d2 - d3 * d3)) / ((d1 - d3) * d2));
        const float cphi = (float)((d1 * d3 - d2 * d2) / ((d1 - d3) * d2));
        const float sphi[4] = {aux_sphi, -aux_sphi, -aux_sphi, aux_sphi};
        for (int i = 0; i < 4; ++i) {
            const double Rp[9] = {cphi, 0, sphi[i], 0, -1, 0, sphi[i], 0, -cphi};
            mat_R(sv.U, s, Rp, sv.V, Rs[4 + i]);
            const double tp[3] = {x1[i] * (d1 + d3), 0.0 * (d1 + d3), x3[i] * (d1 + d3)};
            double tt[3];
            for (int r = 0; r < 3; ++r) tt[r] = sv.U[r * 3] * tp[0] + sv.U[r * 3 + 1] * tp[1] + sv.U[r * 3 + 2] * tp[2];
            const double nn = std::sqrt(tt[0] * tt[0] + tt[1] * tt[1] + tt[2] * tt[2]);
            for (int r = 0; r < 3; ++r) ts[4 + i][r] = tt[r] / nn;
        }
        if (candidates)
            for (int i = 0; i < 8; ++i) {
                std::memcpy(candidates + 12 * i, Rs[i], 72);
                std::memcpy(candidates + 12 * i + 9, ts[i], 24);
            }
        int bestGood = 0, secondBestGood = 0, bestIdx = -1;
        float bestParallax = -1;
        for (int i = 0; i < 8; ++i) {
            double par = 0;
            const int nGood = check_rt(n, px1, px2, Rs[i], ts[i], K, 4.0f * sigma2, true, good.data(), pts.data(), &par);
            n_good[i] = nGood;
            if (nGood > bestGood) {
                secondBestGood = bestGood;
                bestGood = nGood;
                bestIdx = i;
                bestParallax = (float)par;
                std::memcpy(p3d, pts.data(), 24 * (size_t)n);
                std::memcpy(triangulated, good.data(), n);
            } else if (nGood > secondBestGood) {
                secondBestGood = nGood;
            }
        }
        *parallax_out = bestParallax;
        if (secondBestGood < 0.75 * bestGood && bestParallax >= min_parallax && bestGood > min_triangulated && bestGood > ratio_h * n) {
            std::memcpy(R21, Rs[bestIdx], 72);
            std::memcpy(t21, ts[bestIdx], 24);
            return 1;
        }
        std::memset(p3d, 0, 24 * (size_t)n);
        std::memset(triangulated, 0, n);
        return 0;
    }
    // ---- ReconstructF
    double Kt[9], t1[9], E[9];
    transpose3(Km, Kt);
    mul3(Kt, model, t1);
    mul3(t1, Km, E);
    const Svd3 sv = svd3_sorted(E);
    double tv[3] = {sv.U[2], sv.U[5], sv.U[8]};
    {
        const double nn = std::sqrt(tv[0] * tv[0] + tv[1] * tv[1] + tv[2] * tv[2]);
        for (int r = 0; r < 3; ++r) tv[r] = tv[r] / nn;
    }
    const double W[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1}, Wt[9] = {0, 1, 0, -1, 0, 0, 0, 0, 1};
    double Vt[9], R1[9], R2[9];
    transpose3(sv.V, Vt);
    mul3(sv.U, W, t1);
    mul3(t1, Vt, R1);
    if (det3(R1) < 0)
        for (int k = 0; k < 9; ++k) R1[k] = -R1[k];
    mul3(sv.U, Wt, t1);
    mul3(t1, Vt, R2);
    if (det3(R2) < 0)
        for (int k = 0; k < 9; ++k) R2[k] = -R2[k];
    const double tneg[3] = {-tv[0], -tv[1], -tv[2]};
    const double* Rc[4] = {R1, R2, R1, R2};
    const double* tc[4] = {tv, tv, tneg, tneg};
    if (candidates)
        for (int i = 0; i < 4; ++i) {
            std::memcpy(candidates + 12 * i, Rc[i], 72);
            std::memcpy(candidates + 12 * i + 9, tc[i], 24);
        }
    std::vector<std::vector<uint8_t>> goods(4, std::vector<uint8_t>(n));
    std::vector<std::vector<double>> ptss(4, std::vector<double>(3 * (size_t)n));
    int g[4];
    double par[4];
    for (int i = 0; i < 4; ++i) {
        g[i] = check_rt(n, px1, px2, Rc[i], tc[i], K, 24.0f * sigma2, false, goods[i].data(), ptss[i].data(), &par[i]);
        n_good[i] = g[i];
    }
    const int maxGood = std::max(g[0], std::max(g[1], std::max(g[2], g[3])));
    const int minGood = std::max((int)(0.9 * N), min_triangulated);
    int similar = 0;
    for (int i = 0; i < 4; ++i)
        if (g[i] > 0.7 * maxGood) ++similar;
    if (maxGood < minGood || similar > 1) return 0;
    for (int i = 0; i < 4; ++i)
        if (maxGood == g[i]) {   // the reference's if / else-if chain: the first candidate with the maximum decides
            *parallax_out = par[i];
            if (par[i] > min_parallax) {
                std::memcpy(p3d, ptss[i].data(), 24 * (size_t)n);
                std::memcpy(triangulated, goods[i].data(), n);
                std::memcpy(R21, Rc[i], 72);
                std::memcpy(t21, tc[i], 24);
                return 1;
            }
            return 0;
        }
    return 0;
}
