// se3.cuh -- device-side Lie-group arithmetic with the operation order of the non-template Sophus the
// reference vendors (thirdparty/Sophus/sophus/so3.cpp:127-202, se3.cpp:59-95,170-220): SO3 = unit
// quaternion (re-normalised after every product), SE3 = SO3 + translation, tangent = [upsilon; omega].
// Eigen's Quaternion::_transformVector / toRotationMatrix / Quaternion(Matrix3) are restated.
#pragma once
#ifdef __CUDACC__
#include <cuda_runtime.h>
#else  // plain host C++ (the shim classes in ../host use the same arithmetic for SE3 composition)
#include <cmath>
#ifndef __host__
#define __host__
#define __device__
#endif
using std::sqrt; using std::sin; using std::cos; using std::tan; using std::atan;
#endif

namespace ygzb {

struct V3d {
    double x, y, z;
};
__host__ __device__ inline V3d v3(double x, double y, double z) { return V3d{x, y, z}; }
__host__ __device__ inline V3d operator+(V3d a, V3d b) { return V3d{a.x + b.x, a.y + b.y, a.z + b.z}; }
__host__ __device__ inline V3d operator-(V3d a, V3d b) { return V3d{a.x - b.x, a.y - b.y, a.z - b.z}; }
__host__ __device__ inline V3d operator*(double s, V3d a) { return V3d{s * a.x, s * a.y, s * a.z}; }
__host__ __device__ inline V3d cross3(V3d a, V3d b) {
    return V3d{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

struct Quatd {
    double w, x, y, z;
};
__host__ __device__ inline Quatd qnormalized(Quatd q) {
    const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return Quatd{q.w / n, q.x / n, q.y / n, q.z / n};
}
__host__ __device__ inline Quatd qmul(Quatd a, Quatd b) {
    return Quatd{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
                 a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

struct SE3d {
    Quatd q;
    V3d t;
};

__host__ __device__ inline V3d rotate(Quatd q, V3d v) {  // Eigen _transformVector
    const V3d qv{q.x, q.y, q.z};
    V3d uv = cross3(qv, v);
    uv = uv + uv;
    return v + q.w * uv + cross3(qv, uv);
}
__host__ __device__ inline V3d transform(const SE3d& T, V3d p) { return rotate(T.q, p) + T.t; }
__host__ __device__ inline SE3d se3_mul(const SE3d& a, const SE3d& b) {
    SE3d r;
    r.t = a.t + rotate(a.q, b.t);
    r.q = qnormalized(qmul(a.q, b.q));
    return r;
}
__host__ __device__ inline SE3d se3_inverse(const SE3d& a) {
    SE3d r;
    r.q = Quatd{a.q.w, -a.q.x, -a.q.y, -a.q.z};
    r.t = rotate(r.q, -1. * a.t);
    return r;
}
__host__ __device__ inline void quat_to_matrix(Quatd q, double R[3][3]) {  // Eigen toRotationMatrix
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
    R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
    R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}
__host__ __device__ inline Quatd matrix_to_quat(const double R[3][3]) {  // Eigen Quaternion(Matrix3)
    Quatd q;
    double t = R[0][0] + R[1][1] + R[2][2];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R[2][1] - R[1][2]) * t;
        q.y = (R[0][2] - R[2][0]) * t;
        q.z = (R[1][0] - R[0][1]) * t;
    } else {
        int i = 0;
        if (R[1][1] > R[0][0]) i = 1;
        if (R[2][2] > R[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R[i][i] - R[j][j] - R[k][k] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R[k][j] - R[j][k]) * t;
        v[j] = (R[j][i] + R[i][j]) * t;
        v[k] = (R[k][i] + R[i][k]) * t;
        q.x = v[0];
        q.y = v[1];
        q.z = v[2];
    }
    // boundary convention (not Sophus): canonical sign w >= 0, as produced by SO3::exp
    if (q.w < 0) q = Quatd{-q.w, -q.x, -q.y, -q.z};
    return qnormalized(q);
}
// 3x4 row-major [R|t] <-> SE3d
__host__ __device__ inline SE3d se3_from_mat(const double* T) {
    const double R[3][3] = {{T[0], T[1], T[2]}, {T[4], T[5], T[6]}, {T[8], T[9], T[10]}};
    SE3d r;
    r.q = matrix_to_quat(R);
    r.t = V3d{T[3], T[7], T[11]};
    return r;
}
__host__ __device__ inline void se3_to_mat(const SE3d& a, double* T) {
    double R[3][3];
    quat_to_matrix(a.q, R);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[4 * i + j] = R[i][j];
    T[3] = a.t.x;
    T[7] = a.t.y;
    T[11] = a.t.z;
}

constexpr double kSophusEps = 1e-10;

__host__ __device__ inline Quatd so3_exp(V3d om, double* theta) {
    *theta = sqrt(om.x * om.x + om.y * om.y + om.z * om.z);
    const double half = 0.5 * (*theta);
    double imag;
    const double real = cos(half);
    if (*theta < kSophusEps) {
        const double t2 = (*theta) * (*theta), t4 = t2 * t2;
        imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
    } else {
        imag = sin(half) / (*theta);
    }
    return qnormalized(Quatd{real, imag * om.x, imag * om.y, imag * om.z});
}

__host__ __device__ inline V3d so3_log(Quatd q, double* theta) {
    const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    const double w = q.w;
    double f;
    if (n < kSophusEps) f = 2. / w - 2. * (n * n) / (w * w * w);
    else f = 2 * atan(n / w) / n;
    *theta = f * n;
    return V3d{f * q.x, f * q.y, f * q.z};
}

// SE3::exp([upsilon; omega])
__host__ __device__ inline SE3d se3_exp(const double* u) {
    const V3d up{u[0], u[1], u[2]}, om{u[3], u[4], u[5]};
    double theta;
    SE3d r;
    r.q = so3_exp(om, &theta);
    const double Om[3][3] = {{0, -om.z, om.y}, {om.z, 0, -om.x}, {-om.y, om.x, 0}};
    double Om2[3][3], V[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Om2[i][j] = Om[i][0] * Om[0][j] + Om[i][1] * Om[1][j] + Om[i][2] * Om[2][j];
    if (theta < kSophusEps) {
        quat_to_matrix(r.q, V);
    } else {
        const double t2 = theta * theta;
        const double a = (1 - cos(theta)) / t2, b = (theta - sin(theta)) / (t2 * theta);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) V[i][j] = (i == j ? 1.0 : 0.0) + a * Om[i][j] + b * Om2[i][j];
    }
    r.t = V3d{V[0][0] * up.x + V[0][1] * up.y + V[0][2] * up.z, V[1][0] * up.x + V[1][1] * up.y + V[1][2] * up.z,
              V[2][0] * up.x + V[2][1] * up.y + V[2][2] * up.z};
    return r;
}

// SE3::log -> [upsilon; omega]
__host__ __device__ inline void se3_log(const SE3d& T, double* out) {
    double theta;
    const V3d om = so3_log(T.q, &theta);
    const double Om[3][3] = {{0, -om.z, om.y}, {om.z, 0, -om.x}, {-om.y, om.x, 0}};
    double Om2[3][3], Vi[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Om2[i][j] = Om[i][0] * Om[0][j] + Om[i][1] * Om[1][j] + Om[i][2] * Om[2][j];
    const double c = (theta < kSophusEps) ? (1. / 12.) : (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Vi[i][j] = (i == j ? 1.0 : 0.0) - 0.5 * Om[i][j] + c * Om2[i][j];
    out[0] = Vi[0][0] * T.t.x + Vi[0][1] * T.t.y + Vi[0][2] * T.t.z;
    out[1] = Vi[1][0] * T.t.x + Vi[1][1] * T.t.y + Vi[1][2] * T.t.z;
    out[2] = Vi[2][0] * T.t.x + Vi[2][1] * T.t.y + Vi[2][2] * T.t.z;
    out[3] = om.x;
    out[4] = om.y;
    out[5] = om.z;
}

}  // namespace ygzb
