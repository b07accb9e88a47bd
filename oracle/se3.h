// oracle/se3.h -- restatement of the non-template Sophus (thirdparty/Sophus, a621ff) types used on the
// hot path: SO3 = unit quaternion, SE3 = SO3 + translation; tangent order [upsilon(3); omega(3)].
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//   SO3::exp / expAndTheta   thirdparty/Sophus/sophus/so3.cpp:161-190
//   SO3::log / logAndTheta   thirdparty/Sophus/sophus/so3.cpp:127-159
//   SE3::operator*, inverse  thirdparty/Sophus/sophus/se3.cpp:59-95
//   SE3::exp                 thirdparty/Sophus/sophus/se3.cpp:170-198
//   SE3::log                 thirdparty/Sophus/sophus/se3.cpp:200-220
// Eigen pieces restated: Quaternion product / normalize / _transformVector / toRotationMatrix.
#pragma once
#include <cmath>

namespace ora {

const double kSmallEps = 1e-10;  // Sophus SMALL_EPS

struct V3 {
    double x, y, z;
    double& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 {
    double m[3][3];
};
inline V3 operator*(const M3& A, V3 v) {
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
inline M3 hat(V3 v) { return {{{0, -v.z, v.y}, {v.z, 0, -v.x}, {-v.y, v.x, 0}}}; }
inline M3 identity3() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }

struct Quat {
    double w, x, y, z;
    void normalize() {
        const double n = std::sqrt(w * w + x * x + y * y + z * z);
        w /= n;
        x /= n;
        y /= n;
        z /= n;
    }
};
inline Quat qmul(const Quat& a, const Quat& b) {
    return {a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
            a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}

struct SO3 {
    Quat q{1, 0, 0, 0};
    // Eigen Quaternion::_transformVector
    V3 operator*(V3 v) const {
        V3 qv{q.x, q.y, q.z};
        V3 uv = cross(qv, v);
        uv = uv + uv;
        return v + q.w * uv + cross(qv, uv);
    }
    SO3 operator*(const SO3& o) const {
        SO3 r;
        r.q = qmul(q, o.q);
        r.q.normalize();  // SO3::operator*= normalises (so3.cpp:64-70)
        return r;
    }
    SO3 inverse() const { return SO3{Quat{q.w, -q.x, -q.y, -q.z}}; }
    M3 matrix() const {  // Eigen toRotationMatrix
        const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
        const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
        const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
        const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
        return {{{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}}};
    }
    static SO3 exp_theta(V3 omega, double* theta) {
        *theta = norm(omega);
        const double half = 0.5 * (*theta);
        double imag;
        const double real = std::cos(half);
        if (*theta < kSmallEps) {
            const double t2 = (*theta) * (*theta), t4 = t2 * t2;
            imag = 0.5 - 0.0208333 * t2 + 0.000260417 * t4;
        } else {
            imag = std::sin(half) / (*theta);
        }
        SO3 r;
        r.q = Quat{real, imag * omega.x, imag * omega.y, imag * omega.z};
        r.q.normalize();  // SO3(Quaterniond) normalises (so3.cpp:43-48)
        return r;
    }
    static SO3 exp(V3 omega) {
        double th;
        return exp_theta(omega, &th);
    }
    static V3 log_theta(const SO3& o, double* theta) {
        const double n = norm(V3{o.q.x, o.q.y, o.q.z});
        const double w = o.q.w;
        double f;
        if (n < kSmallEps) f = 2. / w - 2. * (n * n) / (w * w * w);
        else f = 2 * std::atan(n / w) / n;  // (the |w| < eps branch is overwritten in the reference, so3.cpp:143-155)
        *theta = f * n;
        return {f * o.q.x, f * o.q.y, f * o.q.z};
    }
    V3 log() const {
        double th;
        return log_theta(*this, &th);
    }
    // from a rotation matrix (Eigen Quaternion(Matrix3d)), used only at the oracle's C boundary
    static SO3 from_matrix(const M3& R) {
        Quat q;
        double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
        if (t > 0) {
            t = std::sqrt(t + 1.0);
            q.w = 0.5 * t;
            t = 0.5 / t;
            q.x = (R.m[2][1] - R.m[1][2]) * t;
            q.y = (R.m[0][2] - R.m[2][0]) * t;
            q.z = (R.m[1][0] - R.m[0][1]) * t;
        } else {
            int i = 0;
            if (R.m[1][1] > R.m[0][0]) i = 1;
            if (R.m[2][2] > R.m[i][i]) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
            double v[3];
            v[i] = 0.5 * t;
            t = 0.5 / t;
            q.w = (R.m[k][j] - R.m[j][k]) * t;
            v[j] = (R.m[j][i] + R.m[i][j]) * t;
            v[k] = (R.m[k][i] + R.m[i][k]) * t;
            q.x = v[0];
            q.y = v[1];
            q.z = v[2];
        }
        // boundary convention (not Sophus): canonical sign w >= 0, so that logAndTheta's theta is >= 0 as it
        // always is for quaternions produced by SO3::exp (a negative theta would take SE3::log's small-angle branch)
        if (q.w < 0) q = Quat{-q.w, -q.x, -q.y, -q.z};
        SO3 r;
        r.q = q;
        r.q.normalize();
        return r;
    }
};

struct SE3 {
    SO3 so3;
    V3 t{0, 0, 0};
    V3 operator*(V3 p) const { return so3 * p + t; }
    SE3 operator*(const SE3& o) const {
        SE3 r;
        r.t = t + so3 * o.t;
        r.so3 = so3 * o.so3;
        return r;
    }
    SE3 inverse() const {
        SE3 r;
        r.so3 = so3.inverse();
        r.t = r.so3 * (-1. * t);
        return r;
    }
    static SE3 exp(const double* u /* upsilon(3), omega(3) */) {
        V3 upsilon{u[0], u[1], u[2]}, omega{u[3], u[4], u[5]};
        double theta;
        SE3 r;
        r.so3 = SO3::exp_theta(omega, &theta);
        const M3 Om = hat(omega), Om2 = Om * Om;
        M3 V;
        if (theta < kSmallEps) {
            V = r.so3.matrix();
        } else {
            const double t2 = theta * theta;
            const double a = (1 - std::cos(theta)) / t2, b = (theta - std::sin(theta)) / (t2 * theta);
            V = identity3();
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) V.m[i][j] = V.m[i][j] + a * Om.m[i][j] + b * Om2.m[i][j];
        }
        r.t = V * upsilon;
        return r;
    }
    void log(double* out) const {
        double theta;
        const V3 om = SO3::log_theta(so3, &theta);
        const M3 Om = hat(om), Om2 = Om * Om;
        M3 Vi = identity3();
        const double c = (theta < kSmallEps) ? (1. / 12.) : (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Vi.m[i][j] = Vi.m[i][j] - 0.5 * Om.m[i][j] + c * Om2.m[i][j];
        const V3 up = Vi * t;
        out[0] = up.x;
        out[1] = up.y;
        out[2] = up.z;
        out[3] = om.x;
        out[4] = om.y;
        out[5] = om.z;
    }
    // 3x4 row-major [R|t] at the C boundary
    static SE3 from_mat(const double* T) {
        M3 R = {{{T[0], T[1], T[2]}, {T[4], T[5], T[6]}, {T[8], T[9], T[10]}}};
        SE3 r;
        r.so3 = SO3::from_matrix(R);
        r.t = {T[3], T[7], T[11]};
        return r;
    }
    void to_mat(double* T) const {
        const M3 R = so3.matrix();
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) T[4 * i + j] = R.m[i][j];
            T[4 * i + 3] = t[i];
        }
    }
};

}  // namespace ora
