// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg.h).
//
// kin.h: rotation / quaternion helpers the reference takes from vikit_cg
// (kinematics/{quarternion,rotation_matrix,convertor,transform}.h — not in the tree).
// SPEC: the reference is a port of KumarRobotics msckf_vio, whose math_utils.hpp fixes the
// conventions (JPL quaternion [x y z w], world->body); those published conventions are
// restated here.  Call sites: msckf_vio.cpp:236,422,442,497-503,546,876-892,1251.
#pragma once
#include "linalg.h"

namespace orc {

struct Quat {  // JPL, [x y z w]
    double x = 0, y = 0, z = 0, w = 1;
    Quat() {}
    Quat(double x_, double y_, double z_, double w_) : x(x_), y(y_), z(z_), w(w_) {}
    double norm() const { return std::sqrt(x * x + y * y + z * z + w * w); }
    Quat normalized() const {
        double n = norm();
        return Quat(x / n, y / n, z / n, w / n);
    }
};

// cg::Quarternion::rotation_matrix(): R = (2w^2-1) I - 2w [q]x + 2 q q^T
inline M3 quat_to_rot(const Quat &q) {
    V3 qv(q.x, q.y, q.z);
    return M3::eye() * (2 * q.w * q.w - 1) - skew(qv) * (2 * q.w) + outer(qv, qv) * 2.0;
}

// cg::RotationMatrix::quarternion()
inline Quat rot_to_quat(const M3 &R) {
    double tr = R(0, 0) + R(1, 1) + R(2, 2);
    double score[4] = {R(0, 0), R(1, 1), R(2, 2), tr};
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (score[i] > score[best]) best = i;
    double q[4];
    if (best == 0) {
        q[0] = std::sqrt(1 + 2 * R(0, 0) - tr) / 2.0;
        q[1] = (R(0, 1) + R(1, 0)) / (4 * q[0]);
        q[2] = (R(0, 2) + R(2, 0)) / (4 * q[0]);
        q[3] = (R(1, 2) - R(2, 1)) / (4 * q[0]);
    } else if (best == 1) {
        q[1] = std::sqrt(1 + 2 * R(1, 1) - tr) / 2.0;
        q[0] = (R(0, 1) + R(1, 0)) / (4 * q[1]);
        q[2] = (R(1, 2) + R(2, 1)) / (4 * q[1]);
        q[3] = (R(2, 0) - R(0, 2)) / (4 * q[1]);
    } else if (best == 2) {
        q[2] = std::sqrt(1 + 2 * R(2, 2) - tr) / 2.0;
        q[0] = (R(0, 2) + R(2, 0)) / (4 * q[2]);
        q[1] = (R(1, 2) + R(2, 1)) / (4 * q[2]);
        q[3] = (R(0, 1) - R(1, 0)) / (4 * q[2]);
    } else {
        q[3] = std::sqrt(1 + tr) / 2.0;
        q[0] = (R(1, 2) - R(2, 1)) / (4 * q[3]);
        q[1] = (R(2, 0) - R(0, 2)) / (4 * q[3]);
        q[2] = (R(0, 1) - R(1, 0)) / (4 * q[3]);
    }
    if (q[3] < 0)
        for (int i = 0; i < 4; ++i) q[i] = -q[i];
    return Quat(q[0], q[1], q[2], q[3]).normalized();
}

// cg::Quarternion operator* (JPL product), normalised
inline Quat quat_mul(const Quat &a, const Quat &b) {
    Quat r;
    r.x = a.w * b.x + a.z * b.y - a.y * b.z + a.x * b.w;
    r.y = -a.z * b.x + a.w * b.y + a.x * b.z + a.y * b.w;
    r.z = a.y * b.x - a.x * b.y + a.w * b.z + a.z * b.w;
    r.w = -a.x * b.x - a.y * b.y - a.z * b.z + a.w * b.w;
    return r.normalized();
}

// cg::Quarternion::small_angle_quaternion
inline Quat small_angle_quat(const V3 &dtheta) {
    V3 dq = dtheta / 2.0;
    double n2 = dq.dot(dq);
    if (n2 <= 1) return Quat(dq[0], dq[1], dq[2], std::sqrt(1 - n2));
    double s = std::sqrt(1 + n2);
    return Quat(dq[0] / s, dq[1] / s, dq[2] / s, 1.0 / s);
}

// cg::RotationMatrix::quarternion_hamilton(): Hamilton [x y z w] of R (body->world use)
inline Quat rot_to_quat_hamilton(const M3 &R) {
    double tr = R(0, 0) + R(1, 1) + R(2, 2);
    double x, y, z, w;
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0) * 2;
        w = 0.25 * s;
        x = (R(2, 1) - R(1, 2)) / s;
        y = (R(0, 2) - R(2, 0)) / s;
        z = (R(1, 0) - R(0, 1)) / s;
    } else if (R(0, 0) > R(1, 1) && R(0, 0) > R(2, 2)) {
        double s = std::sqrt(1.0 + R(0, 0) - R(1, 1) - R(2, 2)) * 2;
        w = (R(2, 1) - R(1, 2)) / s;
        x = 0.25 * s;
        y = (R(0, 1) + R(1, 0)) / s;
        z = (R(0, 2) + R(2, 0)) / s;
    } else if (R(1, 1) > R(2, 2)) {
        double s = std::sqrt(1.0 + R(1, 1) - R(0, 0) - R(2, 2)) * 2;
        w = (R(0, 2) - R(2, 0)) / s;
        x = (R(0, 1) + R(1, 0)) / s;
        y = 0.25 * s;
        z = (R(1, 2) + R(2, 1)) / s;
    } else {
        double s = std::sqrt(1.0 + R(2, 2) - R(0, 0) - R(1, 1)) * 2;
        w = (R(1, 0) - R(0, 1)) / s;
        x = (R(0, 2) + R(2, 0)) / s;
        y = (R(1, 2) + R(2, 1)) / s;
        z = 0.25 * s;
    }
    return Quat(x, y, z, w);
}

// cg::rodrigues(v): rotation vector -> matrix (cv::Rodrigues semantics)
inline M3 rodrigues(const V3 &v) {
    double th = v.norm();
    if (th < 1e-12) return M3::eye() + skew(v);
    V3 k = v / th;
    double c = std::cos(th), s = std::sin(th);
    return M3::eye() * c + outer(k, k) * (1 - c) + skew(k) * s;
}

// cg::from_two_vector(a, b): rotation R with R a || b (Eigen FromTwoVectors semantics)
inline M3 from_two_vector(const V3 &a_, const V3 &b_) {
    V3 a = a_ / a_.norm(), b = b_ / b_.norm();
    double c = a.dot(b);
    if (c < -1 + 1e-12) {  // opposite: rotate pi about any axis orthogonal to a
        V3 ax = std::fabs(a[0]) < 0.9 ? V3(1, 0, 0) : V3(0, 1, 0);
        V3 k(a[1] * ax[2] - a[2] * ax[1], a[2] * ax[0] - a[0] * ax[2], a[0] * ax[1] - a[1] * ax[0]);
        k = k / k.norm();
        return outer(k, k) * 2.0 - M3::eye();
    }
    V3 v(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
    M3 K = skew(v);
    return M3::eye() + K + (K * K) * (1.0 / (1.0 + c));
}

// angle of Eigen::AngleAxisd(R) (msckf_vio.cpp:1054)
inline double rotation_angle(const M3 &R) {
    Quat q = rot_to_quat_hamilton(R);
    double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z);
    return 2.0 * std::atan2(n, std::fabs(q.w));
}

}  // namespace orc
