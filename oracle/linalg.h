// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is shipped or linked into the
// product library; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use it, as the checker.
//
// linalg.h: dense fp64 matrix + the factorizations the reference takes from vikit_cg
// (cg::Matrix, cg::svd_fulluv), Eigen (LDLT, msckf_vio.cpp:850,924; feature.hpp:395) and
// SuiteSparse SPQR (msckf_vio.cpp:800-810).  None of those libraries is available here
// (SURVEY 8c); the EKF posterior does not depend on which orthonormal null-space basis or
// QR row signs are used because the measurement noise is sigma^2 I (msckf_vio.cpp:833,911),
// so Householder QR stands in for SVD and SPQR.  "parity unpinned" for these primitives.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

namespace orc {

struct Mat {
    int r = 0, c = 0;
    std::vector<double> d;
    Mat() {}
    Mat(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, 0.0) {}
    double &operator()(int i, int j) { return d[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return d[(size_t)i * c + j]; }
    static Mat eye(int n) {
        Mat m(n, n);
        for (int i = 0; i < n; ++i) m(i, i) = 1.0;
        return m;
    }
    Mat t() const {
        Mat m(c, r);
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j) m(j, i) = (*this)(i, j);
        return m;
    }
    Mat block(int i0, int j0, int nr, int nc) const {
        Mat m(nr, nc);
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc; ++j) m(i, j) = (*this)(i0 + i, j0 + j);
        return m;
    }
    void set(int i0, int j0, const Mat &m) {
        for (int i = 0; i < m.r; ++i)
            for (int j = 0; j < m.c; ++j) (*this)(i0 + i, j0 + j) = m(i, j);
    }
    // cg::Matrix::conservative_resize: keep the top-left overlap, zero the rest
    void conservative_resize(int nr, int nc) {
        Mat m(nr, nc);
        for (int i = 0; i < std::min(r, nr); ++i)
            for (int j = 0; j < std::min(c, nc); ++j) m(i, j) = (*this)(i, j);
        *this = m;
    }
};

inline Mat operator*(const Mat &a, const Mat &b) {
    assert(a.c == b.r);
    Mat m(a.r, b.c);
    for (int i = 0; i < a.r; ++i)
        for (int k = 0; k < a.c; ++k) {
            double aik = a(i, k);
            if (aik == 0.0) continue;
            const double *bp = &b.d[(size_t)k * b.c];
            double *mp = &m.d[(size_t)i * m.c];
            for (int j = 0; j < b.c; ++j) mp[j] += aik * bp[j];
        }
    return m;
}
inline Mat operator+(const Mat &a, const Mat &b) {
    assert(a.r == b.r && a.c == b.c);
    Mat m = a;
    for (size_t i = 0; i < m.d.size(); ++i) m.d[i] += b.d[i];
    return m;
}
inline Mat operator-(const Mat &a, const Mat &b) {
    assert(a.r == b.r && a.c == b.c);
    Mat m = a;
    for (size_t i = 0; i < m.d.size(); ++i) m.d[i] -= b.d[i];
    return m;
}
inline Mat operator*(const Mat &a, double s) {
    Mat m = a;
    for (auto &v : m.d) v *= s;
    return m;
}
inline Mat operator*(double s, const Mat &a) { return a * s; }
inline Mat operator-(const Mat &a) { return a * -1.0; }

struct V3 {
    double v[3] = {0, 0, 0};  // cg::Vector3() is assumed zero-initialised
    V3() {}
    V3(double a, double b, double c) { v[0] = a; v[1] = b; v[2] = c; }
    double &operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double dot(const V3 &o) const { return v[0] * o.v[0] + v[1] * o.v[1] + v[2] * o.v[2]; }
    double norm() const { return std::sqrt(dot(*this)); }
};
inline V3 operator+(const V3 &a, const V3 &b) { return V3(a[0] + b[0], a[1] + b[1], a[2] + b[2]); }
inline V3 operator-(const V3 &a, const V3 &b) { return V3(a[0] - b[0], a[1] - b[1], a[2] - b[2]); }
inline V3 operator-(const V3 &a) { return V3(-a[0], -a[1], -a[2]); }
inline V3 operator*(const V3 &a, double s) { return V3(a[0] * s, a[1] * s, a[2] * s); }
inline V3 operator*(double s, const V3 &a) { return a * s; }
inline V3 operator/(const V3 &a, double s) { return V3(a[0] / s, a[1] / s, a[2] / s); }

struct M3 {
    double m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    double &operator()(int i, int j) { return m[i * 3 + j]; }
    double operator()(int i, int j) const { return m[i * 3 + j]; }
    static M3 eye() {
        M3 r;
        r(0, 0) = r(1, 1) = r(2, 2) = 1.0;
        return r;
    }
    M3 t() const {
        M3 r;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) r(j, i) = (*this)(i, j);
        return r;
    }
};
inline M3 operator*(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += a(i, k) * b(k, j);
            r(i, j) = s;
        }
    return r;
}
inline V3 operator*(const M3 &a, const V3 &x) {
    V3 r;
    for (int i = 0; i < 3; ++i) r[i] = a(i, 0) * x[0] + a(i, 1) * x[1] + a(i, 2) * x[2];
    return r;
}
inline M3 operator*(const M3 &a, double s) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * s;
    return r;
}
inline M3 operator-(const M3 &a) { return a * -1.0; }
inline M3 operator+(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i];
    return r;
}
inline M3 operator-(const M3 &a, const M3 &b) {
    M3 r;
    for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i];
    return r;
}
inline M3 outer(const V3 &a, const V3 &b) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r(i, j) = a[i] * b[j];
    return r;
}
// cg::skew_symmetric
inline M3 skew(const V3 &w) {
    M3 r;
    r(0, 1) = -w[2]; r(0, 2) = w[1];
    r(1, 0) = w[2];  r(1, 2) = -w[0];
    r(2, 0) = -w[1]; r(2, 1) = w[0];
    return r;
}
// closed-form adjugate inverse
inline M3 inv3(const M3 &a) {
    M3 r;
    double c00 = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
    double c01 = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
    double c02 = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
    double det = a(0, 0) * c00 + a(0, 1) * c01 + a(0, 2) * c02;
    double id = 1.0 / det;
    r(0, 0) = c00 * id;
    r(0, 1) = (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id;
    r(0, 2) = (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id;
    r(1, 0) = c01 * id;
    r(1, 1) = (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id;
    r(1, 2) = (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id;
    r(2, 0) = c02 * id;
    r(2, 1) = (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id;
    r(2, 2) = (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id;
    return r;
}
inline Mat toMat(const M3 &a) {
    Mat m(3, 3);
    for (int i = 0; i < 9; ++i) m.d[i] = a.m[i];
    return m;
}
inline Mat toMat(const V3 &a) {
    Mat m(3, 1);
    for (int i = 0; i < 3; ++i) m.d[i] = a[i];
    return m;
}

// Rigid transform (cg::EuclideanTransform): x -> R x + t
struct SE3 {
    M3 R = M3::eye();
    V3 t;
    SE3() {}
    SE3(const M3 &R_, const V3 &t_) : R(R_), t(t_) {}
    static SE3 from16(const double *m) {
        SE3 T;
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) T.R(i, j) = m[i * 4 + j];
            T.t[i] = m[i * 4 + 3];
        }
        return T;
    }
    SE3 inv() const {
        SE3 T;
        T.R = R.t();
        T.t = -(T.R * t);
        return T;
    }
};
inline SE3 operator*(const SE3 &a, const SE3 &b) { return SE3(a.R * b.R, a.R * b.t + a.t); }

// ---------------------------------------------------------------------------------------
// Householder QR, in place.  On return the upper triangle of A holds R; the essential parts
// of the reflectors are below the diagonal, `tau` their scalars (LAPACK dgeqr2 convention).
inline void householder_qr(Mat &A, std::vector<double> &tau) {
    int m = A.r, n = A.c, k = std::min(m, n);
    tau.assign(k, 0.0);
    for (int j = 0; j < k; ++j) {
        double xnorm2 = 0;
        for (int i = j + 1; i < m; ++i) xnorm2 += A(i, j) * A(i, j);
        double alpha = A(j, j);
        if (xnorm2 == 0.0) { tau[j] = 0.0; continue; }
        double beta = -std::copysign(std::sqrt(alpha * alpha + xnorm2), alpha);
        tau[j] = (beta - alpha) / beta;
        double scale = 1.0 / (alpha - beta);
        for (int i = j + 1; i < m; ++i) A(i, j) *= scale;
        A(j, j) = beta;
        // apply H = I - tau v v^T to the trailing columns
        for (int c = j + 1; c < n; ++c) {
            double s = A(j, c);
            for (int i = j + 1; i < m; ++i) s += A(i, j) * A(i, c);
            s *= tau[j];
            A(j, c) -= s;
            for (int i = j + 1; i < m; ++i) A(i, c) -= s * A(i, j);
        }
    }
}
// B <- Q^T B, with Q from householder_qr(A)
inline void apply_qt(const Mat &A, const std::vector<double> &tau, Mat &B) {
    int m = A.r, k = (int)tau.size();
    for (int j = 0; j < k; ++j) {
        if (tau[j] == 0.0) continue;
        for (int c = 0; c < B.c; ++c) {
            double s = B(j, c);
            for (int i = j + 1; i < m; ++i) s += A(i, j) * B(i, c);
            s *= tau[j];
            B(j, c) -= s;
            for (int i = j + 1; i < m; ++i) B(i, c) -= s * A(i, j);
        }
    }
}
// Full m x m Q from householder_qr(A)
inline Mat form_q(const Mat &A, const std::vector<double> &tau) {
    int m = A.r, k = (int)tau.size();
    Mat Q = Mat::eye(m);
    for (int j = k - 1; j >= 0; --j) {
        if (tau[j] == 0.0) continue;
        for (int c = 0; c < m; ++c) {
            double s = Q(j, c);
            for (int i = j + 1; i < m; ++i) s += A(i, j) * Q(i, c);
            s *= tau[j];
            Q(j, c) -= s;
            for (int i = j + 1; i < m; ++i) Q(i, c) -= s * A(i, j);
        }
    }
    return Q;
}

// LDL^T (no pivoting) of a symmetric positive-definite matrix; solves S X = B.
// Stands in for Eigen's pivoted LDLT (msckf_vio.cpp:850,924): same solution for SPD S.
inline Mat ldlt_solve(const Mat &S, const Mat &B) {
    int n = S.r;
    Mat L = Mat::eye(n);
    std::vector<double> D(n);
    for (int j = 0; j < n; ++j) {
        double dj = S(j, j);
        for (int k = 0; k < j; ++k) dj -= L(j, k) * L(j, k) * D[k];
        D[j] = dj;
        for (int i = j + 1; i < n; ++i) {
            double s = S(i, j);
            for (int k = 0; k < j; ++k) s -= L(i, k) * L(j, k) * D[k];
            L(i, j) = s / dj;
        }
    }
    Mat X = B;
    int nc = X.c;
    for (int i = 0; i < n; ++i) {  // L y = b, row-oriented
        double *xi = &X.d[(size_t)i * nc];
        for (int k = 0; k < i; ++k) {
            double l = L(i, k);
            if (l == 0.0) continue;
            const double *xk = &X.d[(size_t)k * nc];
            for (int c = 0; c < nc; ++c) xi[c] -= l * xk[c];
        }
    }
    for (int i = 0; i < n; ++i) {
        double *xi = &X.d[(size_t)i * nc];
        for (int c = 0; c < nc; ++c) xi[c] /= D[i];
    }
    for (int i = n - 1; i >= 0; --i) {  // L^T x = z
        double *xi = &X.d[(size_t)i * nc];
        for (int k = i + 1; k < n; ++k) {
            double l = L(k, i);
            if (l == 0.0) continue;
            const double *xk = &X.d[(size_t)k * nc];
            for (int c = 0; c < nc; ++c) xi[c] -= l * xk[c];
        }
    }
    return X;
}

}  // namespace orc
