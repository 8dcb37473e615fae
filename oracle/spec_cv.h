// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg.h).
//
// spec_cv.h: CPU restatement of the pixel / point primitives that the reference calls but
// does not contain — they live in github.com/cggos/vikit_cg (unpinned, not vendored;
// README.md:9-13, msckf_core/CMakeLists.txt:59,64).  PARITY UNPINNED: the arithmetic below
// follows SPEC.md, which adopts the semantics of the OpenCV call that each reference call
// site replaced (commented-out code at image_processor.cpp:217-227, :130, :399-408,
// :809-816, :837-844).  pyr_down, FAST-9 score/NMS and the (un)distortion maps are
// cross-checked against Python cv2 in tests/golden/; the KLT and corner response are
// SPEC-defined (integer fixed point, so CPU and GPU agree bit for bit).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "linalg.h"

namespace orc {

struct Img {
    int rows = 0, cols = 0;
    std::vector<uint8_t> d;
    Img() {}
    Img(int r, int c) : rows(r), cols(c), d((size_t)r * c) {}
    uint8_t at(int y, int x) const { return d[(size_t)y * cols + x]; }
    uint8_t atc(int y, int x) const {  // replicate border
        y = y < 0 ? 0 : (y >= rows ? rows - 1 : y);
        x = x < 0 ? 0 : (x >= cols ? cols - 1 : x);
        return d[(size_t)y * cols + x];
    }
};
struct Pt {
    float x = 0, y = 0;
    Pt() {}
    Pt(float x_, float y_) : x(x_), y(y_) {}
};

// ---- cg::pyr_down (image_processor.cpp:239,242).  SPEC = cv::pyrDown: separable
// [1 4 6 4 1]/16 twice, (sum + 128) >> 8, BORDER_REFLECT_101, out = ((c+1)/2, (r+1)/2).
inline int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}
inline void pyr_down(const Img &in, Img &out) {
    static const int k[5] = {1, 4, 6, 4, 1};
    out = Img((in.rows + 1) / 2, (in.cols + 1) / 2);
    for (int y = 0; y < out.rows; ++y)
        for (int x = 0; x < out.cols; ++x) {
            int s = 0;
            for (int j = -2; j <= 2; ++j) {
                int yy = reflect101(2 * y + j, in.rows);
                int rs = 0;
                for (int i = -2; i <= 2; ++i) rs += k[i + 2] * in.at(yy, reflect101(2 * x + i, in.cols));
                s += k[j + 2] * rs;
            }
            out.d[(size_t)y * out.cols + x] = (uint8_t)((s + 128) >> 8);
        }
}

// ---- CornerDetector(n_rows, n_cols, thr) (image_processor.cpp:132,647,657).
// SPEC: FAST-9/16 segment test with threshold t, score = largest threshold for which the
// pixel is still a corner (cv::FAST response), strict 3x3 non-max suppression, then one
// corner per fine cell (cell = rows/n_rows+1 by cols/n_cols+1 pixels) with the best
// Shi-Tomasi response over an 8x8 box, skipping cells marked by set_grid_position.
static const int kFastDx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int kFastDy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// S = max over the 16 arcs of 9 contiguous circle pixels of min(d) and of min(-d), d = centre - ring
inline int fast_arc_strength(const Img &im, int y, int x) {
    int v = im.at(y, x), d[25];
    for (int k = 0; k < 16; ++k) d[k] = v - im.at(y + kFastDy[k], x + kFastDx[k]);
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -255;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) {
            mn = std::min(mn, d[k + j]);
            mx = std::max(mx, d[k + j]);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return best;
}

inline float shi_tomasi(const Img &im, int y, int x) {
    if (x < 5 || y < 5 || x > im.cols - 6 || y > im.rows - 6) return 0.0f;
    int dXX = 0, dYY = 0, dXY = 0;
    for (int yy = y - 4; yy < y + 4; ++yy)
        for (int xx = x - 4; xx < x + 4; ++xx) {
            int dx = (int)im.at(yy, xx + 1) - (int)im.at(yy, xx - 1);
            int dy = (int)im.at(yy + 1, xx) - (int)im.at(yy - 1, xx);
            dXX += dx * dx;
            dYY += dy * dy;
            dXY += dx * dy;
        }
    float fXX = (float)dXX / 128.0f, fYY = (float)dYY / 128.0f, fXY = (float)dXY / 128.0f;
    float tr = fXX + fYY;
    float d1 = fXX - fYY;
    float xy2 = fXY * fXY;
    float disc = d1 * d1 + 4.0f * xy2;
    return 0.5f * (tr - std::sqrt(disc));
}

struct CornerDetector {
    int n_rows = 30, n_cols = 47, fast_threshold = 10;
    double detection_threshold = 10.0;
    int rows = 0, cols = 0, cell_h = 1, cell_w = 1;
    std::vector<uint8_t> occupancy;
    void configure(int img_rows, int img_cols) {
        rows = img_rows; cols = img_cols;
        cell_h = rows / n_rows + 1;
        cell_w = cols / n_cols + 1;
        occupancy.assign((size_t)n_rows * n_cols, 0);
    }
    int sub2ind(int x, int y) const { return (y / cell_h) * n_cols + (x / cell_w); }
    void set_grid_position(const Pt &p) {
        int x = (int)p.x, y = (int)p.y;
        if (x < 0 || y < 0 || x >= cols || y >= rows) return;
        occupancy[sub2ind(x, y)] = 1;
    }
    // scores: optional dump of the per-pixel FAST score map after the threshold (for tests)
    void detect_features(const Img &im, std::vector<Pt> &pts, std::vector<double> &resp,
                         std::vector<uint8_t> *score_map = nullptr) {
        std::vector<uint8_t> sc((size_t)rows * cols, 0);
        for (int y = 3; y < rows - 3; ++y)
            for (int x = 3; x < cols - 3; ++x) {
                int s = fast_arc_strength(im, y, x);
                if (s > fast_threshold) sc[(size_t)y * cols + x] = (uint8_t)(s - 1);
            }
        if (score_map) *score_map = sc;
        std::vector<float> best((size_t)n_rows * n_cols, 0.0f);
        std::vector<Pt> best_pt((size_t)n_rows * n_cols);
        for (int y = 3; y < rows - 3; ++y)
            for (int x = 3; x < cols - 3; ++x) {
                int s = sc[(size_t)y * cols + x];
                if (s == 0) continue;
                bool is_max = true;
                for (int j = -1; j <= 1 && is_max; ++j)
                    for (int i = -1; i <= 1; ++i)
                        if ((i || j) && sc[(size_t)(y + j) * cols + x + i] >= s) { is_max = false; break; }
                if (!is_max) continue;
                int k = sub2ind(x, y);
                if (occupancy[k]) continue;
                float st = shi_tomasi(im, y, x);
                if (st > best[k]) { best[k] = st; best_pt[k] = Pt((float)x, (float)y); }
            }
        pts.clear();
        resp.clear();
        for (int k = 0; k < n_rows * n_cols; ++k)
            if ((double)best[k] > detection_threshold) {
                pts.push_back(best_pt[k]);
                resp.push_back((double)best[k]);
            }
        std::fill(occupancy.begin(), occupancy.end(), 0);
    }
};

// ---- cg::optical_flow_multi_level(pyrA, pyrB, ptsA, ptsB, status, win, max_iters)
// (image_processor.cpp:410,569).  SPEC: pyramidal Lucas-Kanade in the style of
// cv::calcOpticalFlowPyrLK with OPTFLOW_USE_INITIAL_FLOW, restated in integer fixed point:
// 14-bit bilinear weights, samples kept with 5 fractional bits, central-difference template
// gradients, int64 sums for the 2x2 system and mismatch vector, fp64 solve.
struct KltParams {
    int win = 15, max_iters = 30;
    double eps = 0.01, min_eig = 1e-4;
};

struct BilinW { int ix, iy, w00, w01, w10, w11; };
inline BilinW bilin_weights(float x, float y) {
    BilinW b;
    float fx = std::floor(x), fy = std::floor(y);
    b.ix = (int)fx; b.iy = (int)fy;
    float a = x - fx, c = y - fy;
    b.w00 = (int)std::lrintf((1.f - a) * (1.f - c) * 16384.f);
    b.w01 = (int)std::lrintf(a * (1.f - c) * 16384.f);
    b.w10 = (int)std::lrintf((1.f - a) * c * 16384.f);
    b.w11 = 16384 - b.w00 - b.w01 - b.w10;
    return b;
}
inline int sample_fx(const Img &im, const BilinW &b, int i, int j) {
    int x = b.ix + i, y = b.iy + j;
    int s = b.w00 * im.atc(y, x) + b.w01 * im.atc(y, x + 1) + b.w10 * im.atc(y + 1, x) + b.w11 * im.atc(y + 1, x + 1);
    return (s + 256) >> 9;
}

inline void klt_track_one(const std::vector<Img> &pa, const std::vector<Img> &pb, const Pt &p0, Pt &q0,
                          uint8_t &status, const KltParams &kp) {
    const int L = (int)pa.size(), win = kp.win, half = win / 2, tw = win + 2;
    std::vector<int> T((size_t)tw * tw), Ix((size_t)win * win), Iy((size_t)win * win);
    status = 1;
    float top = 1.0f / (float)(1 << (L - 1));
    float qx = q0.x * top, qy = q0.y * top;
    for (int l = L - 1; l >= 0; --l) {
        const Img &A = pa[l], &B = pb[l];
        float s = 1.0f / (float)(1 << l);
        float px = p0.x * s, py = p0.y * s;
        BilinW wa = bilin_weights(px, py);
        for (int j = 0; j < tw; ++j)
            for (int i = 0; i < tw; ++i) T[(size_t)j * tw + i] = sample_fx(A, wa, i - half - 1, j - half - 1);
        int64_t A11 = 0, A12 = 0, A22 = 0;
        for (int j = 0; j < win; ++j)
            for (int i = 0; i < win; ++i) {
                int gx = T[(size_t)(j + 1) * tw + i + 2] - T[(size_t)(j + 1) * tw + i];
                int gy = T[(size_t)(j + 2) * tw + i + 1] - T[(size_t)j * tw + i + 1];
                Ix[(size_t)j * win + i] = gx;
                Iy[(size_t)j * win + i] = gy;
                A11 += (int64_t)gx * gx;
                A12 += (int64_t)gx * gy;
                A22 += (int64_t)gy * gy;
            }
        double a11 = (double)A11, a12 = (double)A12, a22 = (double)A22;
        double m1 = a11 * a22, m2 = a12 * a12;
        double D = m1 - m2;
        double df = a11 - a22;
        double disc = df * df + 4.0 * m2;
        double lam = (a11 + a22 - std::sqrt(disc)) * 0.5;
        double min_eig = lam / (4194304.0 * (double)(win * win));
        bool ok = !(min_eig < kp.min_eig || D < 1.1920929e-07);
        if (!ok) {
            if (l == 0) status = 0;
        } else {
            double Dinv = 1.0 / D;
            double pdx = 0, pdy = 0;
            for (int it = 0; it < kp.max_iters; ++it) {
                if (qx < 0.f || qy < 0.f || qx > (float)(B.cols - 1) || qy > (float)(B.rows - 1)) {
                    if (l == 0) status = 0;
                    break;
                }
                BilinW wb = bilin_weights(qx, qy);
                int64_t b1 = 0, b2 = 0;
                for (int j = 0; j < win; ++j)
                    for (int i = 0; i < win; ++i) {
                        int diff = sample_fx(B, wb, i - half, j - half) - T[(size_t)(j + 1) * tw + i + 1];
                        b1 += (int64_t)diff * Ix[(size_t)j * win + i];
                        b2 += (int64_t)diff * Iy[(size_t)j * win + i];
                    }
                double fb1 = (double)b1, fb2 = (double)b2;
                double dx = (a12 * fb2 - a22 * fb1) * Dinv * 2.0;
                double dy = (a12 * fb1 - a11 * fb2) * Dinv * 2.0;
                float fdx = (float)dx, fdy = (float)dy;
                qx += fdx;
                qy += fdy;
                if (dx * dx + dy * dy <= kp.eps * kp.eps) break;
                if (it > 0 && std::fabs(dx + pdx) < 0.01 && std::fabs(dy + pdy) < 0.01) {
                    qx -= fdx * 0.5f;
                    qy -= fdy * 0.5f;
                    break;
                }
                pdx = dx;
                pdy = dy;
            }
        }
        if (l > 0) { qx *= 2.0f; qy *= 2.0f; }
    }
    // a lost track at level 0 leaves a possibly out-of-image point; callers test status first
    q0.x = qx;
    q0.y = qy;
}

inline void optical_flow_multi_level(const std::vector<Img> &pa, const std::vector<Img> &pb,
                                     const std::vector<Pt> &pts_a, std::vector<Pt> &pts_b,
                                     std::vector<uint8_t> &status, const KltParams &kp) {
    status.assign(pts_a.size(), 0);
    pts_b.resize(pts_a.size());
    for (size_t i = 0; i < pts_a.size(); ++i) klt_track_one(pa, pb, pts_a[i], pts_b[i], status[i], kp);
}

// ---- cg::undistort_points / undistort_points_fisheye / project_points /
// distort_points_fisheye (image_processor.cpp:810-844).  SPEC = cv::undistortPoints
// (5 fixed-point iterations), cv::fisheye::undistortPoints (10 Newton steps),
// cv::projectPoints with zero rvec/tvec on (x, y, 1), cv::fisheye::distortPoints. fp64
// inside, float in/out.
inline void undistort_points(const std::vector<Pt> &in, std::vector<Pt> &out, const double K[4],
                             int model, const double D[4], const M3 &R, const double Kn[4]) {
    out.resize(in.size());
    for (size_t n = 0; n < in.size(); ++n) {
        double x = ((double)in[n].x - K[2]) / K[0], y = ((double)in[n].y - K[3]) / K[1];
        if (model == 0) {
            double x0 = x, y0 = y;
            for (int it = 0; it < 5; ++it) {
                double r2 = x * x + y * y;
                double icd = 1.0 / (1.0 + (D[1] * r2 + D[0]) * r2);
                double dx = 2.0 * D[2] * x * y + D[3] * (r2 + 2.0 * x * x);
                double dy = D[2] * (r2 + 2.0 * y * y) + 2.0 * D[3] * x * y;
                x = (x0 - dx) * icd;
                y = (y0 - dy) * icd;
            }
        } else {
            double thd = std::sqrt(x * x + y * y);
            thd = std::min(std::max(-M_PI / 2., thd), M_PI / 2.);
            double scale = 1.0;
            if (thd > 1e-8) {
                double th = thd;
                for (int it = 0; it < 10; ++it) {
                    double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
                    double k0t2 = D[0] * t2, k1t4 = D[1] * t4, k2t6 = D[2] * t6, k3t8 = D[3] * t8;
                    double fix = (th * (1 + k0t2 + k1t4 + k2t6 + k3t8) - thd) /
                                 (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8);
                    th = th - fix;
                    if (std::fabs(fix) < 1e-10) break;
                }
                scale = std::tan(th) / thd;
            }
            x *= scale;
            y *= scale;
        }
        double X = R(0, 0) * x + R(0, 1) * y + R(0, 2);
        double Y = R(1, 0) * x + R(1, 1) * y + R(1, 2);
        double W = R(2, 0) * x + R(2, 1) * y + R(2, 2);
        x = X / W;
        y = Y / W;
        out[n].x = (float)(x * Kn[0] + Kn[2]);
        out[n].y = (float)(y * Kn[1] + Kn[3]);
    }
}

inline void distort_points(const std::vector<Pt> &in, std::vector<Pt> &out, const double K[4], int model,
                           const double D[4]) {
    out.resize(in.size());
    for (size_t n = 0; n < in.size(); ++n) {
        double x = (double)in[n].x, y = (double)in[n].y;
        double xd, yd;
        if (model == 0) {
            double r2 = x * x + y * y;
            double cd = 1.0 + (D[1] * r2 + D[0]) * r2;
            xd = x * cd + 2.0 * D[2] * x * y + D[3] * (r2 + 2.0 * x * x);
            yd = y * cd + D[2] * (r2 + 2.0 * y * y) + 2.0 * D[3] * x * y;
        } else {
            double r = std::sqrt(x * x + y * y);
            double th = std::atan(r);
            double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
            double thd = th * (1 + D[0] * t2 + D[1] * t4 + D[2] * t6 + D[3] * t8);
            double s = r > 1e-8 ? thd / r : 1.0;
            xd = x * s;
            yd = y * s;
        }
        out[n].x = (float)(xd * K[0] + K[2]);
        out[n].y = (float)(yd * K[1] + K[3]);
    }
}

}  // namespace orc
