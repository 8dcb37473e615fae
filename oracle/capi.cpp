// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg.h).
//
// capi.cpp: C entry points for ctypes (tests/, smoke(), bench.py's CPU legs).  `Oracle`
// re-hosts System (msckf_core/src/system.cpp:12-54): one ImageProcessor + one MsckfVio and
// the three forwarders; the feed order (IMU rows with t <= t_img plus one overshoot, then
// stereo_callback, then backend_callback; apps/run_euroc_single_thread.cpp:209-254) is
// driven by the caller.
#include "backend.h"

using namespace orc;

struct Oracle {
    ImageProcessor fe;
    MsckfVio be;
    explicit Oracle(const mskf_config &c) : fe(c), be(c) {}
};

extern "C" {

void *orc_create(const mskf_config *cfg) { return new Oracle(*cfg); }
void orc_destroy(void *h) { delete (Oracle *)h; }

// System::imu_callback, system.cpp:45-48
void orc_imu(void *h, double t, const double w[3], const double a[3]) {
    Oracle *o = (Oracle *)h;
    ImuMsg m{t, V3(w[0], w[1], w[2]), V3(a[0], a[1], a[2])};
    o->fe.imuCallback(m);
    o->be.imuCallback(m);
}
// System::stereo_callback, system.cpp:40-43
void orc_stereo(void *h, double t, const uint8_t *cam0, const uint8_t *cam1) { ((Oracle *)h)->fe.stereoCallback(t, cam0, cam1); }
// System::backend_callback, system.cpp:50-54
void orc_backend(void *h) {
    Oracle *o = (Oracle *)h;
    o->be.featureCallback(*o->fe.feature_msg);
}
void orc_backend_features(void *h, double t, const mskf_feature *f, int n) {
    Oracle *o = (Oracle *)h;
    CameraMeasurement msg;
    msg.time_stamp = t;
    msg.features.resize(n);
    for (int i = 0; i < n; ++i) {
        msg.features[i].id = f[i].id;
        msg.features[i].u0 = f[i].u0;
        msg.features[i].v0 = f[i].v0;
        msg.features[i].u1 = f[i].u1;
        msg.features[i].v1 = f[i].v1;
    }
    o->be.featureCallback(msg);
}
void orc_reset(void *h) { ((Oracle *)h)->be.resetCallback(); }

int orc_get_features(void *h, mskf_feature *out, int cap, double *t) {
    Oracle *o = (Oracle *)h;
    const auto &v = o->fe.feature_msg->features;
    if (t) *t = o->fe.feature_msg->time_stamp;
    for (int i = 0; i < (int)v.size() && i < cap; ++i) {
        out[i].id = v[i].id;
        out[i].pad = 0;
        out[i].u0 = v[i].u0; out[i].v0 = v[i].v0; out[i].u1 = v[i].u1; out[i].v1 = v[i].v1;
    }
    return (int)v.size();
}
int orc_get_n_published(void *h) { return ((Oracle *)h)->fe.n_published; }
void orc_get_tracking_info(void *h, mskf_tracking_info *ti) {
    Oracle *o = (Oracle *)h;
    ti->time_stamp = o->fe.feature_msg->time_stamp;
    ti->before_tracking = o->fe.before_tracking;
    ti->after_tracking = o->fe.after_tracking;
    ti->after_matching = o->fe.after_matching;
    ti->after_ransac = o->fe.after_ransac;
}
int orc_get_grid(void *h, mskf_grid_feature *out, int cap) {
    Oracle *o = (Oracle *)h;
    int n = 0;
    for (const auto &cell : *o->fe.prev_features)
        for (const auto &f : cell.second) {
            if (n < cap) {
                out[n].id = f.id; out[n].response = f.response; out[n].lifetime = f.lifetime;
                out[n].cam0_x = f.cam0_point.x; out[n].cam0_y = f.cam0_point.y;
                out[n].cam1_x = f.cam1_point.x; out[n].cam1_y = f.cam1_point.y;
                out[n].cell = cell.first; out[n].pad = 0;
            }
            ++n;
        }
    return n;
}
int orc_get_pyramid(void *h, int cam, int level, uint8_t *out, int cap, int *rows, int *cols) {
    Oracle *o = (Oracle *)h;
    const std::vector<Img> &p = cam == 0 ? o->fe.prev_cam0_pyramid : o->fe.curr_cam1_pyramid;
    if (level < 0 || level >= (int)p.size()) return -1;
    *rows = p[level].rows; *cols = p[level].cols;
    if ((int)p[level].d.size() > cap) return -3;
    std::memcpy(out, p[level].d.data(), p[level].d.size());
    return 0;
}
void orc_get_state(void *h, mskf_state *s) {
    Oracle *o = (Oracle *)h;
    const MsckfVio &b = o->be;
    std::memset(s, 0, sizeof(*s));
    s->time = b.imu_state.time; s->id = b.imu_state.id;
    s->orientation[0] = b.imu_state.orientation.x; s->orientation[1] = b.imu_state.orientation.y;
    s->orientation[2] = b.imu_state.orientation.z; s->orientation[3] = b.imu_state.orientation.w;
    for (int i = 0; i < 3; ++i) {
        s->position[i] = b.imu_state.position[i]; s->velocity[i] = b.imu_state.velocity[i];
        s->gyro_bias[i] = b.imu_state.gyro_bias[i]; s->acc_bias[i] = b.imu_state.acc_bias[i];
        s->t_cam0_imu[i] = b.imu_state.t_cam0_imu[i]; s->gravity[i] = b.gravity[i];
    }
    for (int i = 0; i < 9; ++i) s->R_imu_cam0[i] = b.imu_state.R_imu_cam0.m[i];
    s->n_cam_states = (int)b.cam_states.size(); s->cov_dim = b.state_cov.r;
    s->is_gravity_set = b.is_gravity_set; s->n_map_features = (int)b.map_server.size();
    s->tracking_rate = b.tracking_rate;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) s->T_b_w[i * 4 + j] = b.T_b_w.R(i, j);
        s->T_b_w[i * 4 + 3] = b.T_b_w.t[i];
    }
    s->T_b_w[15] = 1.0;
    s->n_updates = b.n_updates; s->n_resets = b.n_resets;
}
int orc_get_cam_states(void *h, mskf_cam_state *out, int cap) {
    Oracle *o = (Oracle *)h;
    int n = 0;
    for (const auto &kv : o->be.cam_states) {
        if (n < cap) {
            out[n].id = kv.first; out[n].time = kv.second.time;
            out[n].orientation[0] = kv.second.orientation.x; out[n].orientation[1] = kv.second.orientation.y;
            out[n].orientation[2] = kv.second.orientation.z; out[n].orientation[3] = kv.second.orientation.w;
            for (int i = 0; i < 3; ++i) out[n].position[i] = kv.second.position[i];
        }
        ++n;
    }
    return n;
}
int orc_get_cov(void *h, double *out, int cap) {
    Oracle *o = (Oracle *)h;
    const Mat &P = o->be.state_cov;
    if ((int)P.d.size() <= cap) std::memcpy(out, P.d.data(), P.d.size() * sizeof(double));
    return P.r;
}

// ---- stand-alone primitives -------------------------------------------------------------
void orc_pyr_down(const uint8_t *in, int rows, int cols, uint8_t *out) {
    Img a(rows, cols), b;
    std::memcpy(a.d.data(), in, a.d.size());
    pyr_down(a, b);
    std::memcpy(out, b.d.data(), b.d.size());
}
static void build_pyr(const uint8_t *img, int rows, int cols, int levels, std::vector<Img> &p) {
    p.clear();
    Img a(rows, cols);
    std::memcpy(a.d.data(), img, a.d.size());
    p.push_back(a);
    for (int i = 1; i < levels; ++i) {
        Img t;
        pyr_down(p[i - 1], t);
        p.push_back(t);
    }
}
int orc_detect(const mskf_config *cfg, const uint8_t *img, const float *occ_xy, int n_occ, float *out_xy,
               double *out_resp, int cap, uint8_t *score_map) {
    CornerDetector d;
    d.n_rows = cfg->det_rows; d.n_cols = cfg->det_cols;
    d.fast_threshold = cfg->fast_threshold; d.detection_threshold = cfg->detection_threshold;
    d.configure(cfg->img_rows, cfg->img_cols);
    Img a(cfg->img_rows, cfg->img_cols);
    std::memcpy(a.d.data(), img, a.d.size());
    for (int i = 0; i < n_occ; ++i) d.set_grid_position(Pt((float)(int)occ_xy[2 * i], (float)(int)occ_xy[2 * i + 1]));
    std::vector<Pt> pts;
    std::vector<double> resp;
    std::vector<uint8_t> sm;
    d.detect_features(a, pts, resp, score_map ? &sm : nullptr);
    if (score_map) std::memcpy(score_map, sm.data(), sm.size());
    for (int i = 0; i < (int)pts.size() && i < cap; ++i) {
        out_xy[2 * i] = pts[i].x; out_xy[2 * i + 1] = pts[i].y;
        out_resp[i] = resp[i];
    }
    return (int)pts.size();
}
void orc_klt(const mskf_config *cfg, const uint8_t *a, const uint8_t *b, const float *pts_a, float *pts_b,
             uint8_t *status, int n) {
    std::vector<Img> pa, pb;
    build_pyr(a, cfg->img_rows, cfg->img_cols, cfg->pyramid_levels, pa);
    build_pyr(b, cfg->img_rows, cfg->img_cols, cfg->pyramid_levels, pb);
    KltParams kp;
    kp.win = cfg->klt_win; kp.max_iters = cfg->klt_max_iters; kp.eps = cfg->klt_eps; kp.min_eig = cfg->klt_min_eig;
    std::vector<Pt> A(n), B(n);
    for (int i = 0; i < n; ++i) {
        A[i] = Pt(pts_a[2 * i], pts_a[2 * i + 1]);
        B[i] = Pt(pts_b[2 * i], pts_b[2 * i + 1]);
    }
    std::vector<uint8_t> st;
    optical_flow_multi_level(pa, pb, A, B, st, kp);
    for (int i = 0; i < n; ++i) {
        pts_b[2 * i] = B[i].x; pts_b[2 * i + 1] = B[i].y;
        status[i] = st[i];
    }
}
void orc_undistort(const float *in, int n, const double K[4], int model, const double D[4], const double R[9],
                   const double Kn[4], float *out) {
    std::vector<Pt> a(n), b;
    for (int i = 0; i < n; ++i) a[i] = Pt(in[2 * i], in[2 * i + 1]);
    M3 Rm;
    for (int i = 0; i < 9; ++i) Rm.m[i] = R[i];
    undistort_points(a, b, K, model, D, Rm, Kn);
    for (int i = 0; i < n; ++i) { out[2 * i] = b[i].x; out[2 * i + 1] = b[i].y; }
}
void orc_distort(const float *in, int n, const double K[4], int model, const double D[4], float *out) {
    std::vector<Pt> a(n), b;
    for (int i = 0; i < n; ++i) a[i] = Pt(in[2 * i], in[2 * i + 1]);
    distort_points(a, b, K, model, D);
    for (int i = 0; i < n; ++i) { out[2 * i] = b[i].x; out[2 * i + 1] = b[i].y; }
}
// null-space projection + gating ingredients of one feature, for the basis-invariance test
void orc_last_update_debug(void *h, double *delta_x, int cap, int *n, double *gamma) {
    Oracle *o = (Oracle *)h;
    *n = o->be.last_delta_x.r;
    for (int i = 0; i < *n && i < cap; ++i) delta_x[i] = o->be.last_delta_x(i, 0);
    *gamma = o->be.last_gamma;
}

// ---- linear-algebra hooks for the CPU tests (linalg.h / kin.h against numpy, and the
// null-space basis invariance property of SURVEY 8c) ----------------------------------------
void orc_rodrigues(const double v[3], double R[9]) {
    M3 m = rodrigues(V3(v[0], v[1], v[2]));
    for (int i = 0; i < 9; ++i) R[i] = m.m[i];
}
static Mat mat_from(const double *d, int r, int c) {
    Mat m(r, c);
    std::memcpy(m.d.data(), d, sizeof(double) * (size_t)r * c);
    return m;
}
// thin QR: R (n x n, upper) and Q^T b (first n entries) of A (m x n), m >= n
void orc_qr_thin(const double *A, int m, int n, const double *b, double *R, double *qtb) {
    Mat QR = mat_from(A, m, n), rb = mat_from(b, m, 1);
    std::vector<double> tau;
    householder_qr(QR, tau);
    apply_qt(QR, tau, rb);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) R[i * n + j] = j >= i ? QR(i, j) : 0.0;
    for (int i = 0; i < n; ++i) qtb[i] = rb(i, 0);
}
void orc_ldlt_solve(const double *S, int n, const double *B, int nc, double *X) {
    Mat x = ldlt_solve(mat_from(S, n, n), mat_from(B, n, nc));
    std::memcpy(X, x.d.data(), sizeof(double) * (size_t)n * nc);
}
// featureJacobian's projection (msckf_vio.cpp:757-766) + gatingTest (:909-935) +
// update_math on one feature block.  basis == nullptr: Householder basis (what the oracle
// runs); otherwise `basis` is an explicit 4M x (4M-3) orthonormal null-space basis A, e.g.
// the last columns of U from a full SVD as the reference computes it.
void orc_nullspace_update(int n, int rows, const double *Hx, const double *Hf, const double *r, const double *P,
                          double obs_noise, const double *basis, double *out_dx, double *out_P, double *out_gamma) {
    Mat Hxj = mat_from(Hx, rows, n), Hfj = mat_from(Hf, rows, 3), rj = mat_from(r, rows, 1), Pm = mat_from(P, n, n);
    Mat Hp, rp;
    if (!basis) {
        Mat QR = Hfj;
        std::vector<double> tau;
        householder_qr(QR, tau);
        Mat Hq = Hxj, rq = rj;
        apply_qt(QR, tau, Hq);
        apply_qt(QR, tau, rq);
        Hp = Hq.block(3, 0, rows - 3, n);
        rp = rq.block(3, 0, rows - 3, 1);
    } else {
        Mat A = mat_from(basis, rows, rows - 3);
        Hp = A.t() * Hxj;
        rp = A.t() * rj;
    }
    Mat S = Hp * Pm * Hp.t();
    for (int i = 0; i < S.r; ++i) S(i, i) += obs_noise;
    Mat x = ldlt_solve(S, rp);
    double g = 0;
    for (int i = 0; i < rp.r; ++i) g += rp(i, 0) * x(i, 0);
    *out_gamma = g;
    Mat dx, Pn;
    MsckfVio::update_math(Hp, rp, Pm, obs_noise, dx, Pn);
    std::memcpy(out_dx, dx.d.data(), sizeof(double) * n);
    std::memcpy(out_P, Pn.d.data(), sizeof(double) * (size_t)n * n);
}
// measurementUpdate's algebra alone (msckf_vio.cpp:795-857,897-904)
void orc_update_math(int n, int m, const double *H, const double *r, const double *P, double obs_noise, double *out_dx,
                     double *out_P) {
    Mat dx, Pn;
    MsckfVio::update_math(mat_from(H, m, n), mat_from(r, m, 1), mat_from(P, n, n), obs_noise, dx, Pn);
    std::memcpy(out_dx, dx.d.data(), sizeof(double) * n);
    std::memcpy(out_P, Pn.d.data(), sizeof(double) * (size_t)n * n);
}
// test hook: (H, r, P-) of the latest measurementUpdate; returns rows, *n = cols
void orc_keep_last_update(void *h, int on) { ((Oracle *)h)->be.keep_last_update = on != 0; }
int orc_last_update(void *h, double *H, double *r, double *P, int cap, int *n) {
    Oracle *o = (Oracle *)h;
    const MsckfVio &b = o->be;
    *n = b.last_H.c;
    if (H && (int)b.last_H.d.size() <= cap) std::memcpy(H, b.last_H.d.data(), sizeof(double) * b.last_H.d.size());
    if (r) std::memcpy(r, b.last_r.d.data(), sizeof(double) * b.last_r.d.size());
    if (P && (int)b.last_P_prior.d.size() <= cap) std::memcpy(P, b.last_P_prior.d.data(), sizeof(double) * b.last_P_prior.d.size());
    return b.last_H.r;
}
// test hook: ids of the camera states behind the column groups 21 + 6 i of the latest measurementUpdate's H
int orc_last_update_cam_ids(void *h, long long *ids, int cap) {
    const MsckfVio &b = ((Oracle *)h)->be;
    const int n = (int)b.last_cam_ids.size();
    for (int i = 0; i < n && i < cap; ++i) ids[i] = b.last_cam_ids[i];
    return n;
}
// test hook: the feature map (ascending id): id, is_initialized, position, observation count
int orc_get_map(void *h, long long *ids, int *init, double *pos, int *nobs, int cap) {
    Oracle *o = (Oracle *)h;
    int n = 0;
    for (const auto &kv : o->be.map_server) {
        if (n < cap) {
            ids[n] = kv.first;
            init[n] = kv.second.is_initialized ? 1 : 0;
            for (int i = 0; i < 3; ++i) pos[n * 3 + i] = kv.second.position[i];
            nobs[n] = (int)kv.second.observations.size();
        }
        ++n;
    }
    return n;
}
// Feature::checkMotion + initializePosition (feature.hpp:257-450) on caller-supplied camera states
// (ascending state id 0..n_cam-1) and observations obs[n_feat][n_cam][4]; mask bit c = observed by c
void orc_triangulate(const mskf_config *cfg, int n_cam, const double *cam_q, const double *cam_p, int n_feat,
                     const unsigned *mask, const double *obs, double *pos, int *ok) {
    MsckfVio v(*cfg);
    CamStateServer cams;
    for (int c = 0; c < n_cam; ++c) {
        CAMState cs;
        cs.id = c;
        cs.orientation = Quat(cam_q[c * 4], cam_q[c * 4 + 1], cam_q[c * 4 + 2], cam_q[c * 4 + 3]);
        cs.position = V3(cam_p[c * 3], cam_p[c * 3 + 1], cam_p[c * 3 + 2]);
        cams[c] = cs;
    }
    for (int f = 0; f < n_feat; ++f) {
        Feature ft;
        ft.id = f;
        for (int c = 0; c < n_cam; ++c)
            if (mask[f] & (1u << c)) {
                Obs4 o;
                for (int k = 0; k < 4; ++k) o.v[k] = obs[((size_t)f * n_cam + c) * 4 + k];
                ft.observations[c] = o;
            }
        ok[f] = 0;
        if (ft.checkMotion(cams, v.opt_cfg)) ok[f] = ft.initializePosition(cams, v.T_cam0_cam1, v.opt_cfg) ? 1 : 0;
        for (int k = 0; k < 3; ++k) pos[f * 3 + k] = ft.position[k];
    }
}
double orc_chi2(const mskf_config *cfg, int dof) {
    MsckfVio v(*cfg);
    return v.chi2(dof);
}

}  // extern "C"
