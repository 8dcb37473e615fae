// backend.cu — MsckfVio hot path on sm_100a, fp64, batched over independent streams.
//
// Every stream owns one filter (IMU state, <= NS camera states, covariance P, feature map).
// The data-dependent control flow of the reference (which features are lost, how many
// observations each has, gating results, the 1500-row cap, the prune decision) lives in
// device memory: the host launches the same fixed kernel sequence every frame with
// worst-case grids, and CTAs whose stream (or feature) has no work exit at once.
//
// Layout choices (DESIGN.md "back end"):
//   * P is stored by physical camera SLOT: rows/cols [21+6*slot, +6).  Slots that hold no
//     camera state are kept exactly zero, so every product can run over the full LD x LD
//     matrix without masks, augmentation/pruning never shift memory (the reference
//     conservative_resizes and block-moves P every frame: msckf_vio.cpp:564-566,1161-1181),
//     and mskf_get_covariance gathers the logical (ascending state id) order on read-out.
//   * Observations are [feature slot][camera slot][4]; a 32-bit mask per feature says which
//     camera slots observe it.
//   * The Kalman update uses the compact-column, square-root form of the reference's
//     algebra: H only ever touches camera columns (msckf_vio.cpp:709-712), so with T the
//     (QR-compressed) measurement matrix over the k active camera columns,
//         S = T P_cc T^T + sigma^2 I = L L^T,  W = (P_:c T^T) L^-T,
//         P <- P - W W^T,  delta_x = P[:, c] (H^T r) / sigma^2   (K = P+ H^T / sigma^2)
//     which equals K = P H^T S^-1, P <- (I - K H) P, (P + P^T)/2 of msckf_vio.cpp:833-904 in
//     exact arithmetic and is symmetric by construction.  The posterior does not depend on
//     the orthonormal basis chosen for the null-space projection or on the QR row signs
//     (tests/test_oracle_backend.py::test_posterior_independent_of_nullspace_basis).
//
// Reference call sites replaced (msckf_core/src/msckf_vio.cpp, include/feature.hpp):
//   be_gravity_kernel        initializeGravityAndBias :209-241
//   be_propagate_kernel      batchImuProcessing :377-407, processModel :409-480, predictNewState :482-531
//   be_augment_kernel        stateAugmentation :533-585
//   be_add_obs_kernel        addFeatureObservations :587-608
//   be_select_kernel         removeLostFeatures :943-975 (selection), pruneCamStateBuffer :1073-1124,
//                            findRedundantCamStates :1026-1071
//   be_triangulate_kernel    Feature::checkMotion / initializePosition feature.hpp:257-450
//   be_layout_kernel         removeLostFeatures :965-990 bookkeeping
//   be_feature_jac_kernel    measurementJacobian :610-677, featureJacobian :679-775, gatingTest :909-935
//   be_stack_kernel          removeLostFeatures :992-1015 (stacking + row cap), pruneCamStateBuffer :1126-1150
//   be_gram_kernel, be_pchol_kernel   measurementUpdate :795-810 (SPQR compression, here via the Gram matrix)
//   be_gemm_kernel<...>, be_chol_kernel, be_apply_kernel   measurementUpdate :833-904
//   be_prune_finish_kernel   pruneCamStateBuffer :1161-1181
//   be_finish_kernel         publish :1238-1254, onlineReset :1186-1236
#include <math.h>
#include <unordered_set>
#include <stddef.h>
#include <string.h>

#include "chi2_table.h"
#include "common.cuh"

#define NSM 32          // camera-state slots (mask width)
#define BE_IMU_CAP 256  // IMU samples per stream per propagate launch
#define BE_THREADS 256

namespace mskf {

struct BeCam {
    long long id;
    double time;
    double q[4], p[3], qn[4], pn[3];
};

struct BeState {
    double time;
    long long id, next_id;
    double q[4], p[3], v[3], bg[3], ba[3];
    double Ric[9], tci[3];  // R_imu_cam0, t_cam0_imu
    double qn[4], pn[3], vn[3];
    double g[3];
    double tracking_rate;
    double T_b_w[16];
    long long n_updates, n_resets, n_overflow;
    int n_cam, cur_slot;
    unsigned cam_used;
    int order[NSM];  // logical (ascending state id) -> slot
    int n_feat;
    int gravity_set;
    // per-step scratch
    int n_list;          // length of the current feature list
    int n_todo;          // how many of them are not initialised yet (l_todo)
    int m, k, mt;        // stacked rows, compact columns, rows after compression
    int t_upper;         // Tm is upper triangular (the QR compression ran)
    int u_nslots;
    int u_slots[NSM];
    int colpos[NSM];
    int prune_active;
    int rm_slot[2];
    unsigned rm_bits;
    int do_update;
    int dbg_m[2], dbg_k[2], dbg_nlist[2];  // per phase: stacked rows, active columns, listed features (last step)
    // the latest measurementUpdate: stacked rows, active columns, whether its Gram matrix was formed (m > k) and
    // the camera-state id behind every group of six active columns (mskf_debug_last_gram)
    int gram_m, gram_k, gram_valid, gram_pad;
    long long gram_ids[NSM];
    BeCam cam[NSM];
};

struct BeStep {
    int active, first, n_imu, src;  // src 0: front-end message, 1: injected list
    int n_inject, pad;
    double t;
};

struct BeConst {
    int S, NS, LD, KC;  // KC = 6 NS
    int MF, HASH, ML;
    int ecap;      // elements of per-feature Jacobian scratch per stream
    int rcap;      // rows of per-feature residual scratch per stream
    int hst_cap;   // elements of the stacked H per stream
    int hst_rows;  // rows of the stacked residual per stream
    int ent_cap;   // message entries per stream
    int max_rows, max_cam, max_f;
    int chi2_mode;
    double gyro_noise, acc_noise, gyro_bias_noise, acc_bias_noise, obs_noise;
    double R01[9], t01[3];  // T_cam0_cam1 (config T_cn_cnm1), msckf_vio.cpp:118-121
    double Rib[9], tib[3];  // T_imu_body = inverse of the configured matrix, msckf_vio.cpp:124-126
    double pos_std_thr, rot_thr, trans_thr, track_thr, feat_trans_thr;
    double cov_gb, cov_v, cov_ab, cov_er, cov_et;
};

struct BeBuf {
    BeState *st;       // [S]
    BeStep *step;      // [S]
    double *imu;       // [S][BE_IMU_CAP][7]
    double *P;         // [S][LD*LD]
    unsigned *f_id;    // [S][MF]
    unsigned *f_mask;  // [S][MF]
    uint8_t *f_live, *f_init;  // [S][MF]
    double *f_pos;     // [S][MF][3]
    double *f_obs;     // [S][MF][NS][4]
    int *f_last;       // [S][MF]
    int *freelist;     // [S][MF]
    int *e_cell;       // [S][ent_cap+1]
    mskf_feature *inject;  // [ent_cap]
    // feature lists of the current phase
    int *l_slot;       // [S][ML]
    uint8_t *l_ok, *l_pass;  // [S][ML]
    int *l_todo;             // [S][ML] list indices of the features that still need a position (be_select -> be_triangulate)
    int *l_M, *l_eoff, *l_roff, *l_soff;  // [S][ML]
    uint8_t *l_oslots; // [S][ML][NSM]
    double *Hblk;      // [S][ecap] projected per-feature Jacobians (features that pass the gate)
    double *rblk;      // [S][rcap]
    double *Hst;       // [S][hst_cap]
    double *rst;       // [S][hst_rows]
    double *Gm;        // [S][(KC+1)^2] Gram matrix of [H | r], lower 64-tiles (be_gram_kernel)
    double *Rp;        // [S][KC*KC] rows of the pivoted Cholesky factor in the original column order
    int *perm;         // [S][KC] column order of Tm: compact column perm[l] of the stacked H is column l of Tm
    double *Tm;        // [S][KC*KC]
    double *gv;        // [S][KC] H^T r over the compact columns
    double *PHt;       // [S][LD*KC]
    double *Sm;        // [S][KC*KC]
    double *Linv;      // [S][KC*KC]
    double *W;         // [S][LD*KC]
    double *dxv;       // [S][LD] delta_x of the latest update
    double *work;      // [S][MSKF_PROF_TAGS] algorithmic flops done per kernel class (bench roofline)
    // front-end message (fb.stale / stale_hw / msg_total)
    const mskf_feature *fe_msg;
    const int *fe_hw;
    const long long *fe_total;
};

__constant__ double c_chi2[2][99];

// ======================================================================================
// small fp64 helpers (the conventions of oracle/kin.h: JPL quaternion [x y z w])
// ======================================================================================
__device__ __forceinline__ void quat_to_rot(const double q[4], double R[9]) {
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    const double a = 2 * w * w - 1, b = 2 * w;
    // (2w^2-1) I - 2w [q]x + 2 q q^T
    R[0] = a + 2 * x * x;         R[1] = b * z + 2 * x * y;     R[2] = -b * y + 2 * x * z;
    R[3] = -b * z + 2 * y * x;    R[4] = a + 2 * y * y;         R[5] = b * x + 2 * y * z;
    R[6] = b * y + 2 * z * x;     R[7] = -b * x + 2 * z * y;    R[8] = a + 2 * z * z;
}
__device__ __forceinline__ void quat_normalize(double q[4]) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
__device__ void rot_to_quat(const double R[9], double q[4]) {
    double tr = R[0] + R[4] + R[8];
    double score[4] = {R[0], R[4], R[8], tr};
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (score[i] > score[best]) best = i;
    if (best == 0) {
        q[0] = sqrt(1 + 2 * R[0] - tr) / 2.0;
        q[1] = (R[1] + R[3]) / (4 * q[0]);
        q[2] = (R[2] + R[6]) / (4 * q[0]);
        q[3] = (R[5] - R[7]) / (4 * q[0]);
    } else if (best == 1) {
        q[1] = sqrt(1 + 2 * R[4] - tr) / 2.0;
        q[0] = (R[1] + R[3]) / (4 * q[1]);
        q[2] = (R[5] + R[7]) / (4 * q[1]);
        q[3] = (R[6] - R[2]) / (4 * q[1]);
    } else if (best == 2) {
        q[2] = sqrt(1 + 2 * R[8] - tr) / 2.0;
        q[0] = (R[2] + R[6]) / (4 * q[2]);
        q[1] = (R[5] + R[7]) / (4 * q[2]);
        q[3] = (R[1] - R[3]) / (4 * q[2]);
    } else {
        q[3] = sqrt(1 + tr) / 2.0;
        q[0] = (R[5] - R[7]) / (4 * q[3]);
        q[1] = (R[6] - R[2]) / (4 * q[3]);
        q[2] = (R[1] - R[3]) / (4 * q[3]);
    }
    if (q[3] < 0)
        for (int i = 0; i < 4; ++i) q[i] = -q[i];
    quat_normalize(q);
}
__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double r[4]) {
    r[0] = a[3] * b[0] + a[2] * b[1] - a[1] * b[2] + a[0] * b[3];
    r[1] = -a[2] * b[0] + a[3] * b[1] + a[0] * b[2] + a[1] * b[3];
    r[2] = a[1] * b[0] - a[0] * b[1] + a[3] * b[2] + a[2] * b[3];
    r[3] = -a[0] * b[0] - a[1] * b[1] - a[2] * b[2] + a[3] * b[3];
    quat_normalize(r);
}
__device__ __forceinline__ void small_angle_quat(const double dth[3], double q[4]) {
    double d0 = dth[0] / 2.0, d1 = dth[1] / 2.0, d2 = dth[2] / 2.0;
    double n2 = d0 * d0 + d1 * d1 + d2 * d2;
    if (n2 <= 1) {
        q[0] = d0; q[1] = d1; q[2] = d2; q[3] = sqrt(1 - n2);
    } else {
        double s = sqrt(1 + n2);
        q[0] = d0 / s; q[1] = d1 / s; q[2] = d2 / s; q[3] = 1.0 / s;
    }
}
__device__ __forceinline__ void skew3(const double w[3], double K[9]) {
    K[0] = 0; K[1] = -w[2]; K[2] = w[1];
    K[3] = w[2]; K[4] = 0; K[5] = -w[0];
    K[6] = -w[1]; K[7] = w[0]; K[8] = 0;
}
__device__ __forceinline__ void m3mul(const double *a, const double *b, double *c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j] + a[i * 3 + 1] * b[3 + j] + a[i * 3 + 2] * b[6 + j];
}
__device__ __forceinline__ void m3mulT(const double *a, const double *b, double *c) {  // a * b^T
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i * 3] * b[j * 3] + a[i * 3 + 1] * b[j * 3 + 1] + a[i * 3 + 2] * b[j * 3 + 2];
}
__device__ __forceinline__ void m3Tmul(const double *a, const double *b, double *c) {  // a^T * b
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = a[i] * b[j] + a[3 + i] * b[3 + j] + a[6 + i] * b[6 + j];
}
__device__ __forceinline__ void m3v(const double *a, const double *x, double *y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = a[i * 3] * x[0] + a[i * 3 + 1] * x[1] + a[i * 3 + 2] * x[2];
}
__device__ __forceinline__ void m3Tv(const double *a, const double *x, double *y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = a[i] * x[0] + a[3 + i] * x[1] + a[6 + i] * x[2];
}
// angle of Eigen::AngleAxisd(R) (msckf_vio.cpp:1054) via the Hamilton quaternion
__device__ double rotation_angle(const double R[9]) {
    double tr = R[0] + R[4] + R[8];
    double x, y, z, w;
    if (tr > 0) {
        double s = sqrt(tr + 1.0) * 2;
        w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s;
    } else {
        double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s;
    }
    double n = sqrt(x * x + y * y + z * z);
    return 2.0 * atan2(n, fabs(w));
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// exclusive block scan of one int per thread (BE_THREADS threads); returns the prefix, *total = sum
__device__ int block_excl_scan(int v, int *total, int *s_tmp /* [BE_THREADS/32 + 1] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) s_tmp[warp] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int w = 0; w < BE_THREADS / 32; ++w) {
            int t = s_tmp[w];
            s_tmp[w] = acc;
            acc += t;
        }
        s_tmp[BE_THREADS / 32] = acc;
    }
    __syncthreads();
    *total = s_tmp[BE_THREADS / 32];
    return s_tmp[warp] + x - v;
}

__device__ __forceinline__ void reset_cov(const BeConst &bc, double *P) {  // msckf_vio.cpp:102-112
    for (int i = threadIdx.x; i < bc.LD * bc.LD; i += blockDim.x) P[i] = 0.0;
    __syncthreads();
    if (threadIdx.x < 21) {
        int i = threadIdx.x;
        double v = 0.0;
        if (i >= 3 && i < 6) v = bc.cov_gb;
        else if (i >= 6 && i < 9) v = bc.cov_v;
        else if (i >= 9 && i < 12) v = bc.cov_ab;
        else if (i >= 15 && i < 18) v = bc.cov_er;
        else if (i >= 18) v = bc.cov_et;
        P[i * bc.LD + i] = v;
    }
}

// ======================================================================================
// initializeGravityAndBias (msckf_vio.cpp:209-241): sums in buffer order, one thread
// ======================================================================================
__global__ void be_gravity_kernel(BeConst bc, BeBuf bb, int s, int n) {
    if (threadIdx.x != 0) return;
    BeState &st = bb.st[s];
    const double *imu = bb.imu + (size_t)s * BE_IMU_CAP * 7;
    double sw[3] = {0, 0, 0}, sa[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int k = 0; k < 3; ++k) {
            sw[k] = sw[k] + imu[i * 7 + 1 + k];
            sa[k] = sa[k] + imu[i * 7 + 4 + k];
        }
    double gi[3];
    for (int k = 0; k < 3; ++k) {
        st.bg[k] = sw[k] / (double)n;
        gi[k] = sa[k] / (double)n;
    }
    double gn = sqrt(gi[0] * gi[0] + gi[1] * gi[1] + gi[2] * gi[2]);
    st.g[0] = 0.0; st.g[1] = 0.0; st.g[2] = -gn;
    // from_two_vector(gravity_imu, -gravity) (Eigen FromTwoVectors semantics), transposed
    double a[3] = {gi[0] / gn, gi[1] / gn, gi[2] / gn};
    double b[3] = {0.0, 0.0, 1.0};
    double c = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
    double R[9];
    if (c < -1 + 1e-12) {
        double ax[3] = {fabs(a[0]) < 0.9 ? 1.0 : 0.0, fabs(a[0]) < 0.9 ? 0.0 : 1.0, 0.0};
        double kx[3] = {a[1] * ax[2] - a[2] * ax[1], a[2] * ax[0] - a[0] * ax[2], a[0] * ax[1] - a[1] * ax[0]};
        double kn = sqrt(kx[0] * kx[0] + kx[1] * kx[1] + kx[2] * kx[2]);
        for (int i = 0; i < 3; ++i) kx[i] /= kn;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i * 3 + j] = kx[i] * kx[j] * 2.0 - (i == j ? 1.0 : 0.0);
    } else {
        double v[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        double K[9], KK[9];
        skew3(v, K);
        m3mul(K, K, KK);
        double f = 1.0 / (1.0 + c);
        for (int i = 0; i < 9; ++i) R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + K[i] + KK[i] * f;
    }
    double Rt[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = R[j * 3 + i];
    rot_to_quat(Rt, st.q);
    st.gravity_set = 1;
}

// ======================================================================================
// batchImuProcessing / processModel / predictNewState.  One CTA per stream.  P11 and the
// 21x21 work matrices live in shared memory; the per-sample transition matrices are
// accumulated (Phi_tot = Phi_k ... Phi_1) and applied once to the IMU-camera blocks, which
// the reference touches (and re-symmetrises) once per IMU sample.
// ======================================================================================
#define N21 21
__device__ __forceinline__ void mm21(double *C, const double *A, const double *B) {  // C = A B
    for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) {
        int i = e / N21, j = e - i * N21;
        double s = 0;
#pragma unroll
        for (int k = 0; k < N21; ++k) s += A[i * N21 + k] * B[k * N21 + j];
        C[e] = s;
    }
}

__device__ void predict_new_state(BeState &st, double dt, const double gyro[3], const double acc[3]) {
    double gn = sqrt(gyro[0] * gyro[0] + gyro[1] * gyro[1] + gyro[2] * gyro[2]);
    double Om[16];
    for (int i = 0; i < 16; ++i) Om[i] = 0.0;
    double sk[9];
    skew3(gyro, sk);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) Om[i * 4 + j] = -sk[i * 3 + j];
        Om[i * 4 + 3] = gyro[i];
        Om[12 + i] = -gyro[i];
    }
    double *q = st.q, *v = st.v, *p = st.p;
    double dq[4], dq2[4];
    double cI, cO, post, cI2, cO2, post2;
    if (gn > 1e-5) {
        cI = cos(gn * dt * 0.5); cO = 1 / gn * sin(gn * dt * 0.5); post = 1.0;
        cI2 = cos(gn * dt * 0.25); cO2 = 1 / gn * sin(gn * dt * 0.25); post2 = 1.0;
    } else {
        cI = 1.0; cO = 0.5 * dt; post = cos(gn * dt * 0.5);
        cI2 = 1.0; cO2 = 0.25 * dt; post2 = cos(gn * dt * 0.25);
    }
    for (int i = 0; i < 4; ++i) {
        double s1 = 0, s2 = 0;
        for (int j = 0; j < 4; ++j) {
            s1 += ((i == j ? cI : 0.0) + cO * Om[i * 4 + j]) * post * q[j];
            s2 += ((i == j ? cI2 : 0.0) + cO2 * Om[i * 4 + j]) * post2 * q[j];
        }
        dq[i] = s1;
        dq2[i] = s2;
    }
    double Rq[9], Rd[9], Rd2[9];
    quat_to_rot(q, Rq);
    quat_to_rot(dq, Rd);
    quat_to_rot(dq2, Rd2);
    double k1v[3], k2v[3], k4v[3], t[3];
    m3Tv(Rq, acc, t);
    for (int i = 0; i < 3; ++i) k1v[i] = t[i] + st.g[i];
    m3Tv(Rd2, acc, t);
    for (int i = 0; i < 3; ++i) k2v[i] = t[i] + st.g[i];  // k3_v_dot == k2_v_dot (msckf_vio.cpp:516)
    m3Tv(Rd, acc, t);
    for (int i = 0; i < 3; ++i) k4v[i] = t[i] + st.g[i];
    for (int i = 0; i < 3; ++i) {
        double k1p = v[i];
        double k1_v = v[i] + k1v[i] * dt / 2;
        double k2p = k1_v;
        double k2_v = v[i] + k2v[i] * dt / 2;
        double k3p = k2_v;
        double k3_v = v[i] + k2v[i] * dt;
        double k4p = k3_v;
        double vn = v[i] + dt / 6 * (k1v[i] + 2 * k2v[i] + 2 * k2v[i] + k4v[i]);
        double pn = p[i] + dt / 6 * (k1p + 2 * k2p + 2 * k3p + k4p);
        v[i] = vn;
        p[i] = pn;
    }
    for (int i = 0; i < 4; ++i) q[i] = dq[i];
    quat_normalize(q);
}

__global__ void __launch_bounds__(128) be_propagate_kernel(BeConst bc, BeBuf bb, int chunk) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    double *P = bb.P + (size_t)s * bc.LD * bc.LD;
    const int LD = bc.LD;
    __shared__ double F[N21 * N21], A[N21 * N21], B[N21 * N21], Phi[N21 * N21], PhiT[N21 * N21], P11[N21 * N21],
        T[N21 * N21];
    __shared__ double s_dt;
    const int n0 = chunk * BE_IMU_CAP;
    const int n_imu = min(sp.n_imu - n0, BE_IMU_CAP);
    if (chunk == 0 && sp.first && threadIdx.x == 0) st.time = sp.t;  // msckf_vio.cpp:313-316
    if (n_imu <= 0) return;
    const double *imu = bb.imu + (size_t)s * BE_IMU_CAP * 7;
    for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) {
        int i = e / N21, j = e - i * N21;
        P11[e] = P[i * LD + j];
        PhiT[e] = (i == j) ? 1.0 : 0.0;
    }
    __syncthreads();
    for (int n = 0; n < n_imu; ++n) {
        for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) F[e] = 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            const double *m = imu + n * 7;
            double gyro[3], acc[3];
            for (int i = 0; i < 3; ++i) {
                gyro[i] = m[1 + i] - st.bg[i];
                acc[i] = m[4 + i] - st.ba[i];
            }
            double dt = m[0] - st.time;
            s_dt = dt;
            double R[9], sk[9], ska[9], RtSa[9];
            quat_to_rot(st.q, R);
            skew3(gyro, sk);
            skew3(acc, ska);
            m3Tmul(R, ska, RtSa);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    F[i * N21 + j] = -sk[i * 3 + j];
                    F[(6 + i) * N21 + j] = -RtSa[i * 3 + j];
                    F[(6 + i) * N21 + 9 + j] = -R[j * 3 + i];
                }
            for (int i = 0; i < 3; ++i) {
                F[i * N21 + 3 + i] = -1.0;
                F[(12 + i) * N21 + 6 + i] = 1.0;
            }
        }
        __syncthreads();
        const double dt = s_dt;
        for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) A[e] = F[e] * dt;
        __syncthreads();
        mm21(B, A, A);
        __syncthreads();
        mm21(T, B, A);
        __syncthreads();
        for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) {
            int i = e / N21, j = e - i * N21;
            Phi[e] = (i == j ? 1.0 : 0.0) + A[e] + 0.5 * B[e] + (1.0 / 6.0) * T[e];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const double *m = imu + n * 7;
            double gyro[3], acc[3];
            for (int i = 0; i < 3; ++i) {
                gyro[i] = m[1 + i] - st.bg[i];
                acc[i] = m[4 + i] - st.ba[i];
            }
            predict_new_state(st, dt, gyro, acc);
            // observability-constrained modification of Phi (msckf_vio.cpp:441-455)
            double Rkk1[9], Rq[9], B00[9];
            quat_to_rot(st.qn, Rkk1);
            quat_to_rot(st.q, Rq);
            m3mulT(Rq, Rkk1, B00);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Phi[i * N21 + j] = B00[i * 3 + j];
            double u[3], sv[3];
            m3v(Rkk1, st.g, u);
            double uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
            for (int i = 0; i < 3; ++i) sv[i] = (1.0 / uu) * u[i];
            for (int blk = 0; blk < 2; ++blk) {
                const int r0 = blk == 0 ? 6 : 12;
                double A1[9], d[3], w[3], K[9], au[3];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) A1[i * 3 + j] = Phi[(r0 + i) * N21 + j];
                if (blk == 0)
                    for (int i = 0; i < 3; ++i) d[i] = st.vn[i] - st.v[i];
                else
                    for (int i = 0; i < 3; ++i) d[i] = dt * st.vn[i] + st.pn[i] - st.p[i];
                skew3(d, K);
                m3v(K, st.g, w);
                m3v(A1, u, au);
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Phi[(r0 + i) * N21 + j] = A1[i * 3 + j] - (au[i] - w[i]) * sv[j];
            }
            for (int i = 0; i < 4; ++i) st.qn[i] = st.q[i];
            for (int i = 0; i < 3; ++i) {
                st.pn[i] = st.p[i];
                st.vn[i] = st.v[i];
            }
            st.time = m[0];
        }
        __syncthreads();
        // P11 <- Phi P11 Phi^T + Phi G Qc G^T Phi^T dt, symmetrised (msckf_vio.cpp:457-469)
        mm21(T, Phi, P11);
        __syncthreads();
        for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) {
            int i = e / N21, j = e - i * N21;
            double a = 0, q = 0;
#pragma unroll
            for (int k = 0; k < N21; ++k) a += T[i * N21 + k] * Phi[j * N21 + k];
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                double d = k < 3 ? bc.gyro_noise : (k < 6 ? bc.gyro_bias_noise : (k < 9 ? bc.acc_noise : bc.acc_bias_noise));
                q += Phi[i * N21 + k] * d * Phi[j * N21 + k];
            }
            A[e] = a + q * dt;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) {
            int i = e / N21, j = e - i * N21;
            P11[e] = (A[e] + A[j * N21 + i]) * 0.5;
        }
        mm21(T, Phi, PhiT);
        __syncthreads();
        for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) PhiT[e] = T[e];
        __syncthreads();
    }
    for (int e = threadIdx.x; e < N21 * N21; e += blockDim.x) {
        int i = e / N21, j = e - i * N21;
        P[i * LD + j] = P11[e];
    }
    // P12 <- Phi_tot P12, P21 <- its transpose (msckf_vio.cpp:461-466)
    for (int c = N21 + threadIdx.x; c < LD; c += blockDim.x) {
        double x[N21];
#pragma unroll
        for (int k = 0; k < N21; ++k) x[k] = P[k * LD + c];
#pragma unroll 1
        for (int i = 0; i < N21; ++i) {
            double y = 0;
#pragma unroll
            for (int k = 0; k < N21; ++k) y += PhiT[i * N21 + k] * x[k];
            P[i * LD + c] = y;
            P[c * LD + i] = y;
        }
    }
}

// ======================================================================================
// stateAugmentation.  One CTA per stream.
// ======================================================================================
__global__ void __launch_bounds__(BE_THREADS) be_augment_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    double *P = bb.P + (size_t)s * bc.LD * bc.LD;
    const int LD = bc.LD;
    __shared__ double J[6 * N21];
    __shared__ int s_slot;
    for (int e = threadIdx.x; e < 6 * N21; e += blockDim.x) J[e] = 0.0;
    __syncthreads();
    if (threadIdx.x == 0) {
        st.id = st.next_id++;  // batchImuProcessing, msckf_vio.cpp:401
        int slot = __ffs(~st.cam_used) - 1;
        if (slot < 0 || slot >= bc.NS) slot = bc.NS - 1;  // cannot happen: pruning keeps n_cam < NS here
        s_slot = slot;
        st.cam_used |= 1u << slot;
        st.order[st.n_cam++] = slot;
        st.cur_slot = slot;
        double Rwi[9], Rwc[9], tcw[3], v[3];
        quat_to_rot(st.q, Rwi);
        m3mul(st.Ric, Rwi, Rwc);
        m3Tv(Rwi, st.tci, v);
        for (int i = 0; i < 3; ++i) tcw[i] = st.p[i] + v[i];
        BeCam &c = st.cam[slot];
        c.id = st.id;
        c.time = sp.t;
        rot_to_quat(Rwc, c.q);
        for (int i = 0; i < 3; ++i) c.p[i] = tcw[i];
        for (int i = 0; i < 4; ++i) c.qn[i] = c.q[i];
        for (int i = 0; i < 3; ++i) c.pn[i] = c.p[i];
        double K[9];
        skew3(v, K);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                J[i * N21 + j] = st.Ric[i * 3 + j];
                J[(3 + i) * N21 + j] = K[i * 3 + j];
            }
            J[i * N21 + 15 + i] = 1.0;
            J[(3 + i) * N21 + 12 + i] = 1.0;
            J[(3 + i) * N21 + 18 + i] = 1.0;
        }
    }
    __syncthreads();
    const int rb = N21 + 6 * s_slot;
    for (int c = threadIdx.x; c < LD; c += blockDim.x) {
        if (c >= rb && c < rb + 6) continue;
        double x[N21];
#pragma unroll
        for (int k = 0; k < N21; ++k) x[k] = P[k * LD + c];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double y = 0;
#pragma unroll
            for (int k = 0; k < N21; ++k) y += J[i * N21 + k] * x[k];
            P[(rb + i) * LD + c] = y;
            P[c * LD + rb + i] = y;
        }
    }
    __syncthreads();
    __shared__ double C[36];
    if (threadIdx.x < 36) {
        int i = threadIdx.x / 6, j = threadIdx.x % 6;
        double y = 0;
        for (int k = 0; k < N21; ++k) y += P[(rb + i) * LD + k] * J[j * N21 + k];
        C[threadIdx.x] = y;
    }
    __syncthreads();
    if (threadIdx.x < 36) {
        int i = threadIdx.x / 6, j = threadIdx.x % 6;
        P[(rb + i) * LD + rb + j] = (C[i * 6 + j] + C[j * 6 + i]) * 0.5;
    }
}

// ======================================================================================
// addFeatureObservations.  One CTA per stream; open-addressing hash of the live feature ids
// in shared memory; message entries are applied with last-write-wins semantics in message
// order, exactly like the sequential map inserts of the reference (this matters for the
// stale tail of SURVEY F4, where one id can occur twice in a message).
// ======================================================================================
__device__ __forceinline__ unsigned hash_u32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(BE_THREADS) be_add_obs_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    extern __shared__ __align__(16) unsigned char be_smem[];
    unsigned *h_key = (unsigned *)be_smem;           // [HASH] id + 1, 0 = empty
    int *h_slot = (int *)(h_key + bc.HASH);          // [HASH]
    int *h_owner = h_slot + bc.HASH;                 // [HASH]
    __shared__ int s_tmp[BE_THREADS / 32 + 1];
    __shared__ int s_nfree, s_over;
    const size_t fo = (size_t)s * bc.MF;
    unsigned *f_id = bb.f_id + fo, *f_mask = bb.f_mask + fo;
    uint8_t *f_live = bb.f_live + fo, *f_init = bb.f_init + fo;
    int *f_last = bb.f_last + fo, *freelist = bb.freelist + fo;
    int *e_cell = bb.e_cell + (size_t)s * (bc.ent_cap + 1);
    const unsigned hmask = (unsigned)bc.HASH - 1u;
    const mskf_feature *ent;
    int n_real;
    long long n_zero;
    if (sp.src == 0) {
        ent = bb.fe_msg + (size_t)s * bc.max_f;
        n_real = bb.fe_hw[s];
        n_zero = bb.fe_total[s] - (long long)n_real;
    } else {
        ent = bb.inject;
        n_real = sp.n_inject;
        n_zero = 0;
    }
    const int n_ent = n_real + (n_zero > 0 ? 1 : 0);  // the value-initialised tail acts as one entry {id 0, zeros}
    for (int c = threadIdx.x; c < bc.HASH; c += BE_THREADS) {
        h_key[c] = 0u;
        h_slot[c] = -1;
        h_owner[c] = 0x7fffffff;
    }
    if (threadIdx.x == 0) { s_nfree = 0; s_over = 0; }
    __syncthreads();
    // live features -> hash; free slots -> ordered free list
    int base = 0;
    for (int start = 0; start < bc.MF; start += BE_THREADS) {
        int slot = start + threadIdx.x;
        int is_free = 0;
        if (slot < bc.MF) {
            if (f_live[slot]) {
                unsigned key = f_id[slot] + 1u;
                unsigned c = hash_u32(f_id[slot]) & hmask;
                while (atomicCAS(&h_key[c], 0u, key) != 0u) c = (c + 1u) & hmask;
                h_slot[c] = slot;
                f_last[slot] = -1;
            } else {
                is_free = 1;
            }
        }
        int tot;
        int pos = block_excl_scan(is_free, &tot, s_tmp);
        if (is_free) freelist[base + pos] = slot;
        base += tot;
    }
    if (threadIdx.x == 0) s_nfree = base;
    __syncthreads();
    // find-or-insert every entry's id
    for (int i = threadIdx.x; i < n_ent; i += BE_THREADS) {
        unsigned id = i < n_real ? ent[i].id : 0u;
        unsigned key = id + 1u;
        unsigned c = hash_u32(id) & hmask;
        int probes = 0;
        for (; probes < bc.HASH; ++probes) {  // bounded: a full table must not spin (be_step also rejects such messages)
            unsigned old = atomicCAS(&h_key[c], 0u, key);
            if (old == 0u || old == key) break;
            c = (c + 1u) & hmask;
        }
        if (probes == bc.HASH) {
            e_cell[i] = -1;  // no cell: the entry is dropped and counted as overflow
            s_over = 1;
            continue;
        }
        e_cell[i] = (int)c;
        if (h_slot[c] < 0) atomicMin(&h_owner[c], i);
    }
    __syncthreads();
    // new ids get slots in entry order (deterministic)
    base = 0;
    for (int start = 0; start < n_ent; start += BE_THREADS) {
        int i = start + threadIdx.x;
        int own = 0, c = 0;
        if (i < n_ent && e_cell[i] >= 0) {
            c = e_cell[i];
            own = (h_slot[c] < 0 && h_owner[c] == i) ? 1 : 0;
        }
        int tot;
        int pos = block_excl_scan(own, &tot, s_tmp);
        if (own) {
            int rank = base + pos;
            if (rank < s_nfree) {
                int slot = freelist[rank];
                f_id[slot] = i < n_real ? ent[i].id : 0u;
                f_mask[slot] = 0u;
                f_init[slot] = 0;
                f_live[slot] = 1;
                f_last[slot] = -1;
                h_owner[c] = -1 - slot;  // published after the barrier below
            } else {
                s_over = 1;
            }
        }
        base += tot;
    }
    __syncthreads();
    const int n_new = base;
    for (int c = threadIdx.x; c < bc.HASH; c += BE_THREADS)
        if (h_slot[c] < 0 && h_owner[c] < 0) h_slot[c] = -1 - h_owner[c];
    __syncthreads();
    for (int i = threadIdx.x; i < n_ent; i += BE_THREADS) {
        int slot = e_cell[i] >= 0 ? h_slot[e_cell[i]] : -1;
        if (slot >= 0) atomicMax(&f_last[slot], i);
    }
    __syncthreads();
    const unsigned curbit = 1u << st.cur_slot;
    for (int i = threadIdx.x; i < n_ent; i += BE_THREADS) {
        int slot = e_cell[i] >= 0 ? h_slot[e_cell[i]] : -1;
        if (slot < 0 || f_last[slot] != i) continue;
        double *o = bb.f_obs + (((size_t)s * bc.MF + slot) * bc.NS + st.cur_slot) * 4;
        if (i < n_real) {
            o[0] = ent[i].u0; o[1] = ent[i].v0; o[2] = ent[i].u1; o[3] = ent[i].v1;
        } else {
            o[0] = 0.0; o[1] = 0.0; o[2] = 0.0; o[3] = 0.0;
        }
        f_mask[slot] |= curbit;
    }
    if (threadIdx.x == 0) {
        const int curr = st.n_feat;
        const int added = s_over ? min(n_new, s_nfree) : n_new;
        const long long tracked = (long long)n_real + n_zero - (long long)n_new;
        st.n_feat = curr + added;
        st.tracking_rate = (double)tracked / (double)curr;  // msckf_vio.cpp:604-606 (0/0 -> NaN as in the reference)
        if (s_over) st.n_overflow++;
    }
}

// ======================================================================================
// Feature selection for the two update phases.  phase 0: removeLostFeatures, phase 1:
// pruneCamStateBuffer.  The list is sorted by feature id (std::map order).
// ======================================================================================
__device__ void find_redundant(const BeConst &bc, BeState &st) {  // msckf_vio.cpp:1026-1071
    int n = st.n_cam;
    int key_i = n - 4, it = n - 3, first = 0;
    const BeCam &key = st.cam[st.order[key_i]];
    double Rk[9];
    quat_to_rot(key.q, Rk);
    long long rid[2];
    int rslot[2];
    for (int i = 0; i < 2; ++i) {
        const BeCam &c = st.cam[st.order[it]];
        double R[9], RR[9];
        quat_to_rot(c.q, R);
        double d[3] = {c.p[0] - key.p[0], c.p[1] - key.p[1], c.p[2] - key.p[2]};
        double dist = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        m3mulT(R, Rk, RR);
        double ang = rotation_angle(RR);
        if (ang < bc.rot_thr && dist < bc.trans_thr && st.tracking_rate > bc.track_thr) {
            rslot[i] = st.order[it];
            rid[i] = c.id;
            ++it;
        } else {
            rslot[i] = st.order[first];
            rid[i] = st.cam[st.order[first]].id;
            ++first;
        }
    }
    if (rid[1] < rid[0]) {
        int t = rslot[0];
        rslot[0] = rslot[1];
        rslot[1] = t;
    }
    st.rm_slot[0] = rslot[0];
    st.rm_slot[1] = rslot[1];
    st.rm_bits = (1u << rslot[0]) | (1u << rslot[1]);
}

__global__ void __launch_bounds__(BE_THREADS) be_select_kernel(BeConst bc, BeBuf bb, int phase, int sort_n) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    extern __shared__ __align__(16) unsigned char be_smem[];
    unsigned long long *keys = (unsigned long long *)be_smem;  // [sort_n]
    __shared__ int s_n, s_erased;
    const size_t fo = (size_t)s * bc.MF;
    unsigned *f_mask = bb.f_mask + fo;
    if (threadIdx.x == 0) {
        s_n = 0;
        s_erased = 0;
        st.do_update = 0;
        st.m = 0;
        if (phase == 1) {
            st.prune_active = st.n_cam >= bc.max_cam ? 1 : 0;
            if (st.prune_active) find_redundant(bc, st);
        }
    }
    for (int i = threadIdx.x; i < sort_n; i += BE_THREADS) keys[i] = ~0ull;
    __syncthreads();
    if (phase == 1 && !st.prune_active) {
        if (threadIdx.x == 0) st.n_list = 0;
        return;
    }
    const unsigned curbit = 1u << st.cur_slot, rmbits = st.rm_bits;
    for (int slot = threadIdx.x; slot < bc.MF; slot += BE_THREADS) {
        if (!bb.f_live[fo + slot]) continue;
        unsigned mask = f_mask[slot];
        if (phase == 0) {
            if (mask & curbit) continue;
            if (__popc(mask) < 3) {  // msckf_vio.cpp:956-959
                bb.f_live[fo + slot] = 0;
                atomicAdd(&s_erased, 1);
                continue;
            }
        } else {
            unsigned inv = mask & rmbits;
            int c = __popc(inv);
            if (c == 0) continue;
            if (c == 1) {  // msckf_vio.cpp:1096-1099
                f_mask[slot] = mask & ~inv;
                continue;
            }
        }
        int pos = atomicAdd(&s_n, 1);
        if (pos < sort_n) keys[pos] = ((unsigned long long)bb.f_id[fo + slot] << 32) | (unsigned)slot;
    }
    __syncthreads();
    {
        int need = 2;
        while (need < s_n && need < sort_n) need <<= 1;
        sort_n = need;  // keys beyond the candidates are ~0 and sort to the end
    }
    // bitonic sort ascending
    for (int k = 2; k <= sort_n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < sort_n; i += BE_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    unsigned long long a = keys[i], b = keys[ixj];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    const int n = min(s_n, min(sort_n, bc.ML));
    __syncthreads();
    // Most listed features are initialised already (the prune phase lists ~260 per stream, a handful of which
    // still need their position): the ones that do are compacted HERE, from flags that be_triangulate_kernel does
    // not write while it reads them, so that its warps take one each instead of queueing behind each other in
    // list order (a triangulation is a serial Levenberg-Marquardt chain of ~50 us)
    __shared__ int s_ntodo;
    if (threadIdx.x == 0) s_ntodo = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += BE_THREADS) {
        const int slot = (int)(keys[i] & 0xffffffffu);
        bb.l_slot[(size_t)s * bc.ML + i] = slot;
        if (bb.f_init[fo + slot]) bb.l_ok[(size_t)s * bc.ML + i] = 1;
        else bb.l_todo[(size_t)s * bc.ML + atomicAdd(&s_ntodo, 1)] = i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st.n_list = n;
        st.n_todo = s_ntodo;
        st.n_feat -= s_erased;
    }
}

// ======================================================================================
// Feature::checkMotion + initializePosition (feature.hpp:257-450).  One warp per listed
// feature: lane l keeps the poses (relative to the first observing camera) of stereo views l
// and l + 32 in registers; all lanes run the same LM control flow.
//
// The LM step is accepted iff new_cost < total_cost (feature.hpp:417), and at convergence
// that comparison is decided by the last bits of a 2M-term sum, so this kernel reproduces the
// reference's arithmetic exactly: every operation is a separately rounded IEEE fp64 operation
// (type `sd`: __dadd_rn / __dmul_rn / __ddiv_rn / __dsqrt_rn are never contracted into FMAs;
// the rest of this file is compiled with contraction on), in the expression order of
// feature.hpp / oracle/backend.h, and the cost and normal-equation sums run over the views in
// the reference's sequential order (per-view terms go through shared memory; one lane per
// matrix entry adds them up in view order).  Given the same camera states and observations the
// triangulated position equals the oracle's bit for bit (tests/test_gpu_backend.py).
// ======================================================================================
struct sd {
    double v;
    __device__ __forceinline__ sd() {}
    __device__ __forceinline__ sd(double x) : v(x) {}
};
__device__ __forceinline__ sd operator+(sd a, sd b) { return sd(__dadd_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a, sd b) { return sd(__dsub_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator*(sd a, sd b) { return sd(__dmul_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator/(sd a, sd b) { return sd(__ddiv_rn(a.v, b.v)); }
__device__ __forceinline__ sd operator-(sd a) { return sd(-a.v); }
__device__ __forceinline__ sd ssqrt(sd a) { return sd(__dsqrt_rn(a.v)); }

struct Pose { sd R[9], t[3]; };

// oracle/linalg.h operator*(M3, M3): s = 0; s += a(i,k) b(k,j)
__device__ __forceinline__ void s_m3mul(const sd *a, const sd *b, sd *c) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) c[i * 3 + j] = ((sd(0.0) + a[i * 3] * b[j]) + a[i * 3 + 1] * b[3 + j]) + a[i * 3 + 2] * b[6 + j];
}
// operator*(M3, V3)
__device__ __forceinline__ void s_m3v(const sd *a, const sd *x, sd *y) {
#pragma unroll
    for (int i = 0; i < 3; ++i) y[i] = (a[i * 3] * x[0] + a[i * 3 + 1] * x[1]) + a[i * 3 + 2] * x[2];
}
// quat_to_rot(q).t(): (2w^2-1) I - 2w [q]x + 2 q q^T, element by element as oracle/kin.h, transposed
__device__ __forceinline__ void s_quat_to_rot_t(const double q[4], sd Rt[9]) {
    const sd w = q[3];
    const sd a = sd(2.0) * w * w - sd(1.0), b = sd(2.0) * w;
    const double K[9] = {0.0, -q[2], q[1], q[2], 0.0, -q[0], -q[1], q[0], 0.0};
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            sd e = (sd(i == j ? 1.0 : 0.0) * a - sd(K[i * 3 + j]) * b) + (sd(q[i]) * sd(q[j])) * sd(2.0);
            Rt[j * 3 + i] = e;
        }
}
// SE3::inv(): R^T, -(R^T t)
__device__ __forceinline__ void s_pose_inv(const Pose &p, Pose &o) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) o.R[i * 3 + j] = p.R[j * 3 + i];
    sd tt[3];
    s_m3v(o.R, p.t, tt);
#pragma unroll
    for (int i = 0; i < 3; ++i) o.t[i] = -tt[i];
}
// operator*(SE3, SE3): (a.R b.R, a.R b.t + a.t)
__device__ __forceinline__ void s_pose_mul(const Pose &a, const Pose &b, Pose &o) {
    s_m3mul(a.R, b.R, o.R);
    sd tt[3];
    s_m3v(a.R, b.t, tt);
#pragma unroll
    for (int i = 0; i < 3; ++i) o.t[i] = tt[i] + a.t[i];
}
__device__ __forceinline__ void cam_pose_world(const BeCam &c, int cam1, const BeConst &bc, Pose &o) {
    // cam0_pose = (R(q)^T, p); cam1_pose = cam0_pose * T_cam0_cam1.inv()  (feature.hpp:307-318)
    Pose p0;
    s_quat_to_rot_t(c.q, p0.R);
#pragma unroll
    for (int i = 0; i < 3; ++i) p0.t[i] = c.p[i];
    if (!cam1) {
        o = p0;
    } else {
        Pose T01, Ti;
#pragma unroll
        for (int i = 0; i < 9; ++i) T01.R[i] = bc.R01[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) T01.t[i] = bc.t01[i];
        s_pose_inv(T01, Ti);
        s_pose_mul(p0, Ti, o);
    }
}
__device__ __forceinline__ void pose_rel(const Pose &pose, const Pose &Tc0w, Pose &o) {  // pose.inv() * T_c0_w
    Pose pi;
    s_pose_inv(pose, pi);
    s_pose_mul(pi, Tc0w, o);
}
// Feature::cost (feature.hpp:171-190)
__device__ __forceinline__ sd lm_cost(const Pose &T, const sd x[3], const sd z[2]) {
    sd h0 = ((T.R[0] * x[0] + T.R[1] * x[1]) + T.R[2] * sd(1.0)) + T.t[0] * x[2];
    sd h1 = ((T.R[3] * x[0] + T.R[4] * x[1]) + T.R[5] * sd(1.0)) + T.t[1] * x[2];
    sd h2 = ((T.R[6] * x[0] + T.R[7] * x[1]) + T.R[8] * sd(1.0)) + T.t[2] * x[2];
    sd zx = h0 / h2 - z[0], zy = h1 / h2 - z[1];
    return zx * zx + zy * zy;
}
// Feature::jacobian (feature.hpp:192-229) and this view's terms of A = sum w^2 J^T J (lower triangle
// first: 00 10 11 20 21 22) and b = sum w^2 J^T r (feature.hpp:372-386)
__device__ __forceinline__ void lm_terms(const Pose &T, const sd x[3], const sd z[2], sd huber, double *out /* [9] */) {
    sd h1 = ((T.R[0] * x[0] + T.R[1] * x[1]) + T.R[2] * sd(1.0)) + T.t[0] * x[2];
    sd h2 = ((T.R[3] * x[0] + T.R[4] * x[1]) + T.R[5] * sd(1.0)) + T.t[1] * x[2];
    sd h3 = ((T.R[6] * x[0] + T.R[7] * x[1]) + T.R[8] * sd(1.0)) + T.t[2] * x[2];
    const sd W[9] = {T.R[0], T.R[1], T.t[0], T.R[3], T.R[4], T.t[1], T.R[6], T.R[7], T.t[2]};
    sd J[6], r[2];
    const sd ih3 = sd(1.0) / h3, a1 = h1 / (h3 * h3), a2 = h2 / (h3 * h3);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        J[j] = ih3 * W[j] - a1 * W[6 + j];
        J[3 + j] = ih3 * W[3 + j] - a2 * W[6 + j];
    }
    r[0] = h1 / h3 - z[0];
    r[1] = h2 / h3 - z[1];
    sd e = ssqrt(r[0] * r[0] + r[1] * r[1]);
    const bool unit = e.v <= huber.v;
    sd w = unit ? sd(1.0) : ssqrt((sd(2.0) * huber) / e);
    const bool w1 = (w.v == 1.0);
    sd ws = w * w;
    int o = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int c = 0; c <= a; ++c) {
            sd jtj = J[a] * J[c] + J[3 + a] * J[3 + c];
            out[o++] = (w1 ? jtj : ws * jtj).v;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        sd jtr = J[a] * r[0] + J[3 + a] * r[1];
        out[6 + a] = (w1 ? jtr : ws * jtr).v;
    }
}
// Eigen Matrix3d::ldlt().solve (feature.hpp:395) as oracle/linalg.h ldlt_solve for n = 3;
// S given by its lower triangle 00 10 11 20 21 22
__device__ __forceinline__ void ldlt3_solve(const sd S[6], const sd b[3], sd x[3]) {
    sd D0 = S[0];
    sd L10 = S[1] / D0, L20 = S[3] / D0;
    sd D1 = S[2] - L10 * L10 * D0;
    sd L21 = (S[4] - L20 * L10 * D0) / D1;
    sd D2 = (S[5] - L20 * L20 * D0) - L21 * L21 * D1;
    sd y0 = b[0], y1 = b[1], y2 = b[2];
    if (L10.v != 0.0) y1 = y1 - L10 * y0;
    if (L20.v != 0.0) y2 = y2 - L20 * y0;
    if (L21.v != 0.0) y2 = y2 - L21 * y1;
    y0 = y0 / D0; y1 = y1 / D1; y2 = y2 / D2;
    x[2] = y2;
    x[1] = y1;
    if (L21.v != 0.0) x[1] = x[1] - L21 * x[2];
    x[0] = y0;
    if (L10.v != 0.0) x[0] = x[0] - L10 * x[1];
    if (L20.v != 0.0) x[0] = x[0] - L20 * x[2];
}

#define TRI_WARPS 4
#define TRI_VIEWS 64  // 2 * NSM stereo views
struct TriShared {
    double term[TRI_VIEWS][9];  // per-view terms of A (lower triangle) and b
    double cost[TRI_VIEWS];
};
__device__ void triangulate_one(const BeConst &bc, const BeBuf &bb, const BeState &st, int s, int li, int lane, size_t fo, TriShared &sh);
__global__ void __launch_bounds__(TRI_WARPS * 32) be_triangulate_kernel(BeConst bc, BeBuf bb, int phase) {
    __shared__ TriShared s_tri[TRI_WARPS];
    const int s = blockIdx.y;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    const BeState &st = bb.st[s];
    const int lane = threadIdx.x & 31;
    const size_t fo = (size_t)s * bc.MF;
    // the features that still need a position, compacted by be_select_kernel: one per warp
    const int n_todo = st.n_todo;
    const int *todo = bb.l_todo + (size_t)s * bc.ML;
    for (int j = blockIdx.x * TRI_WARPS + (threadIdx.x >> 5); j < n_todo; j += gridDim.x * TRI_WARPS)
        triangulate_one(bc, bb, st, s, todo[j], lane, fo, s_tri[threadIdx.x >> 5]);
}

// sum of the np per-view costs in view order, starting from 0.0 (feature.hpp:359-364, 402-407); every lane
// computes the same value
__device__ __forceinline__ sd tri_cost_sum(TriShared &sh, int np, int lane, const bool have[2], const sd c[2]) {
    __syncwarp();
    if (have[0]) sh.cost[lane] = c[0].v;
    if (have[1]) sh.cost[lane + 32] = c[1].v;
    __syncwarp();
    sd t(0.0);
    for (int i = 0; i < np; ++i) t = t + sd(sh.cost[i]);
    return t;
}

__device__ void triangulate_one(const BeConst &bc, const BeBuf &bb, const BeState &st, int s, int li, int lane, size_t fo, TriShared &sh) {
    const int slot = bb.l_slot[(size_t)s * bc.ML + li];
    uint8_t *ok = bb.l_ok + (size_t)s * bc.ML + li;
    if (bb.f_init[fo + slot]) {
        if (lane == 0) *ok = 1;
        return;
    }
    const unsigned mask = bb.f_mask[fo + slot];
    // observing camera slots in ascending state id order
    int first_slot = -1, last_slot = -1, M = 0;
    int my_slot[2] = {-1, -1};
    for (int i = 0; i < st.n_cam; ++i) {
        int cs = st.order[i];
        if (!(mask & (1u << cs))) continue;
        if (first_slot < 0) first_slot = cs;
        last_slot = cs;
        if (2 * M == lane || 2 * M + 1 == lane) my_slot[0] = cs;
        if (2 * M == lane + 32 || 2 * M + 1 == lane + 32) my_slot[1] = cs;
        ++M;
    }
    const double *obs = bb.f_obs + ((size_t)s * bc.MF + slot) * bc.NS * 4;
    // ---- checkMotion (feature.hpp:257-287)
    {
        const BeCam &c0 = st.cam[first_slot], &c1 = st.cam[last_slot];
        sd R0[9];  // first_pose.R = R(q)^T
        s_quat_to_rot_t(c0.q, R0);
        const double *o0 = obs + first_slot * 4;
        sd dir[3] = {o0[0], o0[1], 1.0};
        sd dn = ssqrt((dir[0] * dir[0] + dir[1] * dir[1]) + dir[2] * dir[2]);
        for (int i = 0; i < 3; ++i) dir[i] = dir[i] / dn;
        sd dw[3];
        s_m3v(R0, dir, dw);
        sd tr[3] = {sd(c1.p[0]) - sd(c0.p[0]), sd(c1.p[1]) - sd(c0.p[1]), sd(c1.p[2]) - sd(c0.p[2])};
        sd par = (tr[0] * dw[0] + tr[1] * dw[1]) + tr[2] * dw[2];
        sd orth[3] = {tr[0] - par * dw[0], tr[1] - par * dw[1], tr[2] - par * dw[2]};
        sd on = ssqrt((orth[0] * orth[0] + orth[1] * orth[1]) + orth[2] * orth[2]);
        if (!(on.v > bc.feat_trans_thr)) {
            if (lane == 0) *ok = 0;
            return;
        }
    }
    // ---- initializePosition (feature.hpp:289-450)
    Pose Tc0w;
    cam_pose_world(st.cam[first_slot], 0, bc, Tc0w);
    const int np = 2 * M;
    Pose my[2];
    sd mz[2][2];
    const bool have[2] = {lane < np, lane + 32 < np};
    for (int h = 0; h < 2; ++h) {
        if (!have[h]) continue;
        int j = lane + 32 * h;
        Pose w;
        cam_pose_world(st.cam[my_slot[h]], j & 1, bc, w);
        pose_rel(w, Tc0w, my[h]);
        const double *o = obs + my_slot[h] * 4 + 2 * (j & 1);
        mz[h][0] = o[0];
        mz[h][1] = o[1];
    }
    sd sol[3];
    {   // generateInitialGuess (feature.hpp:231-255) from the first and the last view
        Pose wl, Tl;
        cam_pose_world(st.cam[last_slot], 1, bc, wl);
        pose_rel(wl, Tc0w, Tl);
        const double *z1 = obs + first_slot * 4, *z2 = obs + last_slot * 4 + 2;
        sd m[3], zz[3] = {z1[0], z1[1], 1.0};
        s_m3v(Tl.R, zz, m);
        sd A0 = m[0] - sd(z2[0]) * m[2], A1 = m[1] - sd(z2[1]) * m[2];
        sd b0 = sd(z2[0]) * Tl.t[2] - Tl.t[0], b1 = sd(z2[1]) * Tl.t[2] - Tl.t[1];
        sd depth = (sd(1.0) / (A0 * A0 + A1 * A1)) * (A0 * b0 + A1 * b1);
        sd ip[3] = {sd(z1[0]) * depth, sd(z1[1]) * depth, depth};
        sol[0] = ip[0] / ip[2];
        sol[1] = ip[1] / ip[2];
        sol[2] = sd(1.0) / ip[2];
    }
    const sd huber(0.01);
    const double est_prec = 5e-7;
    sd lambda(1e-3);
    int inner = 0, outer = 0;
    bool reduced = false;
    sd delta_norm(0.0);
    sd total_cost;
    {
        sd c[2];
        for (int h = 0; h < 2; ++h)
            if (have[h]) c[h] = lm_cost(my[h], sol, mz[h]);
        total_cost = tri_cost_sum(sh, np, lane, have, c);
    }
    do {
        __syncwarp();
        for (int h = 0; h < 2; ++h)
            if (have[h]) lm_terms(my[h], sol, mz[h], huber, sh.term[lane + 32 * h]);
        __syncwarp();
        sd Ab[9];
        {
            sd acc(0.0);
            if (lane < 9)
                for (int i = 0; i < np; ++i) acc = acc + sd(sh.term[i][lane]);
#pragma unroll
            for (int e = 0; e < 9; ++e) Ab[e] = sd(__shfl_sync(0xffffffffu, acc.v, e));
        }
        do {
            sd At[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) At[i] = Ab[i];
            At[0] = At[0] + lambda; At[2] = At[2] + lambda; At[5] = At[5] + lambda;
            sd delta[3], ns[3];
            ldlt3_solve(At, Ab + 6, delta);
            for (int i = 0; i < 3; ++i) ns[i] = sol[i] - delta[i];
            delta_norm = ssqrt((delta[0] * delta[0] + delta[1] * delta[1]) + delta[2] * delta[2]);
            sd c[2];
            for (int h = 0; h < 2; ++h)
                if (have[h]) c[h] = lm_cost(my[h], ns, mz[h]);
            sd new_cost = tri_cost_sum(sh, np, lane, have, c);
            if (new_cost.v < total_cost.v) {
                reduced = true;
                for (int i = 0; i < 3; ++i) sol[i] = ns[i];
                total_cost = new_cost;
                sd l10 = lambda / sd(10.0);
                lambda = l10.v > 1e-10 ? l10 : sd(1e-10);
            } else {
                reduced = false;
                sd l10 = lambda * sd(10.0);
                lambda = l10.v < 1e12 ? l10 : sd(1e12);
            }
        } while (inner++ < 10 && !reduced);
        inner = 0;
    } while (outer++ < 10 && delta_norm.v > est_prec);
    sd fp[3] = {sol[0] / sol[2], sol[1] / sol[2], sd(1.0) / sol[2]};
    bool valid = true;
    for (int h = 0; h < 2; ++h)
        if (have[h]) {
            sd z = ((my[h].R[6] * fp[0] + my[h].R[7] * fp[1]) + my[h].R[8] * fp[2]) + my[h].t[2];
            if (z.v <= 0) valid = false;
        }
    valid = __all_sync(0xffffffffu, valid);
    if (lane == 0) {
        sd pw[3];
        s_m3v(Tc0w.R, fp, pw);
        for (int i = 0; i < 3; ++i) bb.f_pos[(fo + slot) * 3 + i] = (pw[i] + Tc0w.t[i]).v;
        if (valid) bb.f_init[fo + slot] = 1;
        *ok = valid ? 1 : 0;
    }
}

// ======================================================================================
// List compaction + scratch layout for the per-feature Jacobian blocks.
// ======================================================================================
__global__ void __launch_bounds__(BE_THREADS) be_layout_kernel(BeConst bc, BeBuf bb, int phase) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    __shared__ int s_tmp[BE_THREADS / 32 + 1];
    __shared__ int s_fit;
    const int n = st.n_list;
    if (n == 0) return;
    if (threadIdx.x == 0) s_fit = 0;
    __syncthreads();
    const size_t fo = (size_t)s * bc.MF, lo = (size_t)s * bc.ML;
    int base_n = 0, base_e = 0, base_r = 0, erased = 0, overflow = 0;
    for (int start = 0; start < n; start += BE_THREADS) {
        int i = start + threadIdx.x;
        int keep = 0, slot = 0, M = 0;
        if (i < n) {
            slot = bb.l_slot[lo + i];
            if (bb.l_ok[lo + i]) {
                keep = 1;
                M = phase == 0 ? __popc(bb.f_mask[fo + slot]) : 2;
            } else if (phase == 0) {
                bb.f_live[fo + slot] = 0;  // invalid_feature_ids, msckf_vio.cpp:961-972
            } else {
                bb.f_mask[fo + slot] &= ~st.rm_bits;  // msckf_vio.cpp:1102-1113
            }
        }
        int tot, tot_e, tot_r, tot_x;
        int pos = block_excl_scan(keep, &tot, s_tmp);
        int pe = block_excl_scan(keep ? 4 * M * 6 * M : 0, &tot_e, s_tmp);
        int pr = block_excl_scan(keep ? 4 * M : 0, &tot_r, s_tmp);
        block_excl_scan((i < n && !keep) ? 1 : 0, &tot_x, s_tmp);
        __syncthreads();  // every thread has read slot i before slot pos (<= i) is rewritten
        if (keep) {
            int eo = base_e + pe, ro = base_r + pr;
            if (eo + 4 * M * 6 * M <= bc.ecap && ro + 4 * M <= bc.rcap) {
                int d = base_n + pos;
                bb.l_slot[lo + d] = slot;
                bb.l_M[lo + d] = M;
                bb.l_eoff[lo + d] = eo;
                bb.l_roff[lo + d] = ro;
                atomicAdd(&s_fit, 1);
            } else {
                overflow = 1;
            }
        }
        __syncthreads();
        base_n += tot;
        base_e += tot_e;
        base_r += tot_r;
        erased += tot_x;
    }
    overflow = __syncthreads_or(overflow);
    if (threadIdx.x == 0) {
        // scratch offsets grow with the list index, so the blocks that fit are a prefix of the list;
        // anything beyond (never seen with the default capacities) is dropped and counted
        if (overflow) st.n_overflow++;
        st.n_list = s_fit;
        if (phase == 0) st.n_feat -= erased;
    }
}

// ======================================================================================
// CTA-level fp64 GEMM tile on the fp64 tensor pipe (mma.sync m8n8k4, SASS DMMA.8x8x4: one per 4
// cycles per SM = 37 TFLOP/s, tools/micro/dmma_rate.cu):  C (<= 64 x 64) = sum_l A(i, l) B(j, l).
//   * Operands go global -> shared with 8-byte cp.async (LDGSTS; rows of P start on odd multiples of
//     8 bytes, so 16-byte copies are not possible), two stages of 16 k: the copies of chunk c + 1 are in
//     flight while the DMMAs of chunk c run, and no register or thread is spent on staging.
//   * Shared-memory layout [row][k] with a row stride of 20 doubles (B may also be [k][col] with a
//     stride of 68): the 32 addresses of a fragment load (8 rows x 4 k) fall into 32 different 8-byte
//     banks-pairs, i.e. every LDS.64 is conflict-free.
//   * 8 warps as 2 x 4: a warp owns 32 rows x 16 columns = 4 x 2 DMMA tiles, so per k-step of 4 it
//     loads 4 A and 2 B fragments for 8 DMMAs (the first version loaded 9 for 8).
//   * Tiles are not a fixed 64: a dimension n is cut into ceil(n / 64) tiles of equal numbers of 8-row
//     blocks (201 rows -> 56, 48, 48, 49 instead of 64, 64, 64, 9), and DMMA tiles outside the matrix
//     are skipped (warp-uniform).
// ======================================================================================
#define GT 64
#define UG_K 16
#define UG_RS (UG_K + 4)
#define UG_RSB (GT + 4)
struct UpdGemmSmem {
    double a[2][GT * UG_RS];
    double b[2][GT * UG_RS];  // K-slow B needs UG_K * UG_RSB = 1088 <= 1280 doubles
};

__device__ __forceinline__ void dmma_884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
// 8-byte asynchronous copy global -> shared; !valid writes zeros (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async8(double *dst, const double *src, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(src), "r"(sz) : "memory");
}
// tile t of a dimension of n elements: [start, start + count), count <= 64, a multiple of 8 except at the end
__device__ __forceinline__ int ug_tiles(int n) { return ((n + 7) / 8 + 7) / 8; }
__device__ __forceinline__ void ug_range(int n, int t, int &start, int &count) {
    const int nb = (n + 7) / 8, nt = (nb + 7) / 8;
    const int b0 = t * nb / nt, b1 = (t + 1) * nb / nt;
    start = 8 * b0;
    count = min(n, 8 * b1) - start;
}

// acc += A B^T over k in [k_begin, k_end) for a tile of mr rows and nc columns.  pa(row, l) / pb(l, col) give
// the address of an operand element (only called for elements inside the tile and the k range); `dummy` is
// any valid address (used for the zero-filled copies).  B_KFAST: B's k index is the contiguous one.
template <bool B_KFAST, class PA, class PB>
__device__ __forceinline__ void upd_gemm_tile(int mr, int nc, int k_begin, int k_end, PA pa, PB pb, const double *dummy,
                                              UpdGemmSmem &sm, double (&acc)[4][2][2]) {
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const int wr = warp >> 2, wc = warp & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    k_begin &= ~(UG_K - 1);
    const int nchunks = (k_end - k_begin + UG_K - 1) / UG_K;
    if (nchunks <= 0) return;
    auto stage = [&](int buf, int k0) {
        // A (and a K-fast B): thread -> k offset tid & 15, rows (tid >> 4) + 16 u
        const int l = k0 + (tid & 15);
        const bool lv = l < k_end;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = (tid >> 4) + 16 * u;
            const bool v = lv && row < mr;
            cp_async8(&sm.a[buf][row * UG_RS + (tid & 15)], v ? pa(row, l) : dummy, v);
        }
        if (B_KFAST) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int col = (tid >> 4) + 16 * u;
                const bool v = lv && col < nc;
                cp_async8(&sm.b[buf][col * UG_RS + (tid & 15)], v ? pb(l, col) : dummy, v);
            }
        } else {
            // K-slow B: thread -> column tid & 63, k offsets (tid >> 6) + 4 u
            const int col = tid & 63;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int lo = (tid >> 6) + 4 * u, lb = k0 + lo;
                const bool v = lb < k_end && col < nc;
                cp_async8(&sm.b[buf][lo * UG_RSB + col], v ? pb(lb, col) : dummy, v);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    bool mv[4], nv[2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) mv[mi] = wr * 32 + mi * 8 < mr;
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) nv[ni] = wc * 16 + ni * 8 < nc;
    stage(0, k_begin);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            stage((c + 1) & 1, k_begin + (c + 1) * UG_K);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (mv[0] && nv[0]) {
            const double *as = sm.a[c & 1] + (wr * 32 + g) * UG_RS + t4;
            const double *bs = B_KFAST ? sm.b[c & 1] + (wc * 16 + g) * UG_RS + t4 : sm.b[c & 1] + t4 * UG_RSB + wc * 16 + g;
#pragma unroll
            for (int kk = 0; kk < UG_K; kk += 4) {
                double af[4], bf[2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) af[mi] = as[mi * 8 * UG_RS + kk];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) bf[ni] = B_KFAST ? bs[ni * 8 * UG_RS + kk] : bs[kk * UG_RSB + ni * 8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni)
                        if (mv[mi] && nv[ni]) dmma_884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
        }
        __syncthreads();  // everyone is done with buffer c & 1 before chunk c + 2 lands in it
    }
}
// f(row, col, value) for every accumulator element of this thread inside the tile
template <class F>
__device__ __forceinline__ void upd_gemm_store(int mr, int nc, const double (&acc)[4][2][2], F f) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3, wr = warp >> 2, wc = warp & 3;
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int row = wr * 32 + mi * 8 + g, col = wc * 16 + ni * 8 + 2 * t4 + h2;
                if (row < mr && col < nc) f(row, col, acc[mi][ni][h2]);
            }
}

// In-place Cholesky of an n x n matrix in shared memory (row-major, leading dimension ld,
// lower triangle), CTA-wide; then y <- L^-1 y.
__device__ void cta_cholesky(double *S, int n, int ld) {
    for (int k = 0; k < n; ++k) {
        __syncthreads();
        if (threadIdx.x == 0) S[k * ld + k] = sqrt(S[k * ld + k]);
        __syncthreads();
        const double d = S[k * ld + k];
        for (int i = k + 1 + threadIdx.x; i < n; i += blockDim.x) S[i * ld + k] /= d;
        __syncthreads();
        const int rem = n - k - 1;
        for (int e = threadIdx.x; e < rem * rem; e += blockDim.x) {
            int a = e / rem, b = e - a * rem;
            if (b > a) continue;
            int i = k + 1 + a, j = k + 1 + b;
            S[i * ld + j] -= S[i * ld + k] * S[j * ld + k];
        }
    }
    __syncthreads();
}

// 1 / sqrt(a) for a > 0 from the RSQ64H seed and two Newton steps (fp64 round-off)
__device__ __forceinline__ double chol_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double h = 0.5 * a;
    y = fma(y, fma(-h, y * y, 0.5), y);
    y = fma(y, fma(-h, y * y, 0.5), y);
    return y;
}

#define CH_T 6
__device__ __forceinline__ void chol_load6(const double *p, double (&v)[CH_T]) {  // p is 16-byte aligned (6-double chunks)
    const double2 *q = reinterpret_cast<const double2 *>(p);
    const double2 a = q[0], b = q[1], c = q[2];
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y;
}
__device__ __forceinline__ void chol_store6(double *p, const double (&v)[CH_T]) {
    double2 *q = reinterpret_cast<double2 *>(p);
    q[0] = make_double2(v[0], v[1]);
    q[1] = make_double2(v[2], v[3]);
    q[2] = make_double2(v[4], v[5]);
}

// Tile (ti, tj), tj <= ti, of thread t = ti (ti + 1) / 2 + tj of a lower tile triangle
__device__ __forceinline__ void chol_tile_of_thread(int t, int &ti, int &tj) {
    ti = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((ti + 1) * (ti + 2) / 2 <= t) ++ti;
    while (ti * (ti + 1) / 2 > t) --ti;
    tj = t - ti * (ti + 1) / 2;
}

// One 6x6 diagonal tile in registers: D = L L^T (lower triangle of g on entry), then g = X = L^-1 (lower; the upper
// part is set to zero).  A non-positive pivot propagates NaN, as sqrt would.  Serial: ~6 rsqrt chains.
__device__ __forceinline__ void chol_diag_tile_inverse(double (&g)[CH_T][CH_T]) {
    double inv[CH_T];
#pragma unroll
    for (int c = 0; c < CH_T; ++c) {
        const double piv = g[c][c];
        inv[c] = piv > 0.0 ? chol_rsqrt(piv) : __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
        for (int a = c + 1; a < CH_T; ++a) g[a][c] *= inv[c];
#pragma unroll
        for (int b = c + 1; b < CH_T; ++b)
#pragma unroll
            for (int a = b; a < CH_T; ++a) g[a][b] = fma(-g[a][c], g[b][c], g[a][b]);
    }
    // X = L^-1 by forward substitution on the identity, column by column; L[i][i] = 1 / inv[i]
    double x[CH_T][CH_T];
#pragma unroll
    for (int j = 0; j < CH_T; ++j) {
        x[j][j] = inv[j];
#pragma unroll
        for (int i = j + 1; i < CH_T; ++i) {
            double acc = 0.0;
#pragma unroll
            for (int m = j; m < i; ++m) acc = fma(g[i][m], x[m][j], acc);
            x[i][j] = -acc * inv[i];
        }
    }
#pragma unroll
    for (int a = 0; a < CH_T; ++a)
#pragma unroll
        for (int b = 0; b < CH_T; ++b) g[a][b] = b <= a ? x[a][b] : 0.0;
}

// gamma = r^T S^-1 r for the n x n SPD matrix S held as rows/columns [off, off + n) of a packed lower triangle in
// shared memory (entry (i, j), j <= i, at i (i + 1) / 2 + j), by a BLOCKED right-looking Cholesky with the whole
// triangle in REGISTERS: thread t owns the 6x6 tile (ti, tj) (231 tiles for n <= 126 on 256 threads).  Block step k:
//   1. the owner of the diagonal tile factors it and inverts the 6x6 factor in registers (X = L_kk^-1), takes
//      y_k = X r_k (forward substitution of the right-hand side) and publishes X and y_k;          -- barrier
//   2. the tiles below it become the panel C_i = G_ik X^T = L_ik, published column-major, and fold their share
//      into the right-hand side, r_i -= C_i y_k;                                                     -- barrier
//   3. every trailing tile takes G_ij -= C_i C_j^T (216 FMAs from two 36-double panel pieces).
// Two barriers per SIX columns; the column-by-column version on the packed triangle in shared memory took three
// barriers per column plus a warp-serial forward substitution afterwards (85 of the kernel's 120 us per 30-view
// feature).  The packed triangle is dead once the tiles are loaded; its storage holds the panel, r and X.
// All BE_THREADS threads must call it; the result is returned to every thread.
__device__ double tile_cholesky_gamma(double *G, int off, int n, const double *r, int ti, int tj) {
    const int nt = (n + CH_T - 1) / CH_T, NP = nt * CH_T;
    const int i0 = CH_T * ti, j0 = CH_T * tj;
    const bool live = ti < nt;
    double g[CH_T][CH_T];
#pragma unroll
    for (int a = 0; a < CH_T; ++a)
#pragma unroll
        for (int b = 0; b < CH_T; ++b) {
            const int i = i0 + a, j = j0 + b;
            // rows past n: identity, so that the padding factors to itself and adds nothing to gamma
            g[a][b] = (live && i < n && j <= i) ? G[(i + off) * (i + off + 1) / 2 + j + off] : ((i == j) ? 1.0 : 0.0);
        }
    __syncthreads();
    double *panel = G;                      // [6][NP]: column c of the current panel at panel + c NP
    double *rs = panel + CH_T * NP;         // [NP]
    double *s_X = rs + NP;                  // [6][6]
    double *s_y = s_X + CH_T * CH_T;        // [6]
    double *s_gamma = s_y + CH_T;           // [1]
    for (int i = threadIdx.x; i < NP; i += BE_THREADS) rs[i] = i < n ? r[i] : 0.0;
    if (threadIdx.x == 0) *s_gamma = 0.0;
    __syncthreads();
    for (int k = 0; k < nt; ++k) {
        if (ti == k && tj == k) {
            chol_diag_tile_inverse(g);
            double rk[CH_T], y[CH_T];
            chol_load6(rs + i0, rk);
            double gsum = 0.0;
#pragma unroll
            for (int a = 0; a < CH_T; ++a) {
                double acc = 0.0;
#pragma unroll
                for (int b = 0; b <= a; ++b) acc = fma(g[a][b], rk[b], acc);
                y[a] = acc;
                gsum = fma(acc, acc, gsum);
            }
            chol_store6(s_y, y);
#pragma unroll
            for (int a = 0; a < CH_T; ++a) chol_store6(s_X + CH_T * a, g[a]);
            *s_gamma += gsum;
        }
        if (k == nt - 1) break;
        __syncthreads();
        if (live && tj == k && ti > k) {
            double y[CH_T], ri[CH_T];
            chol_load6(s_y, y);
            chol_load6(rs + i0, ri);
#pragma unroll
            for (int cc = 0; cc < CH_T; ++cc) {
                double xr[CH_T], col[CH_T];
                chol_load6(s_X + CH_T * cc, xr);  // row cc of X: X[cc][b], b <= cc
#pragma unroll
                for (int a = 0; a < CH_T; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int b = 0; b <= cc; ++b) acc = fma(g[a][b], xr[b], acc);
                    col[a] = acc;
                    ri[a] = fma(-acc, y[cc], ri[a]);
                }
                chol_store6(panel + cc * NP + i0, col);
            }
            chol_store6(rs + i0, ri);
        }
        __syncthreads();
        if (live && tj > k) {
#pragma unroll
            for (int cc = 0; cc < CH_T; ++cc) {
                double ci[CH_T], cj[CH_T];
                chol_load6(panel + cc * NP + i0, ci);
                chol_load6(panel + cc * NP + j0, cj);
#pragma unroll
                for (int a = 0; a < CH_T; ++a)
#pragma unroll
                    for (int b = 0; b < CH_T; ++b) g[a][b] = fma(-ci[a], cj[b], g[a][b]);
            }
        }
    }
    __syncthreads();
    return *s_gamma;
}

// ======================================================================================
// measurementJacobian + featureJacobian + gatingTest for one feature.  One CTA per listed
// feature.  The stacked Jacobian H_xj of a feature is block diagonal over its observing
// cameras (one 4x6 block per camera), which the kernel exploits instead of forming the dense
// products of the reference:
//   * Q = H0 H1 H2 = I - V T V^T are the three Householder reflectors of H_fj (4M x 3); the
//     projected Jacobian is H' = (Q^T H_xj)[3:, :], column c needs V^T x_c with x_c 4-sparse;
//   * gating needs S = H' P H'^T + sigma^2 I = (Q^T G Q)[3:, 3:] + sigma^2 I with
//     G = H_xj P_sub H_xj^T assembled from M(M+1)/2 blocks Hx_a P_ab Hx_b^T (4x4 each), and
//     Q^T G Q = G - V A^T - A V^T + V C V^T,  A = (G V) T,  C = T^T (V^T G V) T.
// G / S live in shared memory as a packed lower triangle; H' is written to the scratch only
// for features that pass the gate.
// ======================================================================================
__global__ void __launch_bounds__(BE_THREADS, 2) be_feature_jac_kernel(BeConst bc, BeBuf bb, int phase, int maxM) {
    const int s = blockIdx.y;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    const BeState &st = bb.st[s];
    const int n_list = st.n_list;
    const size_t fo = (size_t)s * bc.MF, lo = (size_t)s * bc.ML;
    const double *P = bb.P + (size_t)s * bc.LD * bc.LD;
    const int LD = bc.LD;

    extern __shared__ __align__(16) unsigned char be_smem[];
    const int maxR4 = 4 * maxM;
    double *G = (double *)be_smem;                       // packed lower [maxR4 (maxR4 + 1) / 2]
    double *Hx = G + (size_t)maxR4 * (maxR4 + 1) / 2;    // [maxM][4][6]
    double *Hf = Hx + maxM * 24;                         // [maxR4][3]   V below the diagonal after the QR
    double *rv = Hf + maxR4 * 3;                         // [maxR4]
    double *Y = rv + maxR4;                              // [maxR4][3]
    double *Am = Y + maxR4 * 3;                          // [maxR4][3]
    double *U = Am + maxR4 * 3;                          // [6 maxM][3]
    __shared__ int oslot[NSM];
    __shared__ double tau[3], Tm[9], Cm[9], Zm[9], ur[3];
    __shared__ double s_gamma;
    __shared__ int s_pass;
    auto gix = [](int i, int j) { return i * (i + 1) / 2 + j; };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if ((int)blockIdx.x >= n_list) return;
    int ch_ti, ch_tj;
    chol_tile_of_thread(threadIdx.x, ch_ti, ch_tj);

    for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
        __syncthreads();
        const int slot = bb.l_slot[lo + li], M = bb.l_M[lo + li];
        const int R4 = 4 * M, C6 = 6 * M, rows = R4 - 3;
        double *H = bb.Hblk + (size_t)s * bc.ecap + bb.l_eoff[lo + li];   // [R4][C6], rows 3.. used
        double *rg = bb.rblk + (size_t)s * bc.rcap + bb.l_roff[lo + li];  // [R4]
        if (threadIdx.x == 0) {
            if (phase == 0) {
                const unsigned mask = bb.f_mask[fo + slot];
                int m = 0;
                for (int i = 0; i < st.n_cam; ++i) {
                    int cs = st.order[i];
                    if (mask & (1u << cs)) oslot[m++] = cs;
                }
            } else {
                oslot[0] = st.rm_slot[0];
                oslot[1] = st.rm_slot[1];
            }
            for (int i = 0; i < M; ++i) bb.l_oslots[(lo + li) * NSM + i] = (uint8_t)oslot[i];
        }
        __syncthreads();
        if (threadIdx.x < M) {
            // measurementJacobian, msckf_vio.cpp:610-677
            const int t = threadIdx.x, cs = oslot[t];
            const BeCam &c = st.cam[cs];
            const double *z = bb.f_obs + (((size_t)s * bc.MF + slot) * bc.NS + cs) * 4;
            const double *pw = bb.f_pos + (fo + slot) * 3;
            double Rw0[9], Rw1[9], t1w[3], tmp[3];
            quat_to_rot(c.q, Rw0);
            m3mul(bc.R01, Rw0, Rw1);
            m3Tv(Rw1, bc.t01, tmp);
            for (int i = 0; i < 3; ++i) t1w[i] = c.p[i] - tmp[i];
            double d0[3] = {pw[0] - c.p[0], pw[1] - c.p[1], pw[2] - c.p[2]};
            double d1[3] = {pw[0] - t1w[0], pw[1] - t1w[1], pw[2] - t1w[2]};
            double p0[3], p1[3];
            m3v(Rw0, d0, p0);
            m3v(Rw1, d1, p1);
            double dz0[4][3] = {{1 / p0[2], 0, -p0[0] / (p0[2] * p0[2])}, {0, 1 / p0[2], -p0[1] / (p0[2] * p0[2])}, {0, 0, 0}, {0, 0, 0}};
            double dz1[4][3] = {{0, 0, 0}, {0, 0, 0}, {1 / p1[2], 0, -p1[0] / (p1[2] * p1[2])}, {0, 1 / p1[2], -p1[1] / (p1[2] * p1[2])}};
            double sk0[9], R01sk[9];
            skew3(p0, sk0);
            m3mul(bc.R01, sk0, R01sk);
            double Hxl[4][6];
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 3; ++j) {
                    double a = 0, b = 0;
                    for (int k = 0; k < 3; ++k) {
                        a += dz0[i][k] * sk0[k * 3 + j] + dz1[i][k] * R01sk[k * 3 + j];
                        b += dz0[i][k] * (-Rw0[k * 3 + j]) + dz1[i][k] * (-Rw1[k * 3 + j]);
                    }
                    Hxl[i][j] = a;
                    Hxl[i][3 + j] = b;
                }
            // observability projection: H_x <- A - A u (u^T u)^-1 u^T, H_f <- -H_x[:, 3:6]
            double u[6], Rn[9], dn[3] = {pw[0] - c.pn[0], pw[1] - c.pn[1], pw[2] - c.pn[2]}, K[9];
            quat_to_rot(c.qn, Rn);
            m3v(Rn, st.g, u);
            skew3(dn, K);
            m3v(K, st.g, u + 3);
            double utu = 0;
            for (int i = 0; i < 6; ++i) utu += u[i] * u[i];
            for (int i = 0; i < 4; ++i) {
                double au = 0;
                for (int k = 0; k < 6; ++k) au += Hxl[i][k] * u[k];
                au *= (1.0 / utu);
                for (int k = 0; k < 6; ++k) Hx[(t * 4 + i) * 6 + k] = Hxl[i][k] - au * u[k];
                for (int k = 0; k < 3; ++k) Hf[(4 * t + i) * 3 + k] = -(Hxl[i][3 + k] - au * u[3 + k]);
            }
            rv[4 * t + 0] = z[0] - p0[0] / p0[2];
            rv[4 * t + 1] = z[1] - p0[1] / p0[2];
            rv[4 * t + 2] = z[2] - p1[0] / p1[2];
            rv[4 * t + 3] = z[3] - p1[1] / p1[2];
        }
        __syncthreads();
        // ---- warp 0: Householder QR of H_f, T factor, Q^T r.  Other warps: G blocks.
        if (warp == 0) {
            for (int j = 0; j < 3; ++j) {
                double xn = 0;
                for (int i = j + 1 + lane; i < R4; i += 32) xn += Hf[i * 3 + j] * Hf[i * 3 + j];
                xn = warp_sum_d(xn);
                const double alpha = Hf[j * 3 + j];
                double tj = 0.0;
                if (xn != 0.0) {
                    double beta = -copysign(sqrt(alpha * alpha + xn), alpha);
                    tj = (beta - alpha) / beta;
                    double scale = 1.0 / (alpha - beta);
                    __syncwarp();
                    for (int i = j + 1 + lane; i < R4; i += 32) Hf[i * 3 + j] *= scale;
                    if (lane == 0) Hf[j * 3 + j] = beta;
                    __syncwarp();
                    for (int c = j + 1; c < 3; ++c) {
                        double sacc = 0;
                        for (int i = j + 1 + lane; i < R4; i += 32) sacc += Hf[i * 3 + j] * Hf[i * 3 + c];
                        sacc = warp_sum_d(sacc) + Hf[j * 3 + c];
                        sacc *= tj;
                        __syncwarp();
                        if (lane == 0) Hf[j * 3 + c] -= sacc;
                        for (int i = j + 1 + lane; i < R4; i += 32) Hf[i * 3 + c] -= sacc * Hf[i * 3 + j];
                        __syncwarp();
                    }
                } else {
                    // zero reflector: make the stored column an explicit zero vector below the diagonal
                    for (int i = j + 1 + lane; i < R4; i += 32) Hf[i * 3 + j] = 0.0;
                }
                if (lane == 0) tau[j] = tj;
                __syncwarp();
            }
            // overwrite the R part with the implicit unit-lower structure of V: V(i, j) = 0 (i < j), 1 (i == j)
            if (lane == 0) {
                Hf[0 * 3 + 0] = 1.0; Hf[0 * 3 + 1] = 0.0; Hf[0 * 3 + 2] = 0.0;
                Hf[1 * 3 + 1] = 1.0; Hf[1 * 3 + 2] = 0.0;
                Hf[2 * 3 + 2] = 1.0;
            }
            __syncwarp();
            double d01 = 0, d02 = 0, d12 = 0, w0 = 0, w1 = 0, w2 = 0;
            for (int i = lane; i < R4; i += 32) {
                const double a0 = Hf[i * 3], a1 = Hf[i * 3 + 1], a2 = Hf[i * 3 + 2], ri = rv[i];
                d01 += a0 * a1; d02 += a0 * a2; d12 += a1 * a2;
                w0 += a0 * ri; w1 += a1 * ri; w2 += a2 * ri;
            }
            d01 = warp_sum_d(d01); d02 = warp_sum_d(d02); d12 = warp_sum_d(d12);
            w0 = warp_sum_d(w0); w1 = warp_sum_d(w1); w2 = warp_sum_d(w2);
            if (lane == 0) {
                const double t00 = tau[0], t11 = tau[1], t22 = tau[2];
                const double t01 = -t11 * (t00 * d01);
                const double t02 = -t22 * (t00 * d02 + t01 * d12);
                const double t12 = -t22 * (t11 * d12);
                Tm[0] = t00; Tm[1] = t01; Tm[2] = t02;
                Tm[3] = 0.0; Tm[4] = t11; Tm[5] = t12;
                Tm[6] = 0.0; Tm[7] = 0.0; Tm[8] = t22;
                // Q^T r = r - V (T^T (V^T r))
                ur[0] = t00 * w0;
                ur[1] = t01 * w0 + t11 * w1;
                ur[2] = t02 * w0 + t12 * w1 + t22 * w2;
            }
        } else {
            // G = H_xj P_sub H_xj^T, blocks (a >= b), one thread per block
            const int nblk = M * (M + 1) / 2;
            for (int e = threadIdx.x - 32; e < nblk; e += BE_THREADS - 32) {
                int a = 0, t = e;
                while (t > a) { t -= a + 1; ++a; }
                const int b = t;
                const double *Pa = P + (size_t)(N21 + 6 * oslot[a]) * LD + N21 + 6 * oslot[b];
                const double *Ha = Hx + a * 24, *Hb = Hx + b * 24;
                double HP[4][6];
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    double p[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) p[k] = Pa[(size_t)k * LD + j];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double acc = 0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc += Ha[i * 6 + k] * p[k];
                        HP[i][j] = acc;
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (a == b && j > i) continue;
                        double acc = 0;
#pragma unroll
                        for (int k = 0; k < 6; ++k) acc += HP[i][k] * Hb[j * 6 + k];
                        G[gix(4 * a + i, 4 * b + j)] = acc;
                    }
            }
        }
        __syncthreads();
        // diagonal blocks: symmetrise exactly what the two triangles would have held (a == b blocks are
        // computed once, lower part only) -- nothing to do; Y = G V
        for (int e = threadIdx.x; e < R4 * 3; e += BE_THREADS) {
            const int i = e / 3, a = e - i * 3;
            double acc = 0;
            for (int j = 0; j <= i; ++j) acc += G[gix(i, j)] * Hf[j * 3 + a];
            for (int j = i + 1; j < R4; ++j) acc += G[gix(j, i)] * Hf[j * 3 + a];
            Y[e] = acc;
        }
        // U[c] = T^T (V^T x_c) for the 6M sparse columns of H_xj
        for (int c = threadIdx.x; c < C6; c += BE_THREADS) {
            const int t = c / 6, cc = c - t * 6;
            double w0 = 0, w1 = 0, w2 = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double x = Hx[(t * 4 + i) * 6 + cc];
                w0 += Hf[(4 * t + i) * 3] * x;
                w1 += Hf[(4 * t + i) * 3 + 1] * x;
                w2 += Hf[(4 * t + i) * 3 + 2] * x;
            }
            U[c * 3 + 0] = Tm[0] * w0;
            U[c * 3 + 1] = Tm[1] * w0 + Tm[4] * w1;
            U[c * 3 + 2] = Tm[2] * w0 + Tm[5] * w1 + Tm[8] * w2;
        }
        __syncthreads();
        if (threadIdx.x < 9) {  // Z = V^T Y
            const int a = threadIdx.x / 3, b = threadIdx.x % 3;
            double acc = 0;
            for (int i = 0; i < R4; ++i) acc += Hf[i * 3 + a] * Y[i * 3 + b];
            Zm[threadIdx.x] = acc;
        }
        for (int e = threadIdx.x; e < R4 * 3; e += BE_THREADS) {  // A = Y T
            const int i = e / 3, b = e - i * 3;
            double acc = 0;
            for (int a = 0; a <= b; ++a) acc += Y[i * 3 + a] * Tm[a * 3 + b];
            Am[e] = acc;
        }
        for (int i = threadIdx.x; i < R4; i += BE_THREADS)  // r' = Q^T r
            rv[i] = rv[i] - (Hf[i * 3] * ur[0] + Hf[i * 3 + 1] * ur[1] + Hf[i * 3 + 2] * ur[2]);
        __syncthreads();
        if (threadIdx.x < 9) {  // C = T^T Z T
            const int a = threadIdx.x / 3, b = threadIdx.x % 3;
            double acc = 0;
            for (int p = 0; p < 3; ++p)
                for (int q = 0; q < 3; ++q) acc += Tm[p * 3 + a] * Zm[p * 3 + q] * Tm[q * 3 + b];
            Cm[threadIdx.x] = acc;
        }
        __syncthreads();
        // S = (Q^T G Q)[3:, 3:] + sigma^2 I, in place on the packed lower triangle
        {
            const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
            for (int i = 3 + ty; i < R4; i += 16) {
                const double vi0 = Hf[i * 3], vi1 = Hf[i * 3 + 1], vi2 = Hf[i * 3 + 2];
                const double ai0 = Am[i * 3], ai1 = Am[i * 3 + 1], ai2 = Am[i * 3 + 2];
                const double c0 = vi0 * Cm[0] + vi1 * Cm[3] + vi2 * Cm[6], c1 = vi0 * Cm[1] + vi1 * Cm[4] + vi2 * Cm[7],
                             c2 = vi0 * Cm[2] + vi1 * Cm[5] + vi2 * Cm[8];
                const int ib = i * (i + 1) / 2;
                for (int j = 3 + tx; j <= i; j += 16) {
                    const double vj0 = Hf[j * 3], vj1 = Hf[j * 3 + 1], vj2 = Hf[j * 3 + 2];
                    double x = G[ib + j];
                    x -= vi0 * Am[j * 3] + vi1 * Am[j * 3 + 1] + vi2 * Am[j * 3 + 2];
                    x -= ai0 * vj0 + ai1 * vj1 + ai2 * vj2;
                    x += c0 * vj0 + c1 * vj1 + c2 * vj2;
                    if (i == j) x += bc.obs_noise;
                    G[ib + j] = x;
                }
            }
        }
        __syncthreads();
        // gamma = r'^T S^-1 r' (gatingTest, msckf_vio.cpp:909-935): blocked Cholesky of S in registers with the
        // forward substitution of r' folded in
        const double g = tile_cholesky_gamma(G, 3, rows, rv + 3, ch_ti, ch_tj);
        if (warp == 0) {
            if (lane == 0) {
                const int dof = phase == 0 ? M - 1 : M;  // msckf_vio.cpp:1001, :1145
                const double thr = (dof >= 1 && dof <= 99) ? c_chi2[bc.chi2_mode][dof - 1] : 0.0;
                s_gamma = g;
                s_pass = g < thr ? 1 : 0;
                bb.l_pass[lo + li] = (uint8_t)s_pass;
                const double dm = M, dr = rows;
                atomicAdd(bb.work + (size_t)s * MSKF_PROF_TAGS + (phase == 0 ? PK_BE_FEATURE_JAC : PK_BE_FEATURE_JAC_PRUNE),
                          dm * (dm + 1) * 0.5 * 480.0 + 6.0 * 16.0 * dm * dm + 18.0 * dr * dr + dr * dr * dr / 3.0 + 6.0 * dr * 6.0 * dm);
            }
        }
        __syncthreads();
        if (s_pass) {
            // H' = (Q^T H_xj)[3:, :] and r' to the scratch used by the stacking kernel
            for (int c = threadIdx.x; c < C6; c += BE_THREADS) {
                const int t = c / 6, cc = c - 6 * t;
                const double u0 = U[c * 3], u1 = U[c * 3 + 1], u2 = U[c * 3 + 2];
                for (int i = 3; i < R4; ++i) {
                    double x = (i >= 4 * t && i < 4 * t + 4) ? Hx[i * 6 + cc] : 0.0;
                    x -= Hf[i * 3] * u0 + Hf[i * 3 + 1] * u1 + Hf[i * 3 + 2] * u2;
                    H[(size_t)i * C6 + c] = x;
                }
            }
            for (int i = 3 + threadIdx.x; i < R4; i += BE_THREADS) rg[i] = rv[i];
        }
    }  // list loop
}

// ======================================================================================
// Prune-phase variant of the kernel above (pruneCamStateBuffer :1126-1150): every involved feature
// has exactly the two camera states being removed (M = 2: an 8 x 12 Jacobian, 5 projected rows), and
// there are hundreds of them per stream, so one THREAD handles a feature from the measurementJacobians to the
// gate: the whole problem is ~3k flops of fixed-size dense algebra that unrolls into straight-line register
// code (H' 5 x 13 lives in registers), where the warp-per-feature version spent ~1100 warp instructions per
// feature with 2 of 32 lanes in the Jacobians, 13 in the column work and two SHFLs per double in the products
// (0.17 ms per fleet launch).  What depends on the two cameras only (rotations, cam1 pose, the rotated gravity of
// the observability projection, the 12 x 12 covariance block) is evaluated once per CTA.
// ======================================================================================
#define JP_THREADS 64
struct JpCam {
    double Rw0[9], Rw1[9], p[3], t1w[3], ug[3], pn[3];
};
__global__ void __launch_bounds__(JP_THREADS) be_feature_jac_prune_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.y;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    const BeState &st = bb.st[s];
    const int n_list = st.n_list;
    if (n_list == 0) return;
    const size_t fo = (size_t)s * bc.MF, lo = (size_t)s * bc.ML;
    const double *P = bb.P + (size_t)s * bc.LD * bc.LD;
    const int LD = bc.LD;
    __shared__ double Ps[12][12];
    __shared__ JpCam cams[2];
    __shared__ double sHx[48][JP_THREADS], sHf[24][JP_THREADS], sr[8][JP_THREADS];  // per-thread columns
    const int os[2] = {st.rm_slot[0], st.rm_slot[1]};
    for (int e = threadIdx.x; e < 144; e += JP_THREADS) {
        const int i = e / 12, j = e - i * 12;
        Ps[i][j] = P[(size_t)(N21 + 6 * os[i / 6] + i % 6) * LD + N21 + 6 * os[j / 6] + j % 6];
    }
    if (threadIdx.x < 2) {
        const BeCam &c = st.cam[os[threadIdx.x]];
        JpCam &jc = cams[threadIdx.x];
        double tmp[3], Rn[9];
        quat_to_rot(c.q, jc.Rw0);
        m3mul(bc.R01, jc.Rw0, jc.Rw1);
        m3Tv(jc.Rw1, bc.t01, tmp);
        for (int i = 0; i < 3; ++i) {
            jc.p[i] = c.p[i];
            jc.t1w[i] = c.p[i] - tmp[i];
            jc.pn[i] = c.pn[i];
        }
        quat_to_rot(c.qn, Rn);
        m3v(Rn, st.g, jc.ug);
    }
    __syncthreads();
    const double gv[3] = {st.g[0], st.g[1], st.g[2]};
    const double thr = c_chi2[bc.chi2_mode][1];  // dof = involved.size() = 2, msckf_vio.cpp:1145
    for (int li = blockIdx.x * JP_THREADS + threadIdx.x; li < n_list; li += gridDim.x * JP_THREADS) {
        const int slot = bb.l_slot[lo + li];
        const double *pw = bb.f_pos + (fo + slot) * 3;
        const double pwv[3] = {pw[0], pw[1], pw[2]};
        // the two measurementJacobians run one after the other (not interleaved: their temporaries would not fit the
        // register file) and leave H_x, H_f and r in this thread's columns of shared memory
#pragma unroll 1
        for (int t = 0; t < 2; ++t) {
            // measurementJacobian, msckf_vio.cpp:610-677
            const JpCam &jc = cams[t];
            const double *z = bb.f_obs + (((size_t)s * bc.MF + slot) * bc.NS + os[t]) * 4;
            double Rw0[9], Rw1[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                Rw0[i] = jc.Rw0[i];
                Rw1[i] = jc.Rw1[i];
            }
            const double d0[3] = {pwv[0] - jc.p[0], pwv[1] - jc.p[1], pwv[2] - jc.p[2]};
            const double d1[3] = {pwv[0] - jc.t1w[0], pwv[1] - jc.t1w[1], pwv[2] - jc.t1w[2]};
            double p0[3], p1[3];
            m3v(Rw0, d0, p0);
            m3v(Rw1, d1, p1);
            const double dz0[4][3] = {{1 / p0[2], 0, -p0[0] / (p0[2] * p0[2])}, {0, 1 / p0[2], -p0[1] / (p0[2] * p0[2])}, {0, 0, 0}, {0, 0, 0}};
            const double dz1[4][3] = {{0, 0, 0}, {0, 0, 0}, {1 / p1[2], 0, -p1[0] / (p1[2] * p1[2])}, {0, 1 / p1[2], -p1[1] / (p1[2] * p1[2])}};
            double sk0[9], R01sk[9];
            skew3(p0, sk0);
            m3mul(bc.R01, sk0, R01sk);
            double Hxl[4][6];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    double a = 0, b = 0;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        a += dz0[i][k] * sk0[k * 3 + j] + dz1[i][k] * R01sk[k * 3 + j];
                        b += dz0[i][k] * (-Rw0[k * 3 + j]) + dz1[i][k] * (-Rw1[k * 3 + j]);
                    }
                    Hxl[i][j] = a;
                    Hxl[i][3 + j] = b;
                }
            // observability projection: H_x <- A - A u (u^T u)^-1 u^T, H_f <- -H_x[:, 3:6]
            double u[6], K[9];
            const double dn[3] = {pwv[0] - jc.pn[0], pwv[1] - jc.pn[1], pwv[2] - jc.pn[2]};
            u[0] = jc.ug[0]; u[1] = jc.ug[1]; u[2] = jc.ug[2];
            skew3(dn, K);
            m3v(K, gv, u + 3);
            double utu = 0;
#pragma unroll
            for (int i = 0; i < 6; ++i) utu += u[i] * u[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                double au = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) au += Hxl[i][k] * u[k];
                au *= (1.0 / utu);
#pragma unroll
                for (int k = 0; k < 6; ++k) sHx[t * 24 + i * 6 + k][threadIdx.x] = Hxl[i][k] - au * u[k];
#pragma unroll
                for (int k = 0; k < 3; ++k) sHf[(4 * t + i) * 3 + k][threadIdx.x] = -(Hxl[i][3 + k] - au * u[3 + k]);
            }
            sr[4 * t + 0][threadIdx.x] = z[0] - p0[0] / p0[2];
            sr[4 * t + 1][threadIdx.x] = z[1] - p0[1] / p0[2];
            sr[4 * t + 2][threadIdx.x] = z[2] - p1[0] / p1[2];
            sr[4 * t + 3][threadIdx.x] = z[3] - p1[1] / p1[2];
        }
        // Householder QR of H_f (8 x 3): V below the diagonal, tau
        double V[8][3], rv[8], tau[3];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            rv[i] = sr[i][threadIdx.x];
#pragma unroll
            for (int k = 0; k < 3; ++k) V[i][k] = sHf[i * 3 + k][threadIdx.x];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            double xn = 0;
#pragma unroll
            for (int i = j + 1; i < 8; ++i) xn += V[i][j] * V[i][j];
            const double alpha = V[j][j];
            double tj = 0.0;
            if (xn != 0.0) {
                const double beta = -copysign(sqrt(alpha * alpha + xn), alpha);
                tj = (beta - alpha) / beta;
                const double scale = 1.0 / (alpha - beta);
#pragma unroll
                for (int i = j + 1; i < 8; ++i) V[i][j] *= scale;
#pragma unroll
                for (int cc = j + 1; cc < 3; ++cc) {
                    double sacc = V[j][cc];
#pragma unroll
                    for (int i = j + 1; i < 8; ++i) sacc += V[i][j] * V[i][cc];
                    sacc *= tj;
                    V[j][cc] -= sacc;
#pragma unroll
                    for (int i = j + 1; i < 8; ++i) V[i][cc] -= sacc * V[i][j];
                }
            } else {
#pragma unroll
                for (int i = j + 1; i < 8; ++i) V[i][j] = 0.0;
            }
            tau[j] = tj;
        }
        // H' = (Q^T [H_xj | r])[3:, :]: column c < 12 -> camera block c / 6, column c % 6; column 12 -> r
        double Hp[5][13];
#pragma unroll
        for (int c = 0; c < 13; ++c) {
            double col[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) col[i] = c == 12 ? rv[i] : (((i >> 2) == c / 6) ? sHx[(c / 6) * 24 + (i & 3) * 6 + c % 6][threadIdx.x] : 0.0);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                double sacc = col[j];
#pragma unroll
                for (int i = j + 1; i < 8; ++i) sacc += V[i][j] * col[i];
                sacc *= tau[j];
                col[j] -= sacc;
#pragma unroll
                for (int i = j + 1; i < 8; ++i) col[i] -= sacc * V[i][j];
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) Hp[i][c] = col[3 + i];
        }
        // S = H' P_sub H'^T + sigma^2 I (lower triangle)
        double S[15];
#pragma unroll
        for (int q = 0; q < 15; ++q) S[q] = 0.0;
#pragma unroll
        for (int d = 0; d < 12; ++d) {
            double hp[5] = {0, 0, 0, 0, 0};
#pragma unroll
            for (int cc = 0; cc < 12; ++cc) {
                const double pcd = Ps[cc][d];
#pragma unroll
                for (int i = 0; i < 5; ++i) hp[i] += Hp[i][cc] * pcd;
            }
            int q = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) S[q++] += hp[i] * Hp[j][d];
        }
        // Cholesky 5 x 5 + forward substitution: gamma = r'^T S^-1 r' (gatingTest, msckf_vio.cpp:909-935)
        double g = 0;
        {
            double L[15];
            int q = 0;
#pragma unroll
            for (int i = 0; i < 5; ++i)
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    double v = S[q] + (i == j ? bc.obs_noise : 0.0);
#pragma unroll
                    for (int k = 0; k < j; ++k) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
                    L[q] = (i == j) ? sqrt(v) : v / L[j * (j + 1) / 2 + j];
                    ++q;
                }
            double y[5];
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                double v = Hp[i][12];
#pragma unroll
                for (int k = 0; k < i; ++k) v -= L[i * (i + 1) / 2 + k] * y[k];
                y[i] = v / L[i * (i + 1) / 2 + i];
                g += y[i] * y[i];
            }
        }
        const int pass = g < thr ? 1 : 0;
        if (pass) {
            double *H = bb.Hblk + (size_t)s * bc.ecap + bb.l_eoff[lo + li];   // [8][12], rows 3.. used
            double *rg = bb.rblk + (size_t)s * bc.rcap + bb.l_roff[lo + li];  // [8]
#pragma unroll
            for (int i = 0; i < 5; ++i) {
#pragma unroll
                for (int c = 0; c < 12; ++c) H[(3 + i) * 12 + c] = Hp[i][c];
                rg[3 + i] = Hp[i][12];
            }
        }
        bb.l_pass[lo + li] = (uint8_t)pass;
        bb.l_oslots[(lo + li) * NSM + 0] = (uint8_t)os[0];
        bb.l_oslots[(lo + li) * NSM + 1] = (uint8_t)os[1];
    }
    if (threadIdx.x == 0 && blockIdx.x == 0)
        bb.work[(size_t)s * MSKF_PROF_TAGS + PK_BE_FEATURE_JAC_PRUNE] += (double)n_list * (2.0 * 480.0 + 3.0 * 13.0 * 32.0 + 2.0 * 5.0 * 144.0 + 2.0 * 15.0 * 12.0 + 60.0);
}

// ======================================================================================
// Stack the gated blocks (row cap of removeLostFeatures), find the active camera columns.
// ======================================================================================
__global__ void __launch_bounds__(BE_THREADS) be_stack_kernel(BeConst bc, BeBuf bb, int phase) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    const int n = st.n_list;
    const size_t fo = (size_t)s * bc.MF, lo = (size_t)s * bc.ML;
    extern __shared__ __align__(16) unsigned char be_smem[];
    int *s_soff = (int *)be_smem;                  // [ML]
    unsigned *s_mask = (unsigned *)(s_soff + bc.ML);  // [ML]
    uint8_t *s_M = (uint8_t *)(s_mask + bc.ML);    // [ML]
    uint8_t *s_pass = s_M + bc.ML;                 // [ML]
    __shared__ int s_m, s_k, s_nuse;
    __shared__ int s_colpos[NSM];
    for (int i = threadIdx.x; i < n; i += BE_THREADS) {
        s_M[i] = (uint8_t)bb.l_M[lo + i];
        s_pass[i] = bb.l_pass[lo + i];
        s_mask[i] = phase == 0 ? bb.f_mask[fo + bb.l_slot[lo + i]] : st.rm_bits;
        s_soff[i] = -1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int cntr = 0, nuse = n;
        unsigned used = 0;
        for (int i = 0; i < n; ++i) {
            int rows = 4 * (int)s_M[i] - 3;
            if (s_pass[i] && cntr + rows <= bc.hst_rows) {
                s_soff[i] = cntr;
                cntr += rows;
                used |= s_mask[i];
            }
            if (phase == 0 && cntr > bc.max_rows) {  // msckf_vio.cpp:1009
                nuse = i + 1;
                break;
            }
        }
        // active camera columns in ascending state id: a feature lost now was seen by the most recent
        // states, so its rows are zero left of a late column and the QR can start there
        int k = 0;
        for (int c = 0; c < bc.NS; ++c) s_colpos[c] = -1;
        for (int i = 0; i < st.n_cam; ++i) {
            const int c = st.order[i];
            if (used & (1u << c)) {
                st.u_slots[k] = c;
                s_colpos[c] = k++;
            }
        }
        st.u_nslots = k;
        // the stacked matrix must fit: drop trailing blocks if a (pathological) case exceeds it
        while ((long long)cntr * 6 * k > (long long)bc.hst_cap && nuse > 0) {
            --nuse;
            if (s_soff[nuse] >= 0) {
                cntr = s_soff[nuse];
                s_soff[nuse] = -1;
                st.n_overflow++;
            }
        }
        s_m = cntr;
        s_k = 6 * k;
        s_nuse = nuse;
        st.m = cntr;
        st.k = 6 * k;
        st.do_update = cntr > 0 ? 1 : 0;
        st.dbg_m[phase] = cntr;
        st.dbg_k[phase] = 6 * k;
        st.dbg_nlist[phase] = n;
    }
    __syncthreads();
    // row offsets for the scatter kernel (-1: not stacked)
    const int nuse = s_nuse;
    for (int i = threadIdx.x; i < n; i += BE_THREADS) bb.l_soff[lo + i] = i < nuse ? s_soff[i] : -1;
    for (int c = threadIdx.x; c < bc.NS; c += BE_THREADS) st.colpos[c] = s_colpos[c];
    // the processed features leave the map (msckf_vio.cpp:1021-1023)
    if (phase == 0) {
        for (int i = threadIdx.x; i < n; i += BE_THREADS) bb.f_live[fo + bb.l_slot[lo + i]] = 0;
        if (threadIdx.x == 0) st.n_feat -= n;
    }
}

// Copies the gated per-feature blocks into the stacked system (rows from be_stack_kernel, columns
// compacted to the active camera slots).  A group of WG threads takes a feature (every feature owns its rows
// of H, so the group also zero-fills them): the whole CTA for the lost-feature update (blocks of up to
// 121 x 186), ONE WARP for the prune update, whose ~260 blocks per stream are 5 x 12 (a CTA per block spent
// its time in barriers: 132 us per fleet launch for 125 KB per stream).
template <int WG>
__global__ void __launch_bounds__(BE_THREADS) be_scatter_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.y;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    const BeState &st = bb.st[s];
    if (!st.do_update) return;
    const int n = st.n_list, k = st.k;
    const size_t lo = (size_t)s * bc.ML;
    double *Hst = bb.Hst + (size_t)s * bc.hst_cap, *rst = bb.rst + (size_t)s * bc.hst_rows;
    constexpr int GPC = BE_THREADS / WG;  // groups per CTA
    __shared__ int s_col_all[GPC][6 * NSM];  // any view count fits either variant (the prune update has M = 2)
    const int grp = threadIdx.x / WG, t = threadIdx.x % WG;
    int *s_col = s_col_all[grp];
    auto group_sync = [&]() {
        if (WG == 32) __syncwarp();
        else __syncthreads();
    };
    for (int i = blockIdx.x * GPC + grp; i < n; i += gridDim.x * GPC) {
        const int so = bb.l_soff[lo + i];
        if (so < 0) continue;  // (uniform over the group)
        const int M = bb.l_M[lo + i], C6 = 6 * M, rows = 4 * M - 3;
        const double *Hp = bb.Hblk + (size_t)s * bc.ecap + bb.l_eoff[lo + i] + 3 * C6;
        const double *rp = bb.rblk + (size_t)s * bc.rcap + bb.l_roff[lo + i] + 3;
        const uint8_t *os = bb.l_oslots + (lo + i) * NSM;
        group_sync();
        for (int c = t; c < C6; c += WG) s_col[c] = 6 * st.colpos[os[c / 6]] + (c % 6);
        for (int e = t; e < rows * k; e += WG) Hst[(size_t)so * k + e] = 0.0;
        group_sync();
        if (WG == 32) {
            for (int e = t; e < rows * C6; e += WG) {
                const int r = e / C6, c = e - r * C6;
                Hst[(size_t)(so + r) * k + s_col[c]] = Hp[e];
            }
        } else {
            for (int r = t / 32; r < rows; r += WG / 32)
                for (int c = t & 31; c < C6; c += 32) Hst[(size_t)(so + r) * k + s_col[c]] = Hp[r * C6 + c];
        }
        for (int r = t; r < rows; r += WG) rst[so + r] = rp[r];
    }
}

// ======================================================================================
// QR compression of the stacked system (measurementUpdate :795-810) through its Gram matrix.
//
// The reference replaces (H, r), m x k with m > k, by (T, Q1^T r) from a thin QR H = Q1 T.  The
// measurement noise is sigma^2 I (:833, :911), so the posterior is a function of H^T H and H^T r only:
// ANY T with T^T T = H^T H, together with r_t = T^-T H^T r on T's row space, gives the same K r and the
// same P - K S K^T.  Here T comes from the Gram matrix (CholeskyQR):
//     G = [H r]^T [H r]                        be_gram_kernel, fp64 tensor pipe (DMMA), tiled over (k+1)^2
//     G[:k,:k] = Pi R^T R Pi^T, r_t            be_pchol_kernel, Cholesky with diagonal pivoting
// instead of a chain of m-row Householder reflectors (the first version of this file: 2 m k^2 flops on
// a serial chain of k column steps per 64-row block, 1.6 + 0.8 ms per fleet launch at 3 % of the fp64
// peak).  H always has a numerical null space (moving every camera state rigidly changes no residual, so
// cond(H) ~ 1e16 on the filter's own updates): the pivoting moves those directions to the end, where the
// factorization stops once the largest remaining diagonal is below PC_TOL of the largest initial one,
// and mt = rank(H) < k rows go on.  Backward error: T^T T = H^T H + E with |E| of the order of the
// rounding of the Gram sums, the same order as the backward error of Householder QR expressed in
// H^T H; measured on the filter's own updates the posterior agrees with the oracle's Householder path
// to ~1e-14 (tests/test_gpu_backend.py::test_op_ekf_update_*, tools/upd_check.py).  Without pivoting
// (zero pivots skipped in place) the same updates show up to 2.5e-9.
// If m <= k the system is used as it is (:795 runs the QR only for tall systems).
// ======================================================================================
#define GR_K 16
#define GR_RS (GT + 4)  // row stride of a staged [row of H][column] chunk: fragment loads hit 32 distinct bank pairs
struct GramSmem {
    double a[2][GR_K * GR_RS];
    double b[2][GR_K * GR_RS];
};

// Lower tile (ti >= tj) of G = [H r]^T [H r]: C(i, j) = sum over the rows l of H of Hr(l, i0 + i) Hr(l, j0 + j).
// Both operands are rows of the same row-major matrix, staged as [l][column] with coalesced 8-byte
// cp.async (column k is the residual, from rst), two stages of 16 rows.
__global__ void __launch_bounds__(BE_THREADS, 3) be_gram_kernel(BeConst bc, BeBuf bb, int phase) {
    const int s = blockIdx.y;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    if (!st.do_update) return;
    const int m = st.m, k = st.k, KC = bc.KC;
    const double *Hst = bb.Hst + (size_t)s * bc.hst_cap, *rst = bb.rst + (size_t)s * bc.hst_rows;
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        // algorithmic flops of this stream's update (dense-equivalent over the active columns)
        const double dm = m, dk = k, dmt = m <= k ? m : k, dld = bc.LD;
        double *w = bb.work + (size_t)s * MSKF_PROF_TAGS;
        if (m > k) {
            w[phase == 0 ? PK_BE_GRAM : PK_BE_GRAM_PRUNE] += dm * (dk + 1) * (dk + 2);  // lower triangle of (k+1)^2, 2 flops per product
            w[PK_BE_PCHOL] += dk * dk * dk / 3.0;
        }
        w[PK_BE_GEMM_PHT] += 2.0 * dld * dk * dmt;
        w[PK_BE_GEMM_S] += 2.0 * dmt * dk * dmt;
        w[PK_BE_CHOL] += 2.0 * dmt * dmt * dmt / 3.0;
        w[PK_BE_GEMM_W] += dld * dmt * dmt;
        w[PK_BE_GEMM_PUPD] += dld * dld * dmt;
        st.gram_m = m;
        st.gram_k = k;
        st.gram_valid = m > k ? 1 : 0;
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < k / 6) st.gram_ids[threadIdx.x] = st.cam[st.u_slots[threadIdx.x]].id;  // mskf_debug_last_gram
    if (m <= k) {
        if (blockIdx.x != 0) return;
        double *Tm = bb.Tm + (size_t)s * KC * KC;
        int *perm = bb.perm + (size_t)s * KC;
        double *gv = bb.gv + (size_t)s * KC;
        for (int e = threadIdx.x; e < m * k; e += BE_THREADS) Tm[e] = Hst[e];
        for (int i = threadIdx.x; i < k; i += BE_THREADS) {
            perm[i] = i;
            double a = 0.0;  // (H^T r)_i, rows in order
            for (int l = 0; l < m; ++l) a = fma(Hst[(size_t)l * k + i], rst[l], a);
            gv[i] = a;
        }
        if (threadIdx.x == 0) {
            st.mt = m;
            st.t_upper = 0;
        }
        return;
    }
    const int kw = k + 1, ldg = KC + 1;
    int t = blockIdx.x, ti = 0;
    while (t > ti) { t -= ti + 1; ++ti; }
    const int tj = t;
    if (ti >= ug_tiles(kw)) return;
    int i0, mr, j0, nc;
    ug_range(kw, ti, i0, mr);
    ug_range(kw, tj, j0, nc);
    __shared__ __align__(16) GramSmem gs;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3, wr = warp >> 2, wc = warp & 3;
    const bool diag = ti == tj;
    double acc[4][2][2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int ni = 0; ni < 2; ++ni) acc[mi][ni][0] = acc[mi][ni][1] = 0.0;
    const int col = tid & 63, ca = i0 + col, cb = j0 + col;
    const bool va = col < mr, vb = !diag && col < nc;
    // column c of [H | r] at row l
    const double *pa = ca < k ? Hst + ca : rst, *pb = cb < k ? Hst + cb : rst;
    const size_t sa = ca < k ? (size_t)k : 1, sb = cb < k ? (size_t)k : 1;
    auto stage = [&](int buf, int l0) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int lo = (tid >> 6) + 4 * u, l = l0 + lo;
            const bool lv = l < m;
            cp_async8(&gs.a[buf][lo * GR_RS + col], (lv && va) ? pa + (size_t)l * sa : Hst, lv && va);
            if (!diag) cp_async8(&gs.b[buf][lo * GR_RS + col], (lv && vb) ? pb + (size_t)l * sb : Hst, lv && vb);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    bool mv[4], nv[2];
#pragma unroll
    for (int mi = 0; mi < 4; ++mi) mv[mi] = wr * 32 + mi * 8 < mr;
#pragma unroll
    for (int ni = 0; ni < 2; ++ni) nv[ni] = wc * 16 + ni * 8 < nc;
    const int nchunks = (m + GR_K - 1) / GR_K;
    stage(0, 0);
    for (int c = 0; c < nchunks; ++c) {
        if (c + 1 < nchunks) {
            stage((c + 1) & 1, (c + 1) * GR_K);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        if (mv[0] && nv[0]) {
            const double *as = gs.a[c & 1] + t4 * GR_RS + wr * 32 + g;
            const double *bs = (diag ? gs.a[c & 1] : gs.b[c & 1]) + t4 * GR_RS + wc * 16 + g;
#pragma unroll
            for (int kk = 0; kk < GR_K; kk += 4) {
                double af[4], bf[2];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi) af[mi] = as[kk * GR_RS + mi * 8];
#pragma unroll
                for (int ni = 0; ni < 2; ++ni) bf[ni] = bs[kk * GR_RS + ni * 8];
#pragma unroll
                for (int mi = 0; mi < 4; ++mi)
#pragma unroll
                    for (int ni = 0; ni < 2; ++ni)
                        if (mv[mi] && nv[ni]) dmma_884(acc[mi][ni][0], acc[mi][ni][1], af[mi], bf[ni]);
            }
        }
        __syncthreads();
    }
    double *G = bb.Gm + (size_t)s * ldg * ldg;
    upd_gemm_store(mr, nc, acc, [&](int r, int c, double v) { G[(size_t)(i0 + r) * ldg + j0 + c] = v; });
}

// ======================================================================================
// Grouped GEMMs of the update, one tile (<= 64 x 64) per CTA, grid (tiles, S).  T is the compressed
// measurement matrix over the k active camera columns: upper triangular when the QR compression ran
// (st.t_upper), which lets the k loop start at the tile's first row of T.
//   OP 0: PHt (LD x mt) = P[:, cols] T^T          OP 1: S (mt x mt, lower tiles) = T PHt[cols, :] + sigma^2 I
//   OP 2: W (LD x mt)  = PHt Linv^T (Linv lower triangular: k stops at the tile's last column)
//   OP 3: P <- P - W W^T (lower tiles; both halves written straight from the accumulators)
// ======================================================================================
template <int OP>
__global__ void __launch_bounds__(BE_THREADS, 3) be_gemm_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.y;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    const BeState &st = bb.st[s];
    if (!st.do_update) return;
    const int LD = bc.LD, KC = bc.KC, k = st.k, mt = st.mt;
    if (mt <= 0) return;
    __shared__ __align__(16) UpdGemmSmem gs;
    __shared__ int cols[6 * NSM];
    double *P = bb.P + (size_t)s * LD * LD;
    const double *Tm = bb.Tm + (size_t)s * KC * KC;
    double *PHt = bb.PHt + (size_t)s * LD * KC;
    double *Sm = bb.Sm + (size_t)s * KC * KC;
    const double *Linv = bb.Linv + (size_t)s * KC * KC;
    double *W = bb.W + (size_t)s * LD * KC;
    double acc[4][2][2];
    int i0, mr, j0, nc;
    if (OP == 0 || OP == 2) {
        const int tn = ug_tiles(mt);
        const int ti = blockIdx.x / tn, tj = blockIdx.x % tn;
        if (ti >= ug_tiles(LD)) return;
        ug_range(LD, ti, i0, mr);
        ug_range(mt, tj, j0, nc);
    } else {
        // lower-triangular tile pairs: blockIdx.x -> (ti >= tj)
        int t = blockIdx.x, ti = 0;
        while (t > ti) { t -= ti + 1; ++ti; }
        const int tj = t, n = OP == 1 ? mt : LD;
        if (ti >= ug_tiles(n)) return;
        ug_range(n, ti, i0, mr);
        ug_range(n, tj, j0, nc);
    }
    if (OP == 0 || OP == 1) {
        const int *perm = bb.perm + (size_t)s * KC;
        for (int l = threadIdx.x; l < k; l += BE_THREADS) {
            const int c = perm[l];
            cols[l] = N21 + 6 * st.u_slots[c / 6] + (c % 6);
        }
        __syncthreads();
    }
    if (OP == 0) {
        upd_gemm_tile<true>(
            mr, nc, st.t_upper ? j0 : 0, k, [&](int r, int l) { return P + (size_t)(i0 + r) * LD + cols[l]; },
            [&](int l, int c) { return Tm + (size_t)(j0 + c) * k + l; }, P, gs, acc);
        upd_gemm_store(mr, nc, acc, [&](int r, int c, double v) { PHt[(size_t)(i0 + r) * KC + j0 + c] = v; });
    } else if (OP == 1) {
        upd_gemm_tile<false>(
            mr, nc, st.t_upper ? i0 : 0, k, [&](int r, int l) { return Tm + (size_t)(i0 + r) * k + l; },
            [&](int l, int c) { return PHt + (size_t)cols[l] * KC + j0 + c; }, P, gs, acc);
        upd_gemm_store(mr, nc, acc, [&](int r, int c, double v) {
            const int i = i0 + r, j = j0 + c;
            Sm[i * KC + j] = v + (i == j ? bc.obs_noise : 0.0);
        });
    } else if (OP == 2) {
        upd_gemm_tile<true>(
            mr, nc, 0, min(mt, j0 + nc), [&](int r, int l) { return PHt + (size_t)(i0 + r) * KC + l; },
            [&](int l, int c) { return Linv + (size_t)(j0 + c) * KC + l; }, P, gs, acc);
        upd_gemm_store(mr, nc, acc, [&](int r, int c, double v) { W[(size_t)(i0 + r) * KC + j0 + c] = v; });
    } else {
        {
            // the tile of P this CTA updates is read only after the K loop: start it towards L2 now
            const int r = threadIdx.x >> 2, q = threadIdx.x & 3;  // 64 rows x 4 quarters of <= 16 doubles
            if (r < mr)
                for (int c = 4 * q * 4; c < nc && c < 4 * (q + 1) * 4; c += 4)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P + (size_t)(i0 + r) * LD + j0 + c));
        }
        upd_gemm_tile<true>(
            mr, nc, 0, mt, [&](int r, int l) { return W + (size_t)(i0 + r) * KC + l; },
            [&](int l, int c) { return W + (size_t)(j0 + c) * KC + l; }, P, gs, acc);
        // P[i][j] and P[j][i] get the same value x = P[i][j] - (W W^T)[i][j], taken from the lower triangle (on a
        // diagonal tile (i, j) and (j, i) are computed from the same products in the same order anyway), so P
        // stays exactly symmetric.  Straight from the accumulators: a thread's two adjacent columns of eight rows
        // per DMMA tile make full 64-byte row segments per warp in P, and its mirror writes (one column, eight
        // consecutive rows of the tile = eight consecutive doubles of a row of P) do too.  All loads first, then
        // all stores (the compiler must otherwise order every load behind the previous store to P).  The first
        // version staged the tile in shared memory and walked it twice with a division per element: 28 % of the
        // kernel's instructions and, with the unprefetched P reads, 43 % of its stall samples (ncu source page).
        const bool diag = i0 == j0;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        const int g = lane >> 2, t4 = lane & 3, wr = warp >> 2, wc = warp & 3;
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r = wr * 32 + mi * 8 + g, c = wc * 16 + ni * 8 + 2 * t4 + h2;
                    if (r < mr && c < nc && !(diag && c > r)) acc[mi][ni][h2] = P[(size_t)(i0 + r) * LD + j0 + c] - acc[mi][ni][h2];
                }
#pragma unroll
        for (int mi = 0; mi < 4; ++mi)
#pragma unroll
            for (int ni = 0; ni < 2; ++ni)
#pragma unroll
                for (int h2 = 0; h2 < 2; ++h2) {
                    const int r = wr * 32 + mi * 8 + g, c = wc * 16 + ni * 8 + 2 * t4 + h2;
                    if (r < mr && c < nc && !(diag && c > r)) {
                        const double x = acc[mi][ni][h2];
                        P[(size_t)(i0 + r) * LD + j0 + c] = x;
                        if (!(diag && c == r)) P[(size_t)(j0 + c) * LD + i0 + r] = x;
                    }
                }
    }
}

// S = L L^T and X = L^-1 in one pass, with the whole lower triangle in REGISTERS: thread t owns the 6x6
// tile (ti, tj) of the 31x31 lower tile triangle (496 tiles <= 512 threads, 72 registers of data), and the
// factorization is BLOCKED by tile columns.  Block step k of the right-looking scheme:
//   1. the owner of the diagonal tile factors it and inverts the 6x6 factor in registers: X_kk = L_kk^-1;  -- barrier
//   2. the tiles below it become the panel C_i = G_ik X_kk^T = L_ik; the tiles left of it (row block k of the
//      forward substitution L X = I, whose partial sums live in the columns the factorization has finished with)
//      become final, X_kj = X_kk R_kj; both are published row-major per block row (6 x n);                 -- barrier
//   3. every tile below block row k takes ONE "tile -= C_i B_j" with B = C^T (j > k: trailing Cholesky update
//      G_ij -= C_i C_j^T) or B = X_k (j <= k: R_ij -= L_ik X_kj, where tile (i, k) starts from zero).
// Two barriers and one serial 6x6 factor per SIX columns.  The column-by-column version of this layout took two
// barriers per column (1.6k cycles, against 290 cycles of DFMA issue); the version before that kept the packed
// triangle in shared memory (2.6k cycles per column for the factor and 4.1k per row for the inverse).
#define CH_THREADS 512
#define CH_NMAX (NSM * 6)

__global__ void __launch_bounds__(CH_THREADS) be_chol_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    const BeState &st = bb.st[s];
    if (!st.do_update) return;
    const int KC = bc.KC, n = st.mt;
    const double *Sm = bb.Sm + (size_t)s * KC * KC;
    double *Linv = bb.Linv + (size_t)s * KC * KC;
    constexpr int NP = CH_NMAX + CH_T;
    __shared__ __align__(16) double panel[CH_T][NP];   // panel[c][i] = L[i][6 k + c], i below block row k
    __shared__ __align__(16) double xpanel[CH_T][NP];  // xpanel[c][j] = X[6 k + c][j], j < 6 k + 6
    __shared__ __align__(16) double s_X[CH_T][CH_T];
    int ti, tj;
    chol_tile_of_thread(threadIdx.x, ti, tj);
    const int i0 = CH_T * ti, j0 = CH_T * tj;
    const int nt = (n + CH_T - 1) / CH_T;
    const bool live = ti < nt;
    double g[CH_T][CH_T];
#pragma unroll
    for (int a = 0; a < CH_T; ++a)
#pragma unroll
        for (int b = 0; b < CH_T; ++b) {
            const int i = i0 + a, j = j0 + b;
            // rows past n: identity (factors to itself, never written out)
            g[a][b] = (i < n && j <= i) ? Sm[i * KC + j] : ((i == j) ? 1.0 : 0.0);
        }
    // the strictly upper tiles of Linv are zero (be_gemm_kernel<2> reads full rows)
    for (int e = threadIdx.x; e < n * n; e += CH_THREADS) {
        const int i = e / n, j = e - i * n;
        if (j / CH_T > i / CH_T) Linv[i * KC + j] = 0.0;
    }
    for (int k = 0; k < nt; ++k) {
        if (ti == k && tj == k) {
            chol_diag_tile_inverse(g);
#pragma unroll
            for (int a = 0; a < CH_T; ++a) {
                chol_store6(&s_X[a][0], g[a]);
                chol_store6(&xpanel[a][j0], g[a]);
            }
        }
        __syncthreads();
        if (live && ti == k && tj < k) {
            // X_kj = X_kk R_kj, in place from the last row up (row a needs rows <= a of R)
#pragma unroll
            for (int a = CH_T - 1; a >= 0; --a) {
                double xr[CH_T];
                chol_load6(&s_X[a][0], xr);
#pragma unroll
                for (int b = 0; b < CH_T; ++b) {
                    double acc = xr[a] * g[a][b];
#pragma unroll
                    for (int m = 0; m < a; ++m) acc = fma(xr[m], g[m][b], acc);
                    g[a][b] = acc;
                }
                chol_store6(&xpanel[a][j0], g[a]);
            }
        }
        if (live && tj == k && ti > k) {
            // C_i = G_ik X_kk^T, column by column; the tile itself restarts from zero as R_ik
#pragma unroll
            for (int cc = 0; cc < CH_T; ++cc) {
                double xr[CH_T], col[CH_T];
                chol_load6(&s_X[cc][0], xr);
#pragma unroll
                for (int a = 0; a < CH_T; ++a) {
                    double acc = 0.0;
#pragma unroll
                    for (int b = 0; b <= cc; ++b) acc = fma(g[a][b], xr[b], acc);
                    col[a] = acc;
                }
                chol_store6(&panel[cc][i0], col);
            }
#pragma unroll
            for (int a = 0; a < CH_T; ++a)
#pragma unroll
                for (int b = 0; b < CH_T; ++b) g[a][b] = 0.0;
        }
        if (k == nt - 1) break;
        __syncthreads();
        if (live && ti > k) {
            const double(*src)[NP] = tj > k ? panel : xpanel;
#pragma unroll
            for (int cc = 0; cc < CH_T; ++cc) {
                double ci[CH_T], bj[CH_T];
                chol_load6(&panel[cc][i0], ci);
                chol_load6(&src[cc][j0], bj);
#pragma unroll
                for (int a = 0; a < CH_T; ++a)
#pragma unroll
                    for (int b = 0; b < CH_T; ++b) g[a][b] = fma(-ci[a], bj[b], g[a][b]);
            }
        }
    }
    // X = L^-1 out
    if (live) {
#pragma unroll
        for (int a = 0; a < CH_T; ++a) {
            const int i = i0 + a;
            if (i >= n) continue;
#pragma unroll
            for (int b = 0; b < CH_T; ++b) {
                const int j = j0 + b;
                if (j < n) Linv[i * KC + j] = j <= i ? g[a][b] : 0.0;
            }
        }
    }
}

#define CH_FOR_COL(ak, BODY)                                     \
    switch (ak) {                                                \
    case 0: { constexpr int B_ = 0; BODY } break;                \
    case 1: { constexpr int B_ = 1; BODY } break;                \
    case 2: { constexpr int B_ = 2; BODY } break;                \
    case 3: { constexpr int B_ = 3; BODY } break;                \
    case 4: { constexpr int B_ = 4; BODY } break;                \
    default: { constexpr int B_ = 5; BODY } break;               \
    }

// Cholesky with diagonal pivoting of the Gram matrix of the stacked system (see be_gram_kernel), one CTA per
// stream, same register layout as be_chol_kernel: thread t owns the 6x6 tile (ti, tj), tj <= ti, of G[:k,:k]
// (diagonal tiles hold the full symmetric block).  Nothing is ever swapped: step t picks the largest
// remaining diagonal entry p, broadcasts c_i = A_ip / sqrt(A_pp) (zero for earlier pivots) through shared
// memory, every tile takes A_ij -= c_i c_j, and row t of R is c in the ORIGINAL column order; earlier pivots'
// rows and columns are left as they fall (never read again: the `done` flags mask them).  Warp 0 keeps a
// copy of the diagonal (same fma, so bit-identical to the tiles') and does the arg-max for the next pivot
// while the other warps update their tiles: two barriers per step.  The compressed residual Q1^T r is never
// formed (see be_apply_kernel: delta_x comes from H^T r, row k of the Gram matrix).  It stops at the numerical rank: largest remaining diagonal <= PC_TOL x the
// largest initial one.  Output: mt = rank, perm = pivots in order followed by the unused columns, and
// T[t][l] = R[t][perm[l]] (zero for l < t): upper trapezoidal in the permuted column order, which the
// products fold into their column gather (cols[l] = camera column perm[l]).
#define PC_TOL 1e-14

__global__ void __launch_bounds__(CH_THREADS) be_pchol_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    if (!st.do_update) return;
    const int KC = bc.KC, k = st.k, ldg = KC + 1;
    if (st.m <= k) return;
    const double *G = bb.Gm + (size_t)s * ldg * ldg;
    double *Rt = bb.Rp + (size_t)s * KC * KC;  // row t of R, original column order, row stride k
    double *Tm = bb.Tm + (size_t)s * KC * KC;
    int *perm = bb.perm + (size_t)s * KC;
    __shared__ __align__(16) double vec[2][CH_NMAX + CH_T];  // column of the step, double-buffered: ONE barrier per step
    __shared__ unsigned s_open[32];
    __shared__ int s_perm[CH_NMAX + CH_T];
    const int t = threadIdx.x, lane = t & 31;
    int ti, tj;
    chol_tile_of_thread(t, ti, tj);
    const int i0 = CH_T * ti, j0 = CH_T * tj;
    const bool live = i0 < k;
    double g[CH_T][CH_T];
#pragma unroll
    for (int a = 0; a < CH_T; ++a)
#pragma unroll
        for (int b = 0; b < CH_T; ++b) {
            const int i = i0 + a, j = j0 + b;
            g[a][b] = (i < k && j < k) ? G[(size_t)max(i, j) * ldg + min(i, j)] : 0.0;
        }
    for (int i = t; i < 2 * (CH_NMAX + CH_T); i += CH_THREADS) (&vec[0][0])[i] = 0.0;
    __syncthreads();  // the first column is published before the first barrier of the loop
    // EVERY warp mirrors the diagonal (lane l: columns 6 l .. 6 l + 5, same fma as the tiles: bit-identical) and
    // repeats the pivot search, so the pivot, its scale and the stop decision are known to all threads without
    // passing through shared memory: that, and the double-buffered column, leave one barrier per step
    double d[CH_T];
    unsigned open = 0;  // bit a: column 6 lane + a has not been a pivot yet
#pragma unroll
    for (int a = 0; a < CH_T; ++a) {
        const int i = CH_T * lane + a;
        d[a] = i < k ? G[(size_t)i * ldg + i] : 0.0;
        if (i < k) open |= 1u << a;
    }
    // (value, index) arg-max over the open columns; ties go to the smaller index.  Open entries are positive
    // doubles when they matter, so their bit patterns order like integers: three warp reductions.
    auto warp_argmax = [&](double &best, int &bi) {
        double lb = -1.0;
        int li = 0x7fffffff;
#pragma unroll
        for (int a = 0; a < CH_T; ++a)
            if (((open >> a) & 1u) && d[a] > lb) {
                lb = d[a];
                li = CH_T * lane + a;
            }
        // negative or zero candidates never become pivots: treat them as "none"
        const unsigned long long key = lb > 0.0 ? (unsigned long long)__double_as_longlong(lb) : 0ull;
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
        const bool mine = hi == mhi && lo == mlo && key != 0ull;
        const unsigned mi = __reduce_min_sync(0xffffffffu, mine ? (unsigned)li : 0x7fffffffu);
        best = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
        if (mhi == 0u && mlo == 0u) best = -1.0;
        bi = (int)mi;
    };
    double best;
    int p;
    warp_argmax(best, p);
    const double d0 = best;
    bool stop = !(best > 0.0);
    double dinv = stop ? 0.0 : chol_rsqrt(best);
    int rank = 0, buf = 0;
    for (int step = 0; step < k && !stop; ++step) {
        const int tp = p / CH_T, ap = p - tp * CH_T;
        // the six "still open" flags of this thread's tile rows / tile columns (before the pivot itself closes)
        const unsigned open_i = __shfl_sync(0xffffffffu, open, ti & 31), open_j = __shfl_sync(0xffffffffu, open, tj & 31);
        double *v = vec[buf];
        if (tj == tp && ti >= tp) {
            // rows i0.. of column p (the diagonal tile holds the full block)
            CH_FOR_COL(ap,
#pragma unroll
                       for (int a = 0; a < CH_T; ++a) {
                           const int i = i0 + a;
                           if (i < k) v[i] = ((open_i >> a) & 1u) ? g[a][B_] * dinv : 0.0;
                       })
        } else if (ti == tp && tj < tp) {
            // columns j0.. of row p = rows j0.. of column p
            CH_FOR_COL(ap,
#pragma unroll
                       for (int b = 0; b < CH_T; ++b) {
                           const int j = j0 + b;
                           v[j] = ((open_j >> b) & 1u) ? g[B_][b] * dinv : 0.0;
                       })
        }
        if (lane == tp) open &= ~(1u << ap);
        if (t == 0) s_perm[step] = p;
        rank = step + 1;
        __syncthreads();
        // next pivot first (the chain every thread waits for), then the tiles
        {
            double c[CH_T];
            chol_load6(v + CH_T * lane, c);
#pragma unroll
            for (int a = 0; a < CH_T; ++a) d[a] = fma(-c[a], c[a], d[a]);
            warp_argmax(best, p);
            stop = !(step + 1 < k && best > PC_TOL * d0);
            dinv = stop ? 0.0 : chol_rsqrt(best);
        }
        if (live) {
            double ci[CH_T], rj[CH_T];
            chol_load6(v + i0, ci);
            chol_load6(v + j0, rj);
#pragma unroll
            for (int a = 0; a < CH_T; ++a)
#pragma unroll
                for (int b = 0; b < CH_T; ++b) g[a][b] = fma(-ci[a], rj[b], g[a][b]);
        }
        if (t < k) Rt[(size_t)step * k + t] = v[t];
        buf ^= 1;
    }
    if (t < 32) s_open[t] = open;
    __syncthreads();
    if (t == 0) {
        int n = rank;
        for (int i = 0; i < k; ++i)
            if ((s_open[i / CH_T] >> (i % CH_T)) & 1u) s_perm[n++] = i;
        st.mt = rank;
        st.t_upper = 1;
    }
    __syncthreads();
    for (int i = t; i < k; i += CH_THREADS) {
        perm[i] = s_perm[i];
        bb.gv[(size_t)s * KC + i] = G[(size_t)k * ldg + i];  // H^T r: row k of the Gram matrix
    }
    for (int e = t; e < rank * k; e += CH_THREADS) {
        const int i = e / k, l = e - i * k;
        Tm[e] = l >= i ? __ldcg(Rt + (size_t)i * k + s_perm[l]) : 0.0;
    }
}

// delta_x and the state correction (measurementUpdate :860-894).  One CTA per stream, AFTER the covariance
// update: with R = sigma^2 I the gain is K = P+ H^T / sigma^2 (P+ the posterior covariance), so
//     delta_x = K r = P+[:, cols] (H^T r) / sigma^2,
// one product of the k active columns of the new covariance with g = H^T r taken straight from the stacked
// system (row k of its Gram matrix).  This is the reference's K r in exact arithmetic, and it does not pass
// through the compressed residual Q1^T r = T^-T g: directions of H with singular values near 1e-8 of the
// largest carry no information but T^-T amplifies the rounding of g along them, which cost 3e-11 in delta_x
// on the filter's own updates (tools/upd_check.py), whereas this form agrees with the oracle's Householder
// path to ~1e-16 (errors of the order eps |P| |g| / sigma^2).
__global__ void __launch_bounds__(BE_THREADS) be_apply_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    if (!st.do_update) return;
    const int LD = bc.LD, KC = bc.KC, k = st.k;
    __shared__ double dx[N21 + 6 * NSM];
    __shared__ double s_g[6 * NSM];
    __shared__ int s_cols[6 * NSM];
    const double *P = bb.P + (size_t)s * LD * LD, *gv = bb.gv + (size_t)s * KC;
    for (int l = threadIdx.x; l < k; l += BE_THREADS) {
        s_cols[l] = N21 + 6 * st.u_slots[l / 6] + (l % 6);
        s_g[l] = gv[l];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < LD; i += BE_THREADS) {
        const double *row = P + (size_t)i * LD;
        double a0 = 0, a1 = 0;
        int l = 0;
        for (; l + 1 < k; l += 2) {
            a0 = fma(row[s_cols[l]], s_g[l], a0);
            a1 = fma(row[s_cols[l + 1]], s_g[l + 1], a1);
        }
        if (l < k) a0 = fma(row[s_cols[l]], s_g[l], a0);
        const double v = (a0 + a1) / bc.obs_noise;
        dx[i] = v;
        bb.dxv[(size_t)s * LD + i] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double dq[4], qn[4];
        small_angle_quat(dx, dq);
        quat_mul(dq, st.q, qn);
        for (int i = 0; i < 4; ++i) st.q[i] = qn[i];
        for (int i = 0; i < 3; ++i) {
            st.bg[i] += dx[3 + i];
            st.v[i] += dx[6 + i];
            st.ba[i] += dx[9 + i];
            st.p[i] += dx[12 + i];
        }
        double de[4], Re[9], Rn[9];
        small_angle_quat(dx + 15, de);
        quat_to_rot(de, Re);
        m3mul(Re, st.Ric, Rn);
        for (int i = 0; i < 9; ++i) st.Ric[i] = Rn[i];
        for (int i = 0; i < 3; ++i) st.tci[i] += dx[18 + i];
        st.n_updates++;
    }
    if (threadIdx.x >= 32 && threadIdx.x < 32 + bc.NS) {
        const int cs = threadIdx.x - 32;
        if (st.cam_used & (1u << cs)) {
            BeCam &c = st.cam[cs];
            double dq[4], qn[4];
            small_angle_quat(dx + N21 + 6 * cs, dq);
            quat_mul(dq, c.q, qn);
            for (int i = 0; i < 4; ++i) c.q[i] = qn[i];
            for (int i = 0; i < 3; ++i) c.p[i] += dx[N21 + 6 * cs + 3 + i];
        }
    }
}

// pruneCamStateBuffer :1161-1181: drop the two camera states.
__global__ void __launch_bounds__(BE_THREADS) be_prune_finish_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    if (!st.prune_active) return;
    const int LD = bc.LD;
    double *P = bb.P + (size_t)s * LD * LD;
    const unsigned rm = st.rm_bits;
    for (int q = 0; q < 2; ++q) {
        const int rb = N21 + 6 * st.rm_slot[q];
        for (int e = threadIdx.x; e < 6 * LD; e += BE_THREADS) {
            int i = e / LD, c = e - i * LD;
            P[(size_t)(rb + i) * LD + c] = 0.0;
            P[(size_t)c * LD + rb + i] = 0.0;
        }
    }
    const size_t fo = (size_t)s * bc.MF;
    for (int slot = threadIdx.x; slot < bc.MF; slot += BE_THREADS)
        if (bb.f_live[fo + slot]) bb.f_mask[fo + slot] &= ~rm;
    __syncthreads();
    if (threadIdx.x == 0) {
        int w = 0;
        for (int i = 0; i < st.n_cam; ++i) {
            int cs = st.order[i];
            if (rm & (1u << cs)) continue;
            st.order[w++] = cs;
        }
        st.n_cam = w;
        st.cam_used &= ~rm;
        st.prune_active = 0;
    }
}

// publish (pose) + onlineReset
__global__ void __launch_bounds__(BE_THREADS) be_finish_kernel(BeConst bc, BeBuf bb) {
    const int s = blockIdx.x;
    const BeStep sp = bb.step[s];
    if (!sp.active) return;
    BeState &st = bb.st[s];
    const int LD = bc.LD;
    double *P = bb.P + (size_t)s * LD * LD;
    __shared__ int s_reset;
    if (threadIdx.x == 0) {
        // T_b_w = T_imu_body * T_i_w * T_imu_body.inv(), msckf_vio.cpp:1242-1246
        double Rwi[9], Riw[9];
        quat_to_rot(st.q, Rwi);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Riw[i * 3 + j] = Rwi[j * 3 + i];
        double Ra[9], ta[3], tmp[3];
        m3mul(bc.Rib, Riw, Ra);
        m3v(bc.Rib, st.p, tmp);
        for (int i = 0; i < 3; ++i) ta[i] = tmp[i] + bc.tib[i];
        double Rbi[9], tbi[3], Rr[9], tr[3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) Rbi[i * 3 + j] = bc.Rib[j * 3 + i];
        m3v(Rbi, bc.tib, tmp);
        for (int i = 0; i < 3; ++i) tbi[i] = -tmp[i];
        m3mul(Ra, Rbi, Rr);
        m3v(Ra, tbi, tmp);
        for (int i = 0; i < 3; ++i) tr[i] = tmp[i] + ta[i];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) st.T_b_w[i * 4 + j] = Rr[i * 3 + j];
            st.T_b_w[i * 4 + 3] = tr[i];
            st.T_b_w[12 + i] = 0.0;
        }
        st.T_b_w[15] = 1.0;
        int reset = 0;
        if (bc.pos_std_thr > 0) {
            double sx = sqrt(P[12 * LD + 12]), sy = sqrt(P[13 * LD + 13]), sz = sqrt(P[14 * LD + 14]);
            if (!(sx < bc.pos_std_thr && sy < bc.pos_std_thr && sz < bc.pos_std_thr)) reset = 1;
        }
        s_reset = reset;
        if (reset) {
            st.n_resets++;
            st.n_cam = 0;
            st.cam_used = 0;
            st.n_feat = 0;
        }
    }
    __syncthreads();
    if (!s_reset) return;
    const size_t fo = (size_t)s * bc.MF;
    for (int slot = threadIdx.x; slot < bc.MF; slot += BE_THREADS) bb.f_live[fo + slot] = 0;
    reset_cov(bc, P);
}

// resetCallback (msckf_vio.cpp:243-304) / first-time initialisation of one stream
__global__ void __launch_bounds__(BE_THREADS) be_reset_kernel(BeConst bc, BeBuf bb, int s, int full, mskf_config cfg) {
    BeState &st = bb.st[s];
    double *P = bb.P + (size_t)s * bc.LD * bc.LD;
    const size_t fo = (size_t)s * bc.MF;
    for (int slot = threadIdx.x; slot < bc.MF; slot += BE_THREADS) bb.f_live[fo + slot] = 0;
    if (threadIdx.x == 0) {
        st.time = 0.0;
        st.q[0] = st.q[1] = st.q[2] = 0.0; st.q[3] = 1.0;
        st.qn[0] = st.qn[1] = st.qn[2] = 0.0; st.qn[3] = 1.0;
        for (int i = 0; i < 3; ++i) {
            st.p[i] = 0; st.v[i] = 0; st.bg[i] = 0; st.ba[i] = 0; st.pn[i] = 0; st.vn[i] = 0;
        }
        st.n_cam = 0;
        st.cam_used = 0;
        st.n_feat = 0;
        st.gravity_set = 0;
        st.prune_active = 0;
        st.do_update = 0;
        st.n_list = 0;
        if (full) {
            // constructor state: loadParameters (msckf_vio.cpp:58-162) + initialize (:164-188)
            st.id = 0;
            st.next_id = 0;
            st.tracking_rate = 0.0;
            st.n_updates = 0; st.n_resets = 0; st.n_overflow = 0;
            for (int i = 0; i < 3; ++i) st.v[i] = cfg.initial_velocity[i];
            st.g[0] = 0; st.g[1] = 0; st.g[2] = -9.81;
            // T_cam0_imu = inverse of the configured cam0/T_cam_imu; R_imu_cam0 = its rotation transposed
            double Rc[9], tc[3];
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) Rc[i * 3 + j] = cfg.T_cam0_imu[i * 4 + j];
                tc[i] = cfg.T_cam0_imu[i * 4 + 3];
            }
            // inv: R^T, -(R^T t); then R_imu_cam0 = (R^T)^T
            double Rt[9], tt[3];
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) Rt[i * 3 + j] = Rc[j * 3 + i];
            m3v(Rt, tc, tt);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) st.Ric[i * 3 + j] = Rt[j * 3 + i];
            for (int i = 0; i < 3; ++i) st.tci[i] = -tt[i];
            for (int i = 0; i < 16; ++i) st.T_b_w[i] = (i % 5 == 0) ? 1.0 : 0.0;
        }
    }
    __syncthreads();
    reset_cov(bc, P);
}

}  // namespace mskf

using namespace mskf;

struct BeBuffers {
    BeConst bc;
    BeBuf bb;
    BeStep *h_step[4];
    double *h_imu[4];
    cudaEvent_t ev[4];
    bool used[4];
    int pos = 0;
    double *h_poses[4];       // pinned [S][16], one slot per back-end step (ring)
    cudaEvent_t ev_poses[4];
    long long pose_seq = 0;   // back-end steps launched so far
    size_t smem_add = 0, smem_sel = 0, smem_jac[2] = {0, 0}, smem_chol = 0;
    int sort_n = 0;
};

#define BE_RING 4

int be_create(mskf_handle *h) {
    const mskf_config &c = h->cfg;
    // 31: the factorization kernels keep one 6x6 tile of the lower tile triangle per thread (31 x 32 / 2 = 496 <= 512)
    if (c.max_cam_state_size < 5 || c.max_cam_state_size > NSM - 1) {
        h->err = "max_cam_state_size must be in [5, 31]";
        return MSKF_ERR_ARG;
    }
    BeBuffers *B = new BeBuffers;
    h->bb = B;
    BeConst &bc = B->bc;
    memset(&bc, 0, sizeof(bc));
    bc.S = h->S;
    bc.NS = c.max_cam_state_size;
    bc.LD = 21 + 6 * bc.NS;
    bc.KC = 6 * bc.NS;
    bc.max_f = h->fc.max_f;
    bc.MF = 2 * (h->fc.max_f + 1) + 64;
    int hs = 1024;
    while (hs < 2 * bc.MF) hs <<= 1;
    bc.HASH = hs;
    bc.ML = bc.MF;
    bc.max_rows = c.max_jacobian_rows;
    bc.max_cam = c.max_cam_state_size;
    bc.ent_cap = h->fc.max_f > 16384 ? h->fc.max_f : 16384;
    const int lost_rows = c.max_jacobian_rows + 4 * bc.NS;      // rows a lost-feature update can stack
    const int prune_rows = 5 * bc.MF;                           // 5 rows per involved feature
    bc.hst_rows = lost_rows > prune_rows ? lost_rows : prune_rows;
    long long a = (long long)lost_rows * bc.KC, b = (long long)prune_rows * 12;
    bc.hst_cap = (int)(a > b ? a : b);
    bc.ecap = 4 * lost_rows * bc.KC;  // four times what one full update can use
    if (bc.ecap < bc.MF * 96) bc.ecap = bc.MF * 96;
    bc.rcap = bc.ecap / 6;
    bc.chi2_mode = c.chi2_mode == MSKF_CHI2_Q95 ? 1 : 0;
    bc.gyro_noise = c.noise_gyro * c.noise_gyro;
    bc.acc_noise = c.noise_acc * c.noise_acc;
    bc.gyro_bias_noise = c.noise_gyro_bias * c.noise_gyro_bias;
    bc.acc_bias_noise = c.noise_acc_bias * c.noise_acc_bias;
    bc.obs_noise = c.noise_feature * c.noise_feature;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            bc.R01[i * 3 + j] = c.T_cn_cnm1[i * 4 + j];
            bc.Rib[i * 3 + j] = c.T_imu_body[j * 4 + i];  // inverse rotation
        }
        bc.t01[i] = c.T_cn_cnm1[i * 4 + 3];
    }
    for (int i = 0; i < 3; ++i) {
        double sacc = 0;
        for (int k = 0; k < 3; ++k) sacc += bc.Rib[i * 3 + k] * c.T_imu_body[k * 4 + 3];
        bc.tib[i] = -sacc;
    }
    bc.pos_std_thr = c.position_std_threshold;
    bc.rot_thr = c.rotation_threshold;
    bc.trans_thr = c.translation_threshold;
    bc.track_thr = c.tracking_rate_threshold;
    bc.feat_trans_thr = c.feature_translation_threshold;
    bc.cov_gb = c.cov_gyro_bias; bc.cov_v = c.cov_velocity; bc.cov_ab = c.cov_acc_bias;
    bc.cov_er = c.cov_ext_rot; bc.cov_et = c.cov_ext_trans;

    BeBuf &bb = B->bb;
    memset(&bb, 0, sizeof(bb));
    const size_t S = h->S;
    int rc;
#define A(p, n) if ((rc = dev_alloc(h, &(p), (size_t)(n))) != MSKF_OK) return rc
    A(bb.st, S); A(bb.step, S); A(bb.imu, S * BE_IMU_CAP * 7);
    A(bb.P, S * bc.LD * bc.LD);
    A(bb.f_id, S * bc.MF); A(bb.f_mask, S * bc.MF); A(bb.f_live, S * bc.MF); A(bb.f_init, S * bc.MF);
    A(bb.f_pos, S * bc.MF * 3); A(bb.f_obs, S * bc.MF * bc.NS * 4); A(bb.f_last, S * bc.MF); A(bb.freelist, S * bc.MF);
    A(bb.e_cell, S * (bc.ent_cap + 1)); A(bb.inject, bc.ent_cap);
    A(bb.l_slot, S * bc.ML); A(bb.l_ok, S * bc.ML); A(bb.l_todo, S * bc.ML); A(bb.l_pass, S * bc.ML); A(bb.l_M, S * bc.ML);
    A(bb.l_eoff, S * bc.ML); A(bb.l_roff, S * bc.ML); A(bb.l_soff, S * bc.ML); A(bb.l_oslots, S * bc.ML * NSM);
    A(bb.Hblk, S * bc.ecap); A(bb.rblk, S * bc.rcap);
    A(bb.Hst, S * bc.hst_cap); A(bb.rst, S * bc.hst_rows); A(bb.Gm, S * (bc.KC + 1) * (bc.KC + 1)); A(bb.Rp, S * bc.KC * bc.KC); A(bb.perm, S * bc.KC);
    A(bb.Tm, S * bc.KC * bc.KC); A(bb.gv, S * bc.KC); A(bb.PHt, S * bc.LD * bc.KC); A(bb.Sm, S * bc.KC * bc.KC);
    A(bb.Linv, S * bc.KC * bc.KC); A(bb.W, S * bc.LD * bc.KC); A(bb.dxv, S * bc.LD);
#undef A
    bb.work = h->d_work;
    bb.fe_msg = h->fb.stale;
    bb.fe_hw = h->fb.stale_hw;
    bb.fe_total = h->fb.msg_total;
    MSKF_CUDA_CHECK(h, cudaMemcpyToSymbol(c_chi2, kChi2Q05, sizeof(double) * 99, 0));
    MSKF_CUDA_CHECK(h, cudaMemcpyToSymbol(c_chi2, kChi2Q95, sizeof(double) * 99, sizeof(double) * 99));
    for (int i = 0; i < BE_RING; ++i) {
        MSKF_CUDA_CHECK(h, cudaMallocHost((void **)&B->h_step[i], sizeof(BeStep) * S));
        MSKF_CUDA_CHECK(h, cudaMallocHost((void **)&B->h_imu[i], sizeof(double) * S * BE_IMU_CAP * 7));
        MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&B->ev[i], cudaEventDisableTiming));
        MSKF_CUDA_CHECK(h, cudaMallocHost((void **)&B->h_poses[i], sizeof(double) * 16 * S));
        MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&B->ev_poses[i], cudaEventDisableTiming));
        B->used[i] = false;
    }
    // dynamic shared memory sizes
    B->smem_add = (size_t)bc.HASH * 12;
    int sn = 256;
    while (sn < bc.MF) sn <<= 1;
    B->sort_n = sn;
    B->smem_sel = (size_t)sn * 8;
    for (int ph = 0; ph < 2; ++ph) {
        int maxM = ph == 0 ? bc.NS : 2;
        int maxR4 = 4 * maxM;
        B->smem_jac[ph] = sizeof(double) * ((size_t)maxR4 * (maxR4 + 1) / 2 + (size_t)maxM * 24 + (size_t)maxR4 * 10 + (size_t)maxM * 18);
    }
    B->smem_chol = sizeof(double) * ((size_t)bc.KC * (bc.KC + 1) / 2 + bc.KC);
    if ((rc = smem_optin(h, be_add_obs_kernel, B->smem_add)) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, be_select_kernel, B->smem_sel)) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, be_feature_jac_kernel, B->smem_jac[0])) != MSKF_OK) return rc;
    for (int s = 0; s < h->S; ++s) be_reset_kernel<<<1, BE_THREADS, 0, h->be_stream>>>(bc, bb, s, 1, h->cfg);
    MSKF_CUDA_CHECK(h, cudaGetLastError());
    return MSKF_OK;
}

void be_destroy(mskf_handle *h) {
    BeBuffers *B = h->bb;
    if (!B) return;
    for (int i = 0; i < BE_RING; ++i) {
        if (B->h_step[i]) cudaFreeHost(B->h_step[i]);
        if (B->h_imu[i]) cudaFreeHost(B->h_imu[i]);
        if (B->ev[i]) cudaEventDestroy(B->ev[i]);
        if (B->h_poses[i]) cudaFreeHost(B->h_poses[i]);
        if (B->ev_poses[i]) cudaEventDestroy(B->ev_poses[i]);
    }
    delete B;
    h->bb = nullptr;
}

int be_init_gravity(mskf_handle *h, int s) {
    BeBuffers *B = h->bb;
    HostStream &hs = h->hs[s];
    const int n = (int)hs.be_imu.size();
    if (n > BE_IMU_CAP) {
        h->err = "gravity initialisation buffer larger than BE_IMU_CAP";
        return MSKF_ERR_CAPACITY;
    }
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    std::vector<double> buf((size_t)n * 7);
    for (int i = 0; i < n; ++i) {
        buf[i * 7] = hs.be_imu[i].t;
        for (int k = 0; k < 3; ++k) {
            buf[i * 7 + 1 + k] = hs.be_imu[i].w[k];
            buf[i * 7 + 4 + k] = hs.be_imu[i].a[k];
        }
    }
    MSKF_CUDA_CHECK(h, cudaMemcpyAsync(B->bb.imu + (size_t)s * BE_IMU_CAP * 7, buf.data(), sizeof(double) * buf.size(),
                                       cudaMemcpyHostToDevice, h->be_stream));
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->be_stream));  // buf is pageable and local
    be_gravity_kernel<<<1, 32, 0, h->be_stream>>>(B->bc, B->bb, s, n);
    h->launches++;
    MSKF_CUDA_CHECK(h, cudaGetLastError());
    return MSKF_OK;
}

// measurementUpdate on the stacked system of every stream with do_update set
static void launch_update(mskf_handle *h, int phase = 0) {
    BeBuffers *B = h->bb;
    const BeConst &bc = B->bc;
    const BeBuf &bb = B->bb;
    cudaStream_t q = h->be_stream;
    const int S = h->S;
    const int tiles_ld = (bc.LD + GT - 1) / GT, tiles_kc = (bc.KC + GT - 1) / GT;
    const int tiles_kw = (bc.KC + 1 + GT - 1) / GT;
    MSKF_LAUNCH(h, phase == 0 ? PK_BE_GRAM : PK_BE_GRAM_PRUNE,
                (be_gram_kernel<<<dim3(phase == 0 ? tiles_kw * (tiles_kw + 1) / 2 : 1, S), BE_THREADS, 0, q>>>(bc, bb, phase)));
    MSKF_LAUNCH(h, PK_BE_PCHOL, (be_pchol_kernel<<<S, CH_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_GEMM_PHT, (be_gemm_kernel<0><<<dim3(tiles_ld * tiles_kc, S), BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_GEMM_S, (be_gemm_kernel<1><<<dim3(tiles_kc * (tiles_kc + 1) / 2, S), BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_CHOL, (be_chol_kernel<<<S, CH_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_GEMM_W, (be_gemm_kernel<2><<<dim3(tiles_ld * tiles_kc, S), BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_GEMM_PUPD, (be_gemm_kernel<3><<<dim3(tiles_ld * (tiles_ld + 1) / 2, S), BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_APPLY, (be_apply_kernel<<<S, BE_THREADS, 0, q>>>(bc, bb)));
}

int be_step(mskf_handle *h, const std::vector<int> &streams, const mskf_feature *inject, int n_inject, int inject_stream,
            double inject_t) {
    BeBuffers *B = h->bb;
    const BeConst &bc = B->bc;
    const BeBuf &bb = B->bb;
    // overlap on: own stream, ordered against the front end only through the message events;
    // overlap off: everything in the front end's stream (serial, used for per-kernel profiling)
    cudaStream_t q = h->be_stream;
    h->cur = q;
    const int S = h->S;
    if (!h->overlap) {
        MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_join, h->stream));
        MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(q, h->ev_join, 0));
    } else {
        MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(q, h->ev_msg_ready, 0));
    }
    // the find-or-insert hash of be_add_obs_kernel has HASH cells for <= MF live features plus the
    // message's ids: a message that could fill it is refused here (the kernel's probe is bounded as well)
    if (inject && n_inject > bc.ent_cap) {
        h->err = "too many injected measurements";
        return MSKF_ERR_CAPACITY;
    }
    if (inject && n_inject > bc.HASH - bc.MF) {  // only distinct ids take cells (a compat-mode message repeats id 0 in its tail)
        std::unordered_set<unsigned> ids;
        for (int i = 0; i < n_inject; ++i) ids.insert(inject[i].id);
        if ((int)ids.size() > bc.HASH - bc.MF) {
            h->err = "too many distinct feature ids in the injected message for the configured feature-map capacity";
            return MSKF_ERR_CAPACITY;
        }
    }
    const int slot = B->pos;
    B->pos = (B->pos + 1) % BE_RING;
    if (B->used[slot]) MSKF_CUDA_CHECK(h, cudaEventSynchronize(B->ev[slot]));
    BeStep *hstep = B->h_step[slot];
    double *himu = B->h_imu[slot];
    memset(hstep, 0, sizeof(BeStep) * S);
    int max_imu = 0;
    for (int s : streams) {
        HostStream &hs = h->hs[s];
        BeStep &sp = hstep[s];
        sp.active = 1;
        sp.t = (inject_stream == s) ? inject_t : hs.msg_t;
        sp.src = (inject_stream == s) ? 1 : 0;
        sp.n_inject = (inject_stream == s) ? n_inject : 0;
        sp.first = hs.be_first ? 1 : 0;
        if (hs.be_first) {
            hs.be_first = false;
            hs.be_time = sp.t;
        }
        // batchImuProcessing bookkeeping, msckf_vio.cpp:380-406
        size_t used = 0;
        int n = 0;
        for (const HostImu &m : hs.be_imu) {
            if (m.t < hs.be_time) {
                ++used;
                continue;
            }
            if (m.t > sp.t) break;
            if (n >= BE_IMU_CAP) break;  // the rest is taken by the next step (never with a 200 Hz IMU at 20 Hz frames)
            double *d = himu + ((size_t)s * BE_IMU_CAP + n) * 7;
            d[0] = m.t;
            for (int k = 0; k < 3; ++k) {
                d[1 + k] = m.w[k];
                d[4 + k] = m.a[k];
            }
            hs.be_time = m.t;
            ++n;
            ++used;
        }
        hs.be_imu.erase(hs.be_imu.begin(), hs.be_imu.begin() + used);
        sp.n_imu = n;
        if (n > max_imu) max_imu = n;
    }
    {
        // descriptors and the first max_imu samples of every stream: fetched by a kernel, not by the copy engine
        // (common.cuh, desc_fetch_kernel)
        static_assert(sizeof(BeStep) % 8 == 0, "descriptor fetch moves 8-byte words");
        FetchArgs fa;
        fa.n = 1;
        fa.seg[0] = fetch_seg(bb.step, hstep, sizeof(BeStep) * S);
        if (max_imu > 0) {
            fa.seg[1] = FetchSeg{bb.imu, himu, (unsigned)(7 * max_imu), (unsigned)(BE_IMU_CAP * 7), (unsigned)S};
            fa.n = 2;
        }
        desc_fetch_kernel<<<fetch_grid(fa), 256, 0, q>>>(fa);
        h->launches++;
        MSKF_CUDA_CHECK(h, cudaGetLastError());
    }
    if (inject && n_inject > 0)
        MSKF_CUDA_CHECK(h, cudaMemcpyAsync(bb.inject, inject, sizeof(mskf_feature) * n_inject, cudaMemcpyHostToDevice, q));
    MSKF_CUDA_CHECK(h, cudaEventRecord(B->ev[slot], q));
    B->used[slot] = true;
    if (inject && n_inject > 0) MSKF_CUDA_CHECK(h, cudaStreamSynchronize(q));  // caller's buffer may be pageable

    MSKF_LAUNCH(h, PK_BE_PROPAGATE, (be_propagate_kernel<<<S, 128, 0, q>>>(bc, bb, 0)));
    MSKF_LAUNCH(h, PK_BE_AUGMENT, (be_augment_kernel<<<S, BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_ADD_OBS, (be_add_obs_kernel<<<S, BE_THREADS, B->smem_add, q>>>(bc, bb)));
    MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_msg_consumed, q));
    h->msg_consumed_valid = true;
    for (int phase = 0; phase < 2; ++phase) {
        const int maxM = phase == 0 ? bc.NS : 2;
        MSKF_LAUNCH(h, PK_BE_SELECT, (be_select_kernel<<<S, BE_THREADS, B->smem_sel, q>>>(bc, bb, phase, B->sort_n)));
        {
            dim3 g(phase == 0 ? 8 : 16, S);
            MSKF_LAUNCH(h, PK_BE_TRIANGULATE, (be_triangulate_kernel<<<g, TRI_WARPS * 32, 0, q>>>(bc, bb, phase)));
        }
        MSKF_LAUNCH(h, PK_BE_LAYOUT, (be_layout_kernel<<<S, BE_THREADS, 0, q>>>(bc, bb, phase)));
        if (phase == 0) {
            // one CTA per lost feature (p99 of a stream's list is 10 in the fleet bench; longer lists loop).  The CTAs
            // without a feature only read two descriptors, but at 73 KB of shared memory each they still cost a
            // launch slot: 32 per stream made 28 waves, 30 us per launch before any work
            dim3 g(16, S);
            MSKF_LAUNCH(h, PK_BE_FEATURE_JAC, (be_feature_jac_kernel<<<g, BE_THREADS, B->smem_jac[0], q>>>(bc, bb, 0, maxM)));
        } else {
            dim3 g(8, S);
            MSKF_LAUNCH(h, PK_BE_FEATURE_JAC_PRUNE, (be_feature_jac_prune_kernel<<<g, JP_THREADS, 0, q>>>(bc, bb)));
        }
        MSKF_LAUNCH(h, PK_BE_STACK, (be_stack_kernel<<<S, BE_THREADS, (size_t)bc.ML * 10, q>>>(bc, bb, phase)));
        // the prune update's blocks are 5 x 12 (M = 2 <= 5 views: 30 columns fit a warp's table): one warp per block
        if (phase == 0) MSKF_LAUNCH(h, PK_BE_STACK, (be_scatter_kernel<BE_THREADS><<<dim3(16, S), BE_THREADS, 0, q>>>(bc, bb)));
        else MSKF_LAUNCH(h, PK_BE_STACK, (be_scatter_kernel<32><<<dim3(8, S), BE_THREADS, 0, q>>>(bc, bb)));
        launch_update(h, phase);
    }
    MSKF_LAUNCH(h, PK_BE_PRUNE_FINISH, (be_prune_finish_kernel<<<S, BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_LAUNCH(h, PK_BE_FINISH, (be_finish_kernel<<<S, BE_THREADS, 0, q>>>(bc, bb)));
    MSKF_CUDA_CHECK(h, cudaGetLastError());
    // every step's poses go to pinned host memory right away (ring of BE_RING slots)
    {
        B->pose_seq++;
        const int ps = (int)(B->pose_seq % BE_RING);
        const char *src = (const char *)bb.st + offsetof(BeState, T_b_w);
        MSKF_CUDA_CHECK(h, cudaMemcpy2DAsync(B->h_poses[ps], sizeof(double) * 16, src, sizeof(BeState), sizeof(double) * 16, S,
                                             cudaMemcpyDeviceToHost, q));
        MSKF_CUDA_CHECK(h, cudaEventRecord(B->ev_poses[ps], q));
    }
    if (!h->overlap) {
        // serial mode: the next front-end step starts after this back-end step
        MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_join, q));
        MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    }
    return MSKF_OK;
}

static int fetch_state(mskf_handle *h, int s, BeState *st) {
    MSKF_CUDA_CHECK(h, cudaMemcpy(st, h->bb->bb.st + s, sizeof(BeState), cudaMemcpyDeviceToHost));
    return MSKF_OK;
}

int be_get_state(mskf_handle *h, int s, mskf_state *out) {
    BeState st;
    int rc = fetch_state(h, s, &st);
    if (rc != MSKF_OK) return rc;
    memset(out, 0, sizeof(*out));
    out->time = st.time;
    out->id = st.id;
    for (int i = 0; i < 4; ++i) out->orientation[i] = st.q[i];
    for (int i = 0; i < 3; ++i) {
        out->position[i] = st.p[i]; out->velocity[i] = st.v[i]; out->gyro_bias[i] = st.bg[i];
        out->acc_bias[i] = st.ba[i]; out->t_cam0_imu[i] = st.tci[i]; out->gravity[i] = st.g[i];
    }
    for (int i = 0; i < 9; ++i) out->R_imu_cam0[i] = st.Ric[i];
    out->n_cam_states = st.n_cam;
    out->cov_dim = 21 + 6 * st.n_cam;
    out->is_gravity_set = st.gravity_set;
    out->n_map_features = st.n_feat;
    out->tracking_rate = st.tracking_rate;
    for (int i = 0; i < 16; ++i) out->T_b_w[i] = st.T_b_w[i];
    out->n_updates = st.n_updates;
    out->n_resets = st.n_resets;
    return MSKF_OK;
}

int be_get_cam_states(mskf_handle *h, int s, mskf_cam_state *out, int cap, int *n) {
    BeState st;
    int rc = fetch_state(h, s, &st);
    if (rc != MSKF_OK) return rc;
    *n = st.n_cam;
    for (int i = 0; i < st.n_cam && i < cap && out; ++i) {
        const BeCam &c = st.cam[st.order[i]];
        out[i].id = c.id;
        out[i].time = c.time;
        for (int k = 0; k < 4; ++k) out[i].orientation[k] = c.q[k];
        for (int k = 0; k < 3; ++k) out[i].position[k] = c.p[k];
    }
    return MSKF_OK;
}

int be_get_cov(mskf_handle *h, int s, double *out, int cap, int *dim) {
    BeState st;
    int rc = fetch_state(h, s, &st);
    if (rc != MSKF_OK) return rc;
    const int n = 21 + 6 * st.n_cam, LD = h->bb->bc.LD;
    *dim = n;
    if (!out) return MSKF_OK;
    if (cap < n * n) return MSKF_ERR_CAPACITY;
    std::vector<double> P((size_t)LD * LD);
    MSKF_CUDA_CHECK(h, cudaMemcpy(P.data(), h->bb->bb.P + (size_t)s * LD * LD, sizeof(double) * P.size(), cudaMemcpyDeviceToHost));
    std::vector<int> idx(n);
    for (int i = 0; i < 21; ++i) idx[i] = i;
    for (int c = 0; c < st.n_cam; ++c)
        for (int k = 0; k < 6; ++k) idx[21 + 6 * c + k] = 21 + 6 * st.order[c] + k;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) out[(size_t)i * n + j] = P[(size_t)idx[i] * LD + idx[j]];
    return MSKF_OK;
}

int be_reset(mskf_handle *h, int s) {
    BeBuffers *B = h->bb;
    be_reset_kernel<<<1, BE_THREADS, 0, h->be_stream>>>(B->bc, B->bb, s, 0, h->cfg);
    h->launches++;
    MSKF_CUDA_CHECK(h, cudaGetLastError());
    return MSKF_OK;
}

// T_b_w of every stream after the latest (lag 0) or the one-before-latest (lag 1) back-end step: the
// step already copied them to pinned host memory, so this only waits for that copy
int be_get_poses(mskf_handle *h, double *out, int cap_streams, int lag) {
    BeBuffers *B = h->bb;
    const int n = cap_streams < h->S ? cap_streams : h->S;
    if (n <= 0) return MSKF_OK;
    const long long seq = B->pose_seq - lag;
    if (seq <= 0) {
        for (int i = 0; i < n * 16; ++i) out[i] = (i % 5 == 0) ? 1.0 : 0.0;  // identity before the first step
        return MSKF_OK;
    }
    const int ps = (int)(seq % BE_RING);
    MSKF_CUDA_CHECK(h, cudaEventSynchronize(B->ev_poses[ps]));
    memcpy(out, B->h_poses[ps], sizeof(double) * 16 * n);
    return MSKF_OK;
}

#include <algorithm>
int be_get_map(mskf_handle *h, int s, long long *ids, int *init, double *pos, int *nobs, int cap, int *n) {
    const BeConst &bc = h->bb->bc;
    const BeBuf &bb = h->bb->bb;
    const size_t MF = bc.MF, fo = (size_t)s * MF;
    std::vector<uint8_t> live(MF), fin(MF);
    std::vector<unsigned> id(MF), mask(MF);
    std::vector<double> p(MF * 3);
    MSKF_CUDA_CHECK(h, cudaMemcpy(live.data(), bb.f_live + fo, MF, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(fin.data(), bb.f_init + fo, MF, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(id.data(), bb.f_id + fo, MF * 4, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(mask.data(), bb.f_mask + fo, MF * 4, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(p.data(), bb.f_pos + fo * 3, MF * 24, cudaMemcpyDeviceToHost));
    std::vector<std::pair<unsigned, int>> order;
    for (size_t i = 0; i < MF; ++i)
        if (live[i]) order.push_back(std::make_pair(id[i], (int)i));
    std::sort(order.begin(), order.end());
    *n = (int)order.size();
    for (int k = 0; k < *n && k < cap; ++k) {
        int i = order[k].second;
        ids[k] = order[k].first;
        init[k] = fin[i];
        for (int j = 0; j < 3; ++j) pos[k * 3 + j] = p[(size_t)i * 3 + j];
        nobs[k] = __builtin_popcount(mask[i]);
    }
    return MSKF_OK;
}

// Stand-alone measurementUpdate (msckf_vio.cpp:778-907 algebra) on caller-supplied H (m x n), r, P (n x n),
// n = 21 + 6 n_cam, run by the same kernels as the pipeline on stream 0 of a scratch handle.
int be_op_update(mskf_handle *t, int n_cam, int m, const double *H, const double *r, const double *P, double *dx, double *Pn) {
    BeBuffers *B = t->bb;
    const BeConst &bc = B->bc;
    const BeBuf &bb = B->bb;
    const int n = 21 + 6 * n_cam, LD = bc.LD, k = 6 * n_cam;
    if (n_cam < 1 || n_cam > bc.NS || m < 1 || m > bc.hst_rows || (long long)m * k > bc.hst_cap) {
        t->err = "mskf_op_ekf_update: dimensions exceed the configured capacity";
        return MSKF_ERR_CAPACITY;
    }
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->stream));
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->be_stream));
    t->cur = t->be_stream;
    BeState st;
    MSKF_CUDA_CHECK(t, cudaMemcpy(&st, bb.st, sizeof(BeState), cudaMemcpyDeviceToHost));
    st.n_cam = n_cam;
    st.cam_used = n_cam >= 32 ? 0xffffffffu : ((1u << n_cam) - 1u);
    for (int i = 0; i < n_cam; ++i) {
        st.order[i] = i;
        st.u_slots[i] = i;
        st.cam[i].q[0] = st.cam[i].q[1] = st.cam[i].q[2] = 0.0;
        st.cam[i].q[3] = 1.0;
    }
    st.u_nslots = n_cam;
    st.m = m;
    st.k = k;
    st.do_update = 1;
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.st, &st, sizeof(BeState), cudaMemcpyHostToDevice));
    BeStep sp;
    memset(&sp, 0, sizeof(sp));
    sp.active = 1;
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.step, &sp, sizeof(sp), cudaMemcpyHostToDevice));
    std::vector<double> Pl((size_t)LD * LD, 0.0), Hc((size_t)m * k);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Pl[(size_t)i * LD + j] = P[(size_t)i * n + j];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < k; ++j) Hc[(size_t)i * k + j] = H[(size_t)i * n + 21 + j];
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.P, Pl.data(), sizeof(double) * Pl.size(), cudaMemcpyHostToDevice));
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.Hst, Hc.data(), sizeof(double) * Hc.size(), cudaMemcpyHostToDevice));
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.rst, r, sizeof(double) * m, cudaMemcpyHostToDevice));
    launch_update(t);
    MSKF_CUDA_CHECK(t, cudaGetLastError());
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->be_stream));
    std::vector<double> dxl(LD);
    MSKF_CUDA_CHECK(t, cudaMemcpy(Pl.data(), bb.P, sizeof(double) * Pl.size(), cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(t, cudaMemcpy(dxl.data(), bb.dxv, sizeof(double) * LD, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        dx[i] = dxl[i];
        for (int j = 0; j < n; ++j) Pn[(size_t)i * n + j] = Pl[(size_t)i * LD + j];
    }
    return MSKF_OK;
}

// Stand-alone Feature::checkMotion + initializePosition (feature.hpp:257-450) on caller-supplied camera
// states and observations, run by be_triangulate_kernel on stream 0 of a scratch handle.
int be_op_triangulate(mskf_handle *t, int n_cam, const double *cam_q, const double *cam_p, int n_feat, const unsigned *mask,
                      const double *obs, double *pos, int *ok) {
    BeBuffers *B = t->bb;
    const BeConst &bc = B->bc;
    const BeBuf &bb = B->bb;
    if (n_cam < 1 || n_cam > bc.NS || n_feat < 1 || n_feat > bc.MF || n_feat > bc.ML) {
        t->err = "mskf_op_triangulate: dimensions exceed the configured capacity";
        return MSKF_ERR_CAPACITY;
    }
    for (int f = 0; f < n_feat; ++f)
        if (mask[f] == 0 || (n_cam < 32 && (mask[f] >> n_cam))) {
            t->err = "mskf_op_triangulate: a feature needs at least one observation, all of them by the given camera states";
            return MSKF_ERR_ARG;
        }
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->stream));
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->be_stream));
    t->cur = t->be_stream;
    BeState st;
    MSKF_CUDA_CHECK(t, cudaMemcpy(&st, bb.st, sizeof(BeState), cudaMemcpyDeviceToHost));
    st.n_cam = n_cam;
    st.cam_used = n_cam >= 32 ? 0xffffffffu : ((1u << n_cam) - 1u);
    for (int i = 0; i < n_cam; ++i) {
        st.order[i] = i;
        st.cam[i].id = i;
        for (int k = 0; k < 4; ++k) st.cam[i].q[k] = cam_q[i * 4 + k];
        for (int k = 0; k < 3; ++k) st.cam[i].p[k] = cam_p[i * 3 + k];
    }
    st.n_list = n_feat;
    st.n_todo = n_feat;
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.st, &st, sizeof(BeState), cudaMemcpyHostToDevice));
    BeStep sp;
    memset(&sp, 0, sizeof(sp));
    sp.active = 1;
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.step, &sp, sizeof(sp), cudaMemcpyHostToDevice));
    std::vector<double> o((size_t)n_feat * bc.NS * 4, 0.0);
    std::vector<int> slots(n_feat);
    for (int f = 0; f < n_feat; ++f) {
        slots[f] = f;
        for (int c = 0; c < n_cam; ++c)
            for (int k = 0; k < 4; ++k) o[((size_t)f * bc.NS + c) * 4 + k] = obs[((size_t)f * n_cam + c) * 4 + k];
    }
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.f_obs, o.data(), sizeof(double) * o.size(), cudaMemcpyHostToDevice));
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.f_mask, mask, sizeof(unsigned) * n_feat, cudaMemcpyHostToDevice));
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.l_slot, slots.data(), sizeof(int) * n_feat, cudaMemcpyHostToDevice));
    MSKF_CUDA_CHECK(t, cudaMemcpy(bb.l_todo, slots.data(), sizeof(int) * n_feat, cudaMemcpyHostToDevice));  // identity: every feature
    MSKF_CUDA_CHECK(t, cudaMemset(bb.f_init, 0, n_feat));
    MSKF_CUDA_CHECK(t, cudaMemset(bb.f_pos, 0, sizeof(double) * 3 * n_feat));
    MSKF_CUDA_CHECK(t, cudaMemset(bb.l_ok, 0, n_feat));
    MSKF_CUDA_CHECK(t, cudaDeviceSynchronize());
    MSKF_LAUNCH(t, PK_BE_TRIANGULATE, (be_triangulate_kernel<<<dim3(16, 1), TRI_WARPS * 32, 0, t->be_stream>>>(bc, bb, 0)));
    MSKF_CUDA_CHECK(t, cudaGetLastError());
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->be_stream));
    std::vector<uint8_t> okb(n_feat);
    MSKF_CUDA_CHECK(t, cudaMemcpy(okb.data(), bb.l_ok, n_feat, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(t, cudaMemcpy(pos, bb.f_pos, sizeof(double) * 3 * n_feat, cudaMemcpyDeviceToHost));
    for (int f = 0; f < n_feat; ++f) ok[f] = okb[f];
    return MSKF_OK;
}

// test hook (mskf_debug_last_gram): Gram matrix of the stacked system of the latest measurementUpdate of a stream
int be_debug_last_gram(mskf_handle *h, int s, double *G, int cap, int *m, int *k, long long *cam_ids, int *valid) {
    BeState st;
    MSKF_CUDA_CHECK(h, cudaMemcpy(&st, h->bb->bb.st + s, sizeof(BeState), cudaMemcpyDeviceToHost));
    *m = st.gram_m;
    *k = st.gram_k;
    *valid = st.gram_valid;
    for (int g = 0; g < st.gram_k / 6; ++g) cam_ids[g] = st.gram_ids[g];
    if (!st.gram_valid || !G) return MSKF_OK;
    const int kw = st.gram_k + 1, ldg = h->bb->bc.KC + 1;
    if (cap < kw * kw) return MSKF_ERR_CAPACITY;
    std::vector<double> buf((size_t)kw * ldg);
    MSKF_CUDA_CHECK(h, cudaMemcpy(buf.data(), h->bb->bb.Gm + (size_t)s * ldg * ldg, sizeof(double) * buf.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < kw; ++i)
        for (int j = 0; j <= i; ++j) G[i * kw + j] = G[j * kw + i] = buf[(size_t)i * ldg + j];  // the kernel fills the lower triangle
    return MSKF_OK;
}

// bring-up / analysis: per stream {m, k, listed features} of the lost-feature and the prune update of the last step
int be_debug_update_dims(mskf_handle *h, int *out6) {
    std::vector<BeState> st(h->S);
    MSKF_CUDA_CHECK(h, cudaMemcpy(st.data(), h->bb->bb.st, sizeof(BeState) * h->S, cudaMemcpyDeviceToHost));
    for (int s = 0; s < h->S; ++s)
        for (int ph = 0; ph < 2; ++ph) {
            out6[s * 6 + ph * 3 + 0] = st[s].dbg_m[ph];
            out6[s * 6 + ph * 3 + 1] = st[s].dbg_k[ph];
            out6[s * 6 + ph * 3 + 2] = st[s].dbg_nlist[ph];
        }
    return MSKF_OK;
}
