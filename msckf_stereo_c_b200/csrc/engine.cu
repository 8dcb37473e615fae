// engine.cu — the C ABI of include/msckf_b200.h: handle life cycle, host-side IMU
// buffering (the strictly sequential, negligible parts of the path: SURVEY a8, a14's
// bookkeeping) and step orchestration.  All image and filter arithmetic runs in the CUDA
// kernels of frontend.cu / backend.cu; there is no CPU fallback.
#include <math.h>
#include <string.h>

#include "../../include/msckf_b200_presets.h"
#include "common.cuh"

using namespace mskf;

#define STEP_RING 4

struct EngineExtra {
    cudaEvent_t ring_ev[STEP_RING];
    bool ring_used[STEP_RING];
    int ring_pos = 0;
    FeStep *h_step_ring[STEP_RING];
    const uint8_t **h_src_ring[STEP_RING];
    std::vector<mskf_feature> scratch;
};
static EngineExtra *extra(mskf_handle *h) { return (EngineExtra *)h->h_be_step; }

// ---- small host math, written to evaluate exactly like the reference's cg::Matrix code
// paths restated in the oracle (same operation order => same bits) ------------------------
static void m3_mul(const double *a, const double *b, double *c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += a[i * 3 + k] * b[k * 3 + j];
            c[i * 3 + j] = s;
        }
}
static void m3_inv(const double *a, double *r) {
    double c00 = a[4] * a[8] - a[5] * a[7];
    double c01 = a[5] * a[6] - a[3] * a[8];
    double c02 = a[3] * a[7] - a[4] * a[6];
    double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    double id = 1.0 / det;
    r[0] = c00 * id;
    r[1] = (a[2] * a[7] - a[1] * a[8]) * id;
    r[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    r[3] = c01 * id;
    r[4] = (a[0] * a[8] - a[2] * a[6]) * id;
    r[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    r[6] = c02 * id;
    r[7] = (a[1] * a[6] - a[0] * a[7]) * id;
    r[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}
// cv::Rodrigues semantics, transposed result (image_processor.cpp:882)
static void rodrigues_t(const double v[3], double Rt[9]) {
    double th = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double R[9];
    const double sk[9] = {0, -v[2], v[1], v[2], 0, -v[0], -v[1], v[0], 0};
    if (th < 1e-12) {
        for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + sk[i];
    } else {
        double k[3] = {v[0] / th, v[1] / th, v[2] / th};
        double c = cos(th), s = sin(th);
        const double skk[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j)
                R[i * 3 + j] = (i == j ? 1.0 : 0.0) * c + (k[i] * k[j]) * (1 - c) + skk[i * 3 + j] * s;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Rt[j * 3 + i] = R[i * 3 + j];
}

// ImageProcessor::integrateImuData (image_processor.cpp:850-889) + the homography of
// predictFeatureTracking (:335-340) for cam0.
static void host_predict_homography(mskf_handle *h, HostStream &hs, double H0[9], double R0[9], double R1[9]) {
    const mskf_config &c = h->cfg;
    const double prev_t = c.fix_prev_image_alias ? hs.fe_prev_t : hs.fe_curr_t;  // defect F6
    size_t begin = 0;
    while (begin < hs.fe_imu.size()) {
        if (hs.fe_imu[begin].t - prev_t < -0.01) ++begin;
        else break;
    }
    size_t end = begin;
    while (end < hs.fe_imu.size()) {
        if (hs.fe_imu[end].t - hs.fe_curr_t < 0.005) ++end;
        else break;
    }
    double mean[3] = {0, 0, 0};
    for (size_t i = begin; i < end; ++i)
        for (int k = 0; k < 3; ++k) mean[k] = mean[k] + hs.fe_imu[i].w[k];
    if (end > begin) {
        double sc = (double)(1.0f / (float)(end - begin));
        for (int k = 0; k < 3; ++k) mean[k] = mean[k] * sc;
    }
    // cam0_mean_ang_vel = R_cam0_imu^T * mean ; R_cam0_imu = rot(T_cam0_imu)^T  => rot(T) * mean
    double cam0_w[3];
    for (int i = 0; i < 3; ++i)
        cam0_w[i] = c.T_cam0_imu[i * 4 + 0] * mean[0] + c.T_cam0_imu[i * 4 + 1] * mean[1] + c.T_cam0_imu[i * 4 + 2] * mean[2];
    double dtime = hs.fe_curr_t - prev_t;
    double v[3] = {cam0_w[0] * dtime, cam0_w[1] * dtime, cam0_w[2] * dtime};
    double Rpc[9];
    rodrigues_t(v, Rpc);
    {
        // cam1: R_cam1_imu^T = rot(T_cn_cnm1 * T_cam0_imu) (image_processor.cpp:66-72, :878-883)
        double R1m[9], cam1_w[3], v1[3];
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) {
                double sacc = 0;
                for (int k = 0; k < 3; ++k) sacc += c.T_cn_cnm1[i * 4 + k] * c.T_cam0_imu[k * 4 + j];
                R1m[i * 3 + j] = sacc;
            }
        for (int i = 0; i < 3; ++i) cam1_w[i] = R1m[i * 3 + 0] * mean[0] + R1m[i * 3 + 1] * mean[1] + R1m[i * 3 + 2] * mean[2];
        for (int i = 0; i < 3; ++i) v1[i] = cam1_w[i] * dtime;
        rodrigues_t(v1, R1);
        for (int i = 0; i < 9; ++i) R0[i] = Rpc[i];
    }
    hs.fe_imu.erase(hs.fe_imu.begin(), hs.fe_imu.begin() + end);
    double K[9] = {c.cam0_intrinsics[0], 0, c.cam0_intrinsics[2], 0, c.cam0_intrinsics[1], c.cam0_intrinsics[3], 0, 0, 1.0};
    double KR[9], Ki[9];
    m3_mul(K, Rpc, KR);
    m3_inv(K, Ki);
    m3_mul(KR, Ki, H0);
}

static const char *kProfNames[PK_COUNT] = {
    "pyr_down_l1", "pyr_down_ln", "klt_temporal", "klt_stereo", "klt_new", "detect", "fe_bookkeeping",
    "be_propagate", "be_augment", "be_add_obs", "be_select", "be_triangulate", "be_layout", "be_feature_jac",
    "be_stack", "be_gram", "be_gemm_pht", "be_gemm_s", "be_chol", "be_gemm_w", "be_apply", "be_gemm_pupd",
    "be_prune_finish", "be_finish", "be_feature_jac_prune", "be_gram_prune", "be_pchol"};
const char *mskf_prof_name(int tag) { return tag >= 0 && tag < PK_COUNT ? kProfNames[tag] : "?"; }
void prof_begin(mskf_handle *h, int tag) {
    if (!h->prof_on) return;
    if (h->prof_used + 2 > h->prof_ev.size()) {
        for (int i = 0; i < 2; ++i) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            h->prof_ev.push_back(e);
        }
    }
    h->prof_tag.push_back(tag);
    cudaEventRecord(h->prof_ev[h->prof_used], h->cur ? h->cur : h->stream);
}
void prof_end(mskf_handle *h) {
    if (!h->prof_on) return;
    cudaEventRecord(h->prof_ev[h->prof_used + 1], h->cur ? h->cur : h->stream);
    h->prof_used += 2;
}
void prof_collect(mskf_handle *h) {
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, h->prof_ev[i], h->prof_ev[i + 1]) == cudaSuccess) {
            int tag = h->prof_tag[i / 2];
            h->prof_ms[tag] += ms;
            h->prof_n[tag]++;
        }
    }
    h->prof_used = 0;
    h->prof_tag.clear();
}

// ---- upload staging: copy stream + double-buffered landing area ----------------------------
static uint8_t *stage_slot_base(mskf_handle *h) {
    const size_t img = (size_t)h->cfg.img_rows * h->cfg.img_cols;
    return h->fb.staging + (size_t)h->stage_cur * h->S * 2 * img;
}
static int stage_prepare_write(mskf_handle *h) {
    if (!h->stage_waited) {
        // the kernels that read this slot two steps ago must have finished with it
        if (h->consumed_valid[h->stage_cur]) MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_consumed[h->stage_cur], 0));
        h->stage_waited = true;
    }
    h->stage_dirty = true;
    return MSKF_OK;
}
int stage_begin_consume(mskf_handle *h) {
    if (!h->stage_dirty) return MSKF_OK;
    MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_copied[h->stage_cur], h->copy_stream));
    MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, h->ev_copied[h->stage_cur], 0));
    return MSKF_OK;
}
int stage_end_consume(mskf_handle *h) {
    if (!h->stage_dirty) return MSKF_OK;
    MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_consumed[h->stage_cur], h->stream));
    h->consumed_valid[h->stage_cur] = true;
    h->stage_cur ^= 1;
    h->stage_dirty = false;
    h->stage_waited = false;
    return MSKF_OK;
}

extern "C" {

// ---- bench instrumentation: CUDA-event time per kernel class on the launching stream
int mskf_profile_enable(mskf_handle *h, int on) {
    if (!h) return MSKF_ERR_ARG;
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->be_stream);
    prof_collect(h);
    h->prof_on = on != 0;
    if (on) {
        for (int i = 0; i < MSKF_PROF_TAGS; ++i) { h->prof_ms[i] = 0; h->prof_n[i] = 0; h->work_host[i] = 0; }
        cudaMemsetAsync(h->d_work, 0, sizeof(double) * (size_t)h->S * MSKF_PROF_TAGS, h->stream);
    }
    return MSKF_OK;
}
int mskf_profile_read(mskf_handle *h, int tag, const char **name, double *ms, long long *count) {
    if (!h || tag < 0) return MSKF_ERR_ARG;
    if (tag >= PK_COUNT) return 1;
    cudaStreamSynchronize(h->stream);
    cudaStreamSynchronize(h->be_stream);
    prof_collect(h);
    if (name) *name = kProfNames[tag];
    if (ms) *ms = h->prof_ms[tag];
    if (count) *count = h->prof_n[tag];
    return MSKF_OK;
}

int mskf_default_config(mskf_config *cfg, const char *preset) { return mskf_fill_preset(cfg, preset); }

int mskf_create(const mskf_config *cfg, int n_streams, int device, mskf_handle **out) {
    if (!cfg || !out || n_streams < 1) return MSKF_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return MSKF_ERR_CUDA;
    mskf_handle *h = new mskf_handle;
    h->cfg = *cfg;
    h->S = n_streams;
    h->device = device;
    h->hs.resize(n_streams);
    *out = h;  // returned even on failure so that mskf_last_error can be read
    MSKF_CUDA_CHECK(h, cudaSetDevice(device));
    MSKF_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
    MSKF_CUDA_CHECK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    {
        // the back end is a chain of short, latency-bound kernels: high priority lets its CTAs jump the
        // queue of the front end's large grids, which then fill whatever the chain leaves idle
        int lo = 0, hi = 0;
        MSKF_CUDA_CHECK(h, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        MSKF_CUDA_CHECK(h, cudaStreamCreateWithPriority(&h->be_stream, cudaStreamNonBlocking, hi));
    }
    MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_msg_ready, cudaEventDisableTiming));
    MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_msg_consumed, cudaEventDisableTiming));
    MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    h->cur = h->stream;
    for (int i = 0; i < 2; ++i) {
        MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_copied[i], cudaEventDisableTiming));
        MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&h->ev_consumed[i], cudaEventDisableTiming));
    }
    int rc = dev_alloc(h, &h->d_work, (size_t)n_streams * MSKF_PROF_TAGS);
    if (rc != MSKF_OK) return rc;
    rc = fe_create(h);
    if (rc != MSKF_OK) return rc;
    rc = be_create(h);
    if (rc != MSKF_OK) return rc;
    EngineExtra *ex = new EngineExtra;
    h->h_be_step = ex;
    for (int i = 0; i < STEP_RING; ++i) {
        MSKF_CUDA_CHECK(h, cudaEventCreateWithFlags(&ex->ring_ev[i], cudaEventDisableTiming));
        ex->ring_used[i] = false;
        MSKF_CUDA_CHECK(h, cudaMallocHost((void **)&ex->h_step_ring[i], sizeof(FeStep) * n_streams));
        MSKF_CUDA_CHECK(h, cudaMallocHost((void **)&ex->h_src_ring[i], sizeof(uint8_t *) * 2 * n_streams));
    }
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->be_stream));
    return MSKF_OK;
}

void mskf_destroy(mskf_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->be_stream) cudaStreamSynchronize(h->be_stream);
    be_destroy(h);
    EngineExtra *ex = extra(h);
    if (ex) {
        for (int i = 0; i < STEP_RING; ++i) {
            cudaEventDestroy(ex->ring_ev[i]);
            cudaFreeHost(ex->h_step_ring[i]);
            cudaFreeHost((void *)ex->h_src_ring[i]);
        }
        delete ex;
    }
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    if (h->be_stream) cudaStreamDestroy(h->be_stream);
    if (h->ev_msg_ready) cudaEventDestroy(h->ev_msg_ready);
    if (h->ev_msg_consumed) cudaEventDestroy(h->ev_msg_consumed);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (int i = 0; i < 2; ++i) {
        if (h->ev_copied[i]) cudaEventDestroy(h->ev_copied[i]);
        if (h->ev_consumed[i]) cudaEventDestroy(h->ev_consumed[i]);
    }
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    for (void *p : h->allocs) cudaFree(p);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *mskf_last_error(const mskf_handle *h) { return h ? h->err.c_str() : "null handle"; }

int mskf_set_cuda_stream(mskf_handle *h, void *cuda_stream) {
    if (!h) return MSKF_ERR_ARG;
    cudaStreamSynchronize(h->stream);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    h->stream = (cudaStream_t)cuda_stream;
    h->cur = h->stream;
    h->own_stream = false;
    return MSKF_OK;
}

long long mskf_launch_count(const mskf_handle *h) { return h ? h->launches : 0; }

int mskf_push_imu_to(mskf_handle *h, int s, int halves, double t, const double w[3], const double a[3]) {
    if (!h || s < 0 || s >= h->S || !w || !a || !(halves & (MSKF_IMU_FRONTEND | MSKF_IMU_BACKEND))) return MSKF_ERR_ARG;
    HostStream &hs = h->hs[s];
    HostImu m;
    m.t = t;
    for (int i = 0; i < 3; ++i) { m.w[i] = w[i]; m.a[i] = a[i]; }
    if ((halves & MSKF_IMU_FRONTEND) && !hs.fe_first) hs.fe_imu.push_back(m);  // image_processor.cpp:205-211
    if (halves & MSKF_IMU_BACKEND) {                                            // msckf_vio.cpp:190-207
        hs.be_imu.push_back(m);
        if (!hs.gravity_set && hs.be_imu.size() >= 200) {
            int rc = be_init_gravity(h, s);
            if (rc != MSKF_OK) return rc;
            hs.gravity_set = true;
        }
    }
    return MSKF_OK;
}

int mskf_push_imu(mskf_handle *h, int s, double t, const double w[3], const double a[3]) {
    return mskf_push_imu_to(h, s, MSKF_IMU_FRONTEND | MSKF_IMU_BACKEND, t, w, a);
}

int mskf_push_imu_batch(mskf_handle *h, int stream, int n, const double *samples) {
    if (!h || !samples || n < 0 || stream < -1 || stream >= h->S) return MSKF_ERR_ARG;
    const int s0 = stream < 0 ? 0 : stream, s1 = stream < 0 ? h->S : stream + 1;
    for (int s = s0; s < s1; ++s) {
        const double *p = samples + (stream < 0 ? (size_t)s * n * 7 : 0);
        for (int i = 0; i < n; ++i) {
            int rc = mskf_push_imu(h, s, p[i * 7], p + i * 7 + 1, p + i * 7 + 4);
            if (rc != MSKF_OK) return rc;
        }
    }
    return MSKF_OK;
}

int mskf_push_stereo_batch(mskf_handle *h, const double *t, const uint8_t *cam0, const uint8_t *cam1, size_t stream_stride) {
    if (!h || !t || !cam0 || !cam1) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    const size_t img = (size_t)h->cfg.img_rows * h->cfg.img_cols;
    if (stream_stride < img) return MSKF_ERR_ARG;
    // one strided copy per camera on the copy stream: stream s lands at staging[slot][s][cam]
    int rc = stage_prepare_write(h);
    if (rc != MSKF_OK) return rc;
    uint8_t *base = stage_slot_base(h);
    if (cam1 == cam0 + img && stream_stride == 2 * img) {
        // the caller's frame set is laid out like the landing area ([stream][cam][pixels]): one contiguous copy
        MSKF_CUDA_CHECK(h, cudaMemcpyAsync(base, cam0, 2 * img * (size_t)h->S, cudaMemcpyHostToDevice, h->copy_stream));
    } else {
        MSKF_CUDA_CHECK(h, cudaMemcpy2DAsync(base, 2 * img, cam0, stream_stride, img, h->S, cudaMemcpyHostToDevice, h->copy_stream));
        MSKF_CUDA_CHECK(h, cudaMemcpy2DAsync(base + img, 2 * img, cam1, stream_stride, img, h->S, cudaMemcpyHostToDevice, h->copy_stream));
    }
    for (int s = 0; s < h->S; ++s) {
        HostStream &hs = h->hs[s];
        hs.pending = true;
        hs.pending_t = t[s];
        hs.src0 = base + ((size_t)s * 2 + 0) * img;
        hs.src1 = base + ((size_t)s * 2 + 1) * img;
    }
    return MSKF_OK;
}

int mskf_wait_uploads(mskf_handle *h) {
    if (!h) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->copy_stream));
    return MSKF_OK;
}

int mskf_host_alloc(void **out, size_t bytes) {
    if (!out || bytes == 0) return MSKF_ERR_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes, cudaHostAllocPortable) == cudaSuccess ? MSKF_OK : MSKF_ERR_CUDA;
}

int mskf_host_alloc_wc(void **out, size_t bytes) {
    if (!out || bytes == 0) return MSKF_ERR_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes, cudaHostAllocPortable | cudaHostAllocWriteCombined) == cudaSuccess ? MSKF_OK : MSKF_ERR_CUDA;
}

void mskf_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int mskf_push_stereo_device_batch(mskf_handle *h, const double *t, const uint8_t *d_cam0, const uint8_t *d_cam1, size_t stream_stride) {
    if (!h || !t || !d_cam0 || !d_cam1) return MSKF_ERR_ARG;
    if ((((uintptr_t)d_cam0) | ((uintptr_t)d_cam1) | (uintptr_t)stream_stride) & 15) {
        h->err = "device images (and the stream stride) must be 16-byte aligned";
        return MSKF_ERR_ARG;
    }
    for (int s = 0; s < h->S; ++s) {
        HostStream &hs = h->hs[s];
        hs.pending = true;
        hs.pending_t = t[s];
        hs.src0 = d_cam0 + (size_t)s * stream_stride;
        hs.src1 = d_cam1 + (size_t)s * stream_stride;
    }
    return MSKF_OK;
}

int mskf_get_work(mskf_handle *h, int tag, double *total) {
    if (!h || !total || tag < 0 || tag >= MSKF_PROF_TAGS) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    std::vector<double> w((size_t)h->S * MSKF_PROF_TAGS);
    MSKF_CUDA_CHECK(h, cudaMemcpy(w.data(), h->d_work, sizeof(double) * w.size(), cudaMemcpyDeviceToHost));
    double acc = h->work_host[tag];
    for (int s = 0; s < h->S; ++s) acc += w[(size_t)s * MSKF_PROF_TAGS + tag];
    *total = acc;
    return MSKF_OK;
}

int mskf_push_stereo(mskf_handle *h, int s, double t, const uint8_t *cam0, const uint8_t *cam1, int rows, int cols,
                     int stride) {
    if (!h || s < 0 || s >= h->S || !cam0 || !cam1) return MSKF_ERR_ARG;
    if (rows != h->cfg.img_rows || cols != h->cfg.img_cols || stride < cols) {
        h->err = "image geometry does not match the configuration";
        return MSKF_ERR_ARG;
    }
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    const size_t img = (size_t)rows * cols;
    int rc = stage_prepare_write(h);
    if (rc != MSKF_OK) return rc;
    uint8_t *d0 = stage_slot_base(h) + ((size_t)s * 2 + 0) * img, *d1 = stage_slot_base(h) + ((size_t)s * 2 + 1) * img;
    // pageable sources are staged by the runtime before the call returns; page-locked
    // sources are read asynchronously and must stay unchanged until the next mskf_sync()
    MSKF_CUDA_CHECK(h, cudaMemcpy2DAsync(d0, cols, cam0, stride, cols, rows, cudaMemcpyHostToDevice, h->copy_stream));
    MSKF_CUDA_CHECK(h, cudaMemcpy2DAsync(d1, cols, cam1, stride, cols, rows, cudaMemcpyHostToDevice, h->copy_stream));
    HostStream &hs = h->hs[s];
    hs.pending = true;
    hs.pending_t = t;
    hs.src0 = d0;
    hs.src1 = d1;
    return MSKF_OK;
}

int mskf_push_stereo_device(mskf_handle *h, int s, double t, const uint8_t *d_cam0, const uint8_t *d_cam1) {
    if (!h || s < 0 || s >= h->S || !d_cam0 || !d_cam1) return MSKF_ERR_ARG;
    if ((((uintptr_t)d_cam0) | ((uintptr_t)d_cam1)) & 15) {
        h->err = "device images must be 16-byte aligned";
        return MSKF_ERR_ARG;
    }
    HostStream &hs = h->hs[s];
    hs.pending = true;
    hs.pending_t = t;
    hs.src0 = d_cam0;
    hs.src1 = d_cam1;
    return MSKF_OK;
}

int mskf_frontend_step(mskf_handle *h) {
    if (!h) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    EngineExtra *ex = extra(h);
    const int S = h->S;
    int slot = ex->ring_pos;
    ex->ring_pos = (ex->ring_pos + 1) % STEP_RING;
    if (ex->ring_used[slot]) MSKF_CUDA_CHECK(h, cudaEventSynchronize(ex->ring_ev[slot]));
    FeStep *hstep = ex->h_step_ring[slot];
    const uint8_t **hsrc = ex->h_src_ring[slot];
    bool any = false, any_first = false;
    int max_prev = 0, n_active = 0;
    for (int s = 0; s < S; ++s) {
        HostStream &hs = h->hs[s];
        FeStep &st = hstep[s];
        memset(&st, 0, sizeof(st));
        hsrc[s] = hs.src0;
        hsrc[S + s] = hs.src1;
        if (!hs.pending) continue;
        any = true;
        ++n_active;
        st.active = 1;
        st.is_first = hs.fe_first ? 1 : 0;
        hs.fe_curr_t = hs.pending_t;
        st.t = hs.pending_t;
        if (hs.fe_first) {
            any_first = true;
            hs.slot = 0;
        } else {
            hs.slot ^= 1;  // std::swap(prev_cam0_pyramid_, curr_cam0_pyramid_), image_processor.cpp:194
            host_predict_homography(h, hs, st.H0, st.R0, st.R1);
            max_prev = h->fc.max_f;
        }
        st.slot = hs.slot;
        hs.fe_prev_t = hs.fe_curr_t;
        hs.fe_first = false;
        hs.pending = false;
        hs.published = true;
        hs.msg_t = st.t;
    }
    if (!any) return MSKF_OK;
    {
        int rc = stage_begin_consume(h);
        if (rc != MSKF_OK) return rc;
    }
    {
        static_assert(sizeof(FeStep) % 8 == 0 && sizeof(uint8_t *) == 8, "descriptor fetch moves 8-byte words");
        FetchArgs fa;
        fa.n = 3;
        fa.seg[0] = fetch_seg(h->fb.step, hstep, sizeof(FeStep) * S);
        fa.seg[1] = fetch_seg(h->fb.src0, hsrc, sizeof(uint8_t *) * S);
        fa.seg[2] = fetch_seg(h->fb.src1, hsrc + S, sizeof(uint8_t *) * S);
        desc_fetch_kernel<<<fetch_grid(fa), 256, 0, h->stream>>>(fa);
        h->launches++;
        MSKF_CUDA_CHECK(h, cudaGetLastError());
    }
    MSKF_CUDA_CHECK(h, cudaEventRecord(ex->ring_ev[slot], h->stream));
    ex->ring_used[slot] = true;
    return fe_step(h, any_first, max_prev, n_active);
}

int mskf_backend_step(mskf_handle *h) {
    if (!h) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    std::vector<int> streams;
    for (int s = 0; s < h->S; ++s) {
        HostStream &hs = h->hs[s];
        if (!hs.published) continue;
        hs.published = false;
        if (!hs.gravity_set) continue;  // msckf_vio.cpp:308
        streams.push_back(s);
    }
    if (streams.empty()) return MSKF_OK;
    return be_step(h, streams, nullptr, 0, -1, 0.0);
}

int mskf_step(mskf_handle *h) {
    int rc = mskf_frontend_step(h);
    if (rc != MSKF_OK) return rc;
    return mskf_backend_step(h);
}

int mskf_sync(mskf_handle *h) {
    if (!h) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->copy_stream));
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->stream));
    MSKF_CUDA_CHECK(h, cudaStreamSynchronize(h->be_stream));
    if (h->prof_on) prof_collect(h);
    return MSKF_OK;
}

int mskf_backend_step_features(mskf_handle *h, int s, double t, const mskf_feature *f, int n) {
    if (!h || s < 0 || s >= h->S || n < 0 || (n > 0 && !f)) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    HostStream &hs = h->hs[s];
    hs.published = false;
    if (!hs.gravity_set) return MSKF_OK;
    std::vector<int> streams(1, s);
    return be_step(h, streams, f, n, s, t);
}

int mskf_get_features(mskf_handle *h, int s, mskf_feature *out, int cap, int *n, double *t) {
    if (!h || s < 0 || s >= h->S || !n) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    const FeConst &fc = h->fc;
    int cur = 0, hw = 0;
    long long total = 0;
    MSKF_CUDA_CHECK(h, cudaMemcpy(&cur, h->fb.msg_n + s, sizeof(int), cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(&hw, h->fb.stale_hw + s, sizeof(int), cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(&total, h->fb.msg_total + s, sizeof(long long), cudaMemcpyDeviceToHost));
    *n = total > 0x7fffffffLL ? 0x7fffffff : (int)total;  // the reference's vector grows without bound; see mskf_get_features_head
    if (t) *t = h->hs[s].msg_t;
    if (!out || cap <= 0) return MSKF_OK;
    std::vector<mskf_feature> buf(fc.max_f);
    // entries [0, cur) are this frame's; [cur, hw) are stale leftovers; the rest are value-initialised
    MSKF_CUDA_CHECK(h, cudaMemcpy(buf.data(), h->fb.stale + (size_t)s * fc.max_f, sizeof(mskf_feature) * (size_t)hw, cudaMemcpyDeviceToHost));
    (void)cur;
    for (long long i = 0; i < total && i < cap; ++i) {
        if (i < hw) out[i] = buf[(size_t)i];
        else memset(&out[i], 0, sizeof(mskf_feature));
    }
    return MSKF_OK;
}

// The never-cleared message without its O(run length) value-initialised tail: entries [0, *n_head) are all that
// is not {id 0, zeros}; *n_total is the length of the reference's vector (64-bit: it grows every frame).
int mskf_get_features_head(mskf_handle *h, int s, mskf_feature *out, int cap, int *n_head, long long *n_total, double *t) {
    if (!h || s < 0 || s >= h->S || !n_head) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    int hw = 0;
    long long total = 0;
    MSKF_CUDA_CHECK(h, cudaMemcpy(&hw, h->fb.stale_hw + s, sizeof(int), cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(&total, h->fb.msg_total + s, sizeof(long long), cudaMemcpyDeviceToHost));
    *n_head = hw;
    if (n_total) *n_total = total;
    if (t) *t = h->hs[s].msg_t;
    if (!out || cap <= 0 || hw == 0) return MSKF_OK;
    const int k = hw < cap ? hw : cap;
    MSKF_CUDA_CHECK(h, cudaMemcpy(out, h->fb.stale + (size_t)s * h->fc.max_f, sizeof(mskf_feature) * (size_t)k, cudaMemcpyDeviceToHost));
    return MSKF_OK;
}

int mskf_get_n_published(mskf_handle *h, int s, int *n) {
    if (!h || s < 0 || s >= h->S || !n) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    MSKF_CUDA_CHECK(h, cudaMemcpy(n, h->fb.msg_n + s, sizeof(int), cudaMemcpyDeviceToHost));
    return MSKF_OK;
}

int mskf_get_tracking_info(mskf_handle *h, int s, mskf_tracking_info *out) {
    if (!h || s < 0 || s >= h->S || !out) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    MSKF_CUDA_CHECK(h, cudaMemcpy(out, h->fb.info + s, sizeof(*out), cudaMemcpyDeviceToHost));
    return MSKF_OK;
}

int mskf_get_grid(mskf_handle *h, int s, mskf_grid_feature *out, int cap, int *n) {
    if (!h || s < 0 || s >= h->S || !n) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    const FeConst &fc = h->fc;
    int gp = 0, cnt = 0;
    MSKF_CUDA_CHECK(h, cudaMemcpy(&gp, h->fb.gslot + s, sizeof(int), cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(&cnt, h->fb.g_n[gp] + s, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    if (!out || cap <= 0 || cnt == 0) return MSKF_OK;
    std::vector<unsigned long long> id(cnt);
    std::vector<float> resp(cnt);
    std::vector<int> life(cnt), cell(cnt);
    std::vector<float2> c0(cnt), c1(cnt);
    const size_t go = (size_t)s * fc.max_f;
    MSKF_CUDA_CHECK(h, cudaMemcpy(id.data(), h->fb.g_id[gp] + go, sizeof(unsigned long long) * cnt, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(resp.data(), h->fb.g_resp[gp] + go, sizeof(float) * cnt, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(life.data(), h->fb.g_life[gp] + go, sizeof(int) * cnt, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(cell.data(), h->fb.g_cell[gp] + go, sizeof(int) * cnt, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(c0.data(), h->fb.g_cam0[gp] + go, sizeof(float2) * cnt, cudaMemcpyDeviceToHost));
    MSKF_CUDA_CHECK(h, cudaMemcpy(c1.data(), h->fb.g_cam1[gp] + go, sizeof(float2) * cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt && i < cap; ++i) {
        out[i].id = id[i]; out[i].response = resp[i]; out[i].lifetime = life[i];
        out[i].cam0_x = c0[i].x; out[i].cam0_y = c0[i].y; out[i].cam1_x = c1[i].x; out[i].cam1_y = c1[i].y;
        out[i].cell = cell[i]; out[i].pad = 0;
    }
    return MSKF_OK;
}

int mskf_get_pyramid(mskf_handle *h, int s, int cam, int level, uint8_t *out, int cap, int *rows, int *cols) {
    if (!h || s < 0 || s >= h->S || cam < 0 || cam > 1 || level < 0 || level >= h->fc.levels || !rows || !cols)
        return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    const FeConst &fc = h->fc;
    *rows = fc.lvl_rows[level];
    *cols = fc.lvl_cols[level];
    size_t bytes = (size_t)(*rows) * (*cols);
    if (!out) return MSKF_OK;
    if ((size_t)cap < bytes) return MSKF_ERR_CAPACITY;
    const uint8_t *p = (cam == 0 ? h->fb.pyr[h->hs[s].slot] : h->fb.pyr[2]) + (size_t)s * fc.pyr_bytes + fc.lvl_off[level];
    MSKF_CUDA_CHECK(h, cudaMemcpy2D(out, (size_t)*cols, p, (size_t)fc.pitch, (size_t)*cols, (size_t)*rows, cudaMemcpyDeviceToHost));
    return MSKF_OK;
}

int mskf_get_state(mskf_handle *h, int s, mskf_state *out) {
    if (!h || s < 0 || s >= h->S || !out) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    return be_get_state(h, s, out);
}
int mskf_get_cam_states(mskf_handle *h, int s, mskf_cam_state *out, int cap, int *n) {
    if (!h || s < 0 || s >= h->S || !n) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    return be_get_cam_states(h, s, out, cap, n);
}
int mskf_get_covariance(mskf_handle *h, int s, double *out, int cap, int *dim) {
    if (!h || s < 0 || s >= h->S || !dim) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    return be_get_cov(h, s, out, cap, dim);
}
int mskf_debug_get_map(mskf_handle *h, int s, long long *ids, int *init, double *pos, int *nobs, int cap, int *n) {
    if (!h || s < 0 || s >= h->S || !n) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    return be_get_map(h, s, ids, init, pos, nobs, cap, n);
}
int mskf_debug_update_dims(mskf_handle *h, int *out6) {
    if (!h || !out6) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    return be_debug_update_dims(h, out6);
}
int mskf_debug_last_gram(mskf_handle *h, int stream, double *G, int cap, int *m, int *k, long long *cam_ids, int *valid) {
    if (!h || stream < 0 || stream >= h->S || !m || !k || !cam_ids || !valid) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    return be_debug_last_gram(h, stream, G, cap, m, k, cam_ids, valid);
}
int mskf_get_poses(mskf_handle *h, double *out, int cap_streams) {
    if (!h || !out) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    return be_get_poses(h, out, cap_streams, 0);
}
int mskf_get_poses_prev(mskf_handle *h, double *out, int cap_streams) {
    if (!h || !out) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaSetDevice(h->device));
    return be_get_poses(h, out, cap_streams, 1);
}
int mskf_join(mskf_handle *h) {
    if (!h) return MSKF_ERR_ARG;
    MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_join, h->be_stream));
    MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(h->stream, h->ev_join, 0));
    return MSKF_OK;
}
int mskf_set_overlap(mskf_handle *h, int on) {
    if (!h) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    h->overlap = on != 0;
    return MSKF_OK;
}
int mskf_reset(mskf_handle *h, int s) {
    if (!h || s < 0 || s >= h->S) return MSKF_ERR_ARG;
    int rc = mskf_sync(h);
    if (rc != MSKF_OK) return rc;
    HostStream &hs = h->hs[s];
    hs.be_imu.clear();
    hs.gravity_set = false;
    hs.be_first = true;
    hs.be_time = 0;
    return be_reset(h, s);
}

// ---- stand-alone operators: a private one-stream (or n-stream) engine runs the same kernels
static int op_make(mskf_handle *h, int rows, int cols, int levels, int n_streams, mskf_handle **tmp) {
    mskf_config c = h->cfg;
    c.img_rows = rows;
    c.img_cols = cols;
    if (levels > 0) c.pyramid_levels = levels;
    int rc = mskf_create(&c, n_streams, h->device, tmp);
    if (rc != MSKF_OK && *tmp) {
        h->err = (*tmp)->err;
        mskf_destroy(*tmp);
        *tmp = nullptr;
    }
    return rc;
}

int mskf_op_pyramid(mskf_handle *h, const uint8_t *img, int n_images, int rows, int cols, int levels, uint8_t *out) {
    if (!h || !img || !out || n_images < 1 || levels < 2) return MSKF_ERR_ARG;
    mskf_handle *t = nullptr;
    int S = (n_images + 1) / 2;
    int rc = op_make(h, rows, cols, levels, S, &t);
    if (rc != MSKF_OK) return rc;
    const size_t isz = (size_t)rows * cols;
    for (int s = 0; s < S && rc == MSKF_OK; ++s) {
        const uint8_t *a = img + (size_t)(2 * s) * isz;
        const uint8_t *b = (2 * s + 1 < n_images) ? img + (size_t)(2 * s + 1) * isz : a;
        rc = mskf_push_stereo(t, s, 0.0, a, b, rows, cols, cols);
    }
    if (rc == MSKF_OK) rc = mskf_frontend_step(t);
    size_t per = 0;
    for (int l = 1; l < levels; ++l) per += (size_t)t->fc.lvl_rows[l] * t->fc.lvl_cols[l];
    for (int i = 0; i < n_images && rc == MSKF_OK; ++i) {
        uint8_t *o = out + (size_t)i * per;
        for (int l = 1; l < levels && rc == MSKF_OK; ++l) {
            int r, q;
            rc = mskf_get_pyramid(t, i / 2, i & 1, l, o, (int)((size_t)t->fc.lvl_rows[l] * t->fc.lvl_cols[l]), &r, &q);
            o += (size_t)r * q;
        }
    }
    if (rc != MSKF_OK) h->err = t->err;
    mskf_destroy(t);
    return rc;
}


static int op_detect_impl(mskf_handle *h, const uint8_t *img, int rows, int cols, const float *occupied_xy,
                          int n_occupied, float *out_xy, double *out_response, int cap, int *n, uint8_t *score_map) {
    if (!h || !img || !n) return MSKF_ERR_ARG;
    mskf_handle *t = nullptr;
    int rc = op_make(h, rows, cols, 0, 1, &t);
    if (rc != MSKF_OK) return rc;
    rc = mskf_push_stereo(t, 0, 0.0, img, img, rows, cols, cols);
    if (rc == MSKF_OK) rc = fe_op_detect(t, occupied_xy, n_occupied, out_xy, out_response, cap, n, score_map);
    if (rc != MSKF_OK) h->err = t->err;
    mskf_destroy(t);
    return rc;
}
int mskf_op_detect(mskf_handle *h, const uint8_t *img, int rows, int cols, const float *occupied_xy, int n_occupied,
                   float *out_xy, double *out_response, int cap, int *n) {
    return op_detect_impl(h, img, rows, cols, occupied_xy, n_occupied, out_xy, out_response, cap, n, nullptr);
}
// test hook: also returns the per-pixel FAST score map (0 = not a corner)
int mskf_debug_detect_scores(mskf_handle *h, const uint8_t *img, int rows, int cols, float *out_xy,
                             double *out_response, int cap, int *n, uint8_t *score_map) {
    return op_detect_impl(h, img, rows, cols, nullptr, 0, out_xy, out_response, cap, n, score_map);
}

int mskf_op_klt(mskf_handle *h, const uint8_t *img_a, const uint8_t *img_b, int rows, int cols, const float *pts_a,
                float *pts_b, uint8_t *status, int n) {
    if (!h || !img_a || !img_b || n < 0) return MSKF_ERR_ARG;
    if (n == 0) return MSKF_OK;
    mskf_handle *t = nullptr;
    int rc = op_make(h, rows, cols, 0, 1, &t);
    if (rc != MSKF_OK) return rc;
    if (n > t->fc.cap_k) {
        h->err = "mskf_op_klt: too many points for one call";
        mskf_destroy(t);
        return MSKF_ERR_CAPACITY;
    }
    rc = mskf_push_stereo(t, 0, 0.0, img_a, img_b, rows, cols, cols);
    if (rc == MSKF_OK) rc = fe_op_klt(t, pts_a, pts_b, status, n);
    if (rc != MSKF_OK) h->err = t->err;
    mskf_destroy(t);
    return rc;
}

// measurementUpdate as a stand-alone operator: H is m x n with n = 21 + 6 n_cam (IMU columns must be zero, as
// featureJacobian leaves them: msckf_vio.cpp:709-712), observation noise from the configuration.
int mskf_op_ekf_update(mskf_handle *h, int n_cam, int m, const double *H, const double *r, const double *P,
                       double *out_delta_x, double *out_P) {
    if (!h || !H || !r || !P || !out_delta_x || !out_P) return MSKF_ERR_ARG;
    mskf_handle *t = nullptr;
    mskf_config c = h->cfg;
    if (n_cam > c.max_cam_state_size) c.max_cam_state_size = n_cam;
    if (m > c.max_jacobian_rows) c.max_jacobian_rows = m;
    int rc = mskf_create(&c, 1, h->device, &t);
    if (rc == MSKF_OK) {
        t->prof_on = h->prof_on;  // per-kernel times of the operator are accounted to the caller's handle
        rc = be_op_update(t, n_cam, m, H, r, P, out_delta_x, out_P);
        if (t->prof_on) {
            prof_collect(t);
            for (int i = 0; i < PK_COUNT; ++i) { h->prof_ms[i] += t->prof_ms[i]; h->prof_n[i] += t->prof_n[i]; }
        }
    }
    if (rc != MSKF_OK && t) h->err = t->err;
    if (t) mskf_destroy(t);
    return rc;
}

// Feature::checkMotion + initializePosition as a stand-alone operator (feature.hpp:257-450)
int mskf_op_triangulate(mskf_handle *h, int n_cam, const double *cam_orientation, const double *cam_position, int n_feat,
                        const unsigned *obs_mask, const double *obs, double *out_position, int *out_ok) {
    if (!h || !cam_orientation || !cam_position || !obs_mask || !obs || !out_position || !out_ok) return MSKF_ERR_ARG;
    mskf_handle *t = nullptr;
    mskf_config c = h->cfg;
    if (n_cam > c.max_cam_state_size) c.max_cam_state_size = n_cam;
    int rc = mskf_create(&c, 1, h->device, &t);
    if (rc == MSKF_OK) rc = be_op_triangulate(t, n_cam, cam_orientation, cam_position, n_feat, obs_mask, obs, out_position, out_ok);
    if (rc != MSKF_OK && t) h->err = t->err;
    if (t) mskf_destroy(t);
    return rc;
}

}  // extern "C"
