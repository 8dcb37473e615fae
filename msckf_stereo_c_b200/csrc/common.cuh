// common.cuh — shared declarations of the sm_100a engine behind include/msckf_b200.h.
//
// Memory layout in HBM (per handle, S = n_streams):
//   pyramids   3 sets [S][pyr_bytes]  (cam0 ping, cam0 pong, cam1); level l of a stream
//              starts at lvl_off[l], rows tightly packed (cols bytes per row)
//   staging    [S][2][rows*cols]      level-0 landing area for host uploads
//   front end  per-stream grids, KLT work lists, detector tables (struct FeBuffers)
//   back end   per-stream filter state, covariance [S][LD*LD] fp64, feature map,
//              Jacobian scratch (struct BeBuffers)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/msckf_b200.h"

#define MSKF_PROF_TAGS 32

#define MSKF_CUDA_CHECK(h, expr)                                                            \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            (h)->err = std::string(#expr) + ": " + cudaGetErrorString(_e);                  \
            return MSKF_ERR_CUDA;                                                           \
        }                                                                                   \
    } while (0)

namespace mskf {

// ---- constants handed to the front-end kernels by value -------------------------------
struct FeConst {
    int rows, cols, levels;
    int lvl_rows[MSKF_MAX_LEVELS], lvl_cols[MSKF_MAX_LEVELS];
    unsigned lvl_off[MSKF_MAX_LEVELS];
    unsigned pyr_bytes;  // per image, all levels
    int pitch;           // row pitch of EVERY pyramid level = level-0 width (level l sits at rows sum(lvl_rows[<l]) of one
                         // pitch-wide image), so that the KLT's row offsets are compile-time immediates for the standard widths
    int klt_win, klt_max_iters;
    double klt_eps2, klt_min_eig;
    int grid_row, grid_col, grid_min, grid_max, grid_h, grid_w, n_cells, n_cells_all;
    int det_rows, det_cols, det_cell_h, det_cell_w, det_cells;
    int fast_threshold;
    double detection_threshold;
    int max_f;    // capacity of a grid / tracked list
    int cap_k;    // capacity of a KLT work list (>= det_cells)
    int cam_model[2];
    double K[2][4], D[2][4];
    double R01[9];      // R_cam0_cam1 = R_cam1_imu^T R_cam0_imu (image_processor.cpp:544)
    double E[9];        // essential matrix (image_processor.cpp:591)
    double stereo_gate; // stereo_threshold * norm_pixel_unit (image_processor.cpp:606,615)
    int compat_stale;
    int use_ransac;
    double ransac_threshold;
};

// per-stream, per-step descriptor written by the host before each front-end step
struct FeStep {
    int active, is_first, slot, pad;
    double t;
    double H0[9];  // K R_p_c K^-1 for cam0 (image_processor.cpp:340)
    double R0[9], R1[9];  // cam0_R_p_c, cam1_R_p_c (integrateImuData, image_processor.cpp:882-883): twoPointRansac
};

struct GridSoA {            // one grid of one stream, publish order (cell asc, insertion order)
    unsigned long long *id;
    float *response;
    int *lifetime;
    float2 *cam0, *cam1;
    int *cell;
};

struct FeBuffers {
    uint8_t *pyr[3];        // [S][pyr_bytes] x3 : cam0 slot0, cam0 slot1, cam1
    uint8_t *staging;       // [2 slots][S][2][rows*cols]
    const uint8_t **src0, **src1;  // [S] device pointers to this step's level-0 sources
    FeStep *step;           // [S]
    // grids: prev and curr (index by ping-pong flag gslot[s])
    unsigned long long *g_id[2];
    float *g_resp[2];
    int *g_life[2];
    float2 *g_cam0[2], *g_cam1[2];
    int *g_cell[2];
    int *g_n[2];            // [S]
    int *gslot;             // [S] which grid buffer is "prev"
    unsigned long long *next_id;  // [S]
    // KLT work lists
    float2 *k_a, *k_b;      // [S][cap_k]
    uint8_t *k_status;      // [S][cap_k]
    uint8_t *k_skip;        // [S][cap_k] new-feature candidates whose cell has no vacancy (not matched)
    int *k_n;               // [S]
    int *k_idx;             // [S][cap_k] new-feature candidates that ARE matched (cell with a vacancy), compacted by fe_sieve
    int *k_nm;              // [S] their number
    // tracked meta carried through the temporal track
    unsigned long long *t_id;  // [S][max_f]
    int *t_life;               // [S][max_f]
    float2 *t_p0, *t_p1;       // [S][max_f] previous cam0 / cam1 points of the tracked features (twoPointRansac)
    unsigned *track_calls;     // [S] trackFeatures calls that reached the RANSAC stage (seeds its sampler)
    // detector
    unsigned long long *det_best;  // [S][det_cells] packed (score bits << 32 | ~raster)
    uint8_t *det_occ;              // [S][det_cells]
    float *nf_resp;                // [S][det_cells] responses in detect order
    float *in_resp;                // [S][cap_k] responses of the stereo inliers (fe_finish)
    int *nf_n;                     // [S] number detected
    // outputs
    mskf_feature *msg;      // [S][max_f] current frame measurements
    int *msg_n;             // [S]
    mskf_feature *stale;    // [S][max_f] high-water copy emulating the never-cleared vector (F4)
    int *stale_hw;          // [S] high-water mark
    long long *msg_total;   // [S] total length of the reference's growing vector
    mskf_tracking_info *info;  // [S]
    uint8_t *dbg_score;     // optional [rows*cols] FAST score map of stream 0 (tests only)
    double *work;           // [S][MSKF_PROF_TAGS] algorithmic bytes / flops done, per kernel class (bench roofline)
};

}  // namespace mskf

struct BeBuffers;  // backend.cu

struct HostImu {
    double t, w[3], a[3];
};

struct HostStream {
    // front end (image_processor.cpp:205-211, 850-889)
    bool fe_first = true;
    std::vector<HostImu> fe_imu;
    double fe_prev_t = 0, fe_curr_t = 0;
    int slot = 0;
    bool pending = false;      // a stereo pair is staged
    bool published = false;    // front end produced a message not yet consumed by the back end
    double pending_t = 0;
    const uint8_t *src0 = nullptr, *src1 = nullptr;
    // back end (msckf_vio.cpp:190-241, 377-407)
    std::vector<HostImu> be_imu;
    bool gravity_set = false, be_first = true;
    double be_time = 0;        // mirror of imu_state.time
    double msg_t = 0;
};

struct mskf_handle {
    mskf_config cfg;
    int S = 0, device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // host uploads run on their own stream into a double-buffered landing area, so the upload of
    // frame k+1 overlaps the kernels of frame k (mskf_push_stereo* -> stage_*; engine.cu)
    // The back end of frame k runs on its own stream so that the front end of frame k+1 overlaps it
    // (the only shared data is the CameraMeasurement: ev_msg_ready / ev_msg_consumed).
    cudaStream_t be_stream = nullptr;
    cudaStream_t cur = nullptr;        // stream the MSKF_LAUNCH profiler brackets are recorded on
    cudaEvent_t ev_msg_ready = nullptr, ev_msg_consumed = nullptr, ev_join = nullptr;
    bool msg_consumed_valid = false;
    bool overlap = true;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr};
    bool consumed_valid[2] = {false, false};
    int stage_cur = 0;          // slot the next pushes land in
    bool stage_dirty = false;   // copies issued into stage_cur since the last step
    bool stage_waited = false;  // the copy stream already waited for the slot to be consumed
    std::string err;
    mskf::FeConst fc;
    mskf::FeBuffers fb;
    BeBuffers *bb = nullptr;
    std::vector<HostStream> hs;
    // pinned staging
    uint8_t *h_stage = nullptr;        // [S][2][rows*cols]
    mskf::FeStep *h_step = nullptr;    // [S]
    const uint8_t **h_src = nullptr;   // [2][S]
    void *h_be_step = nullptr;         // backend step descriptors (pinned)
    std::vector<void *> allocs;
    long long launches = 0;            // kernels launched so far (bench: gpu_launches)
    // per-kernel CUDA-event timing on the launching stream (bench.py's roofline leg)
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;  // pairs (start, stop)
    std::vector<int> prof_tag;
    size_t prof_used = 0;
    double *d_work = nullptr;          // [S][MSKF_PROF_TAGS]
    double work_host[MSKF_PROF_TAGS] = {0};  // work of kernels whose size the host knows
    double prof_ms[MSKF_PROF_TAGS] = {0};
    long long prof_n[MSKF_PROF_TAGS] = {0};
};

// kernel tags for the profiler
enum {
    PK_PYR_L1 = 0, PK_PYR_LN, PK_KLT_TEMPORAL, PK_KLT_STEREO, PK_KLT_NEW, PK_DETECT, PK_FE_BOOK,
    PK_BE_PROPAGATE, PK_BE_AUGMENT, PK_BE_ADD_OBS, PK_BE_SELECT, PK_BE_TRIANGULATE, PK_BE_LAYOUT, PK_BE_FEATURE_JAC,
    PK_BE_STACK, PK_BE_GRAM, PK_BE_GEMM_PHT, PK_BE_GEMM_S, PK_BE_CHOL, PK_BE_GEMM_W, PK_BE_APPLY, PK_BE_GEMM_PUPD,
    PK_BE_PRUNE_FINISH, PK_BE_FINISH, PK_BE_FEATURE_JAC_PRUNE, PK_BE_GRAM_PRUNE, PK_BE_PCHOL, PK_COUNT
};
const char *mskf_prof_name(int tag);
void prof_begin(mskf_handle *h, int tag);
void prof_end(mskf_handle *h);
void prof_collect(mskf_handle *h);

// launch + count + optional event bracket
// Per-step descriptors (FeStep / BeStep arrays, image pointers, IMU rows: a few KB per handle) are FETCHED by a
// small kernel from the page-locked host ring (device-visible through UVA) instead of copied by cudaMemcpyAsync:
// a copy-engine transfer queues behind the 46 MB frame-set uploads of the other handles (up to 3 ms at 256
// streams), which stalled every handle's step behind the fleet's uploads; SM loads over PCIe do not.
struct FetchSeg {
    void *dst;
    const void *src;
    unsigned width8;  // 8-byte words per row
    unsigned pitch8;  // row pitch of dst and src in 8-byte words
    unsigned rows;
};
struct FetchArgs {
    FetchSeg seg[4];
    int n;
};
static __global__ void __launch_bounds__(256) desc_fetch_kernel(FetchArgs a) {
    for (int k = 0; k < a.n; ++k) {
        const FetchSeg sg = a.seg[k];
        const unsigned long long *src = (const unsigned long long *)sg.src;
        unsigned long long *dst = (unsigned long long *)sg.dst;
        const unsigned total = sg.width8 * sg.rows;
        for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
            const unsigned r = i / sg.width8, c = i - r * sg.width8;
            dst[(size_t)r * sg.pitch8 + c] = src[(size_t)r * sg.pitch8 + c];
        }
    }
}
// one pass over the largest segment (a load over PCIe costs ~1.5 us: a thread should issue one, not eight in a row)
static inline int fetch_grid(const FetchArgs &a) {
    unsigned most = 1;
    for (int k = 0; k < a.n; ++k) most = a.seg[k].width8 * a.seg[k].rows > most ? a.seg[k].width8 * a.seg[k].rows : most;
    const unsigned g = (most + 255) / 256;
    return (int)(g < 1 ? 1 : g > 64 ? 64 : g);
}
static inline FetchSeg fetch_seg(void *dst, const void *src, size_t bytes) {
    return FetchSeg{dst, src, (unsigned)(bytes / 8), (unsigned)(bytes / 8), 1u};
}

#define MSKF_LAUNCH(h, tag, ...)        \
    do {                                \
        prof_begin((h), (tag));         \
        __VA_ARGS__;                    \
        prof_end((h));                  \
        (h)->launches++;                \
    } while (0)

// frontend.cu
int fe_create(mskf_handle *h);
int stage_begin_consume(mskf_handle *h);  // compute stream waits for the uploads of this step
int stage_end_consume(mskf_handle *h);    // level 0 has been landed: the slot may be overwritten
int fe_step(mskf_handle *h, bool any_first, int max_prev, int n_active);
int fe_op_detect(mskf_handle *t, const float *occ, int n_occ, float *out_xy, double *out_resp, int cap, int *n,
                 uint8_t *score_map);
int fe_op_klt(mskf_handle *t, const float *pts_a, float *pts_b, uint8_t *status, int n);
// backend.cu
int be_create(mskf_handle *h);
void be_destroy(mskf_handle *h);
int be_step(mskf_handle *h, const std::vector<int> &streams, const mskf_feature *inject, int n_inject,
            int inject_stream, double inject_t);
int be_init_gravity(mskf_handle *h, int s);
int be_get_state(mskf_handle *h, int s, mskf_state *out);
int be_get_cam_states(mskf_handle *h, int s, mskf_cam_state *out, int cap, int *n);
int be_get_cov(mskf_handle *h, int s, double *out, int cap, int *dim);
int be_reset(mskf_handle *h, int s);
int be_get_map(mskf_handle *h, int s, long long *ids, int *init, double *pos, int *nobs, int cap, int *n);
int be_op_triangulate(mskf_handle *t, int n_cam, const double *cam_q, const double *cam_p, int n_feat, const unsigned *mask,
                      const double *obs, double *pos, int *ok);
int be_op_update(mskf_handle *t, int n_cam, int m, const double *H, const double *r, const double *P, double *dx, double *Pn);
int be_debug_update_dims(mskf_handle *h, int *out6);
int be_debug_last_gram(mskf_handle *h, int s, double *G, int cap, int *m, int *k, long long *cam_ids, int *valid);
int be_get_poses(mskf_handle *h, double *out, int cap_streams, int lag);

template <typename T>
int dev_alloc(mskf_handle *h, T **p, size_t n) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, n * sizeof(T));
    if (e != cudaSuccess) {
        h->err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return MSKF_ERR_CUDA;
    }
    // the engine's streams are non-blocking (not ordered against the legacy default stream the memset
    // runs on), so the clear is completed here, before any kernel can touch the buffer
    cudaMemset(q, 0, n * sizeof(T));
    cudaStreamSynchronize(cudaStreamLegacy);
    h->allocs.push_back(q);
    *p = (T *)q;
    return MSKF_OK;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per kernel for the whole process, not per handle: every
// kernel that needs more than 48 KB gets the device's opt-in maximum (only ever the same value), so a
// handle with a smaller configuration cannot lower the limit under a live larger one.
template <typename F>
int smem_optin(mskf_handle *h, F *kernel, size_t need) {
    int optin = 0;
    cudaError_t e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, h->device);
    cudaFuncAttributes fa;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, kernel);
    if (e == cudaSuccess) optin -= (int)fa.sharedSizeBytes;  // static + dynamic share the per-block limit
    if (e == cudaSuccess && need > (size_t)optin) {
        h->err = "configuration needs more shared memory per block than the device offers";
        return MSKF_ERR_CAPACITY;
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e != cudaSuccess) {
        h->err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        return MSKF_ERR_CUDA;
    }
    return MSKF_OK;
}
