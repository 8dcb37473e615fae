/*
 * msckf_b200.h — C ABI of the B200-native stereo-MSCKF hot path.
 *
 * Drop-in boundary for mfkiwl/msckf_stereo_c.  The reference has no FFI layer: its
 * boundary is the C++ class API of ImageProcessor / MsckfVio / System plus the plain
 * message structs.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference tree).  Plain pointers and sizes only: no C++ types,
 * no torch types.  One handle drives `n_streams` independent stereo+IMU streams on one
 * GPU; calls on one handle must be serialised by the caller (the reference is
 * single-threaded too: apps/run_euroc_single_thread.cpp:189-254).
 *
 * There is no CPU fallback behind this ABI: every compute entry point launches CUDA
 * kernels and returns MSKF_ERR_CUDA if the device is missing or a launch fails.
 */
#ifndef MSCKF_B200_H
#define MSCKF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSKF_OK 0
#define MSKF_ERR_ARG (-1)
#define MSKF_ERR_CUDA (-2)
#define MSKF_ERR_CAPACITY (-3)
#define MSKF_ERR_STATE (-4)

#define MSKF_MODEL_RADTAN 0
#define MSKF_MODEL_EQUIDISTANT 1

/* chi-square table selection for gatingTest (msckf_vio.cpp:180-185): the table the
 * reference reads lives in vikit_cg and is named "p95", the line it replaces computes
 * the 0.05 quantile.  0 = 0.05 quantile (default), 1 = 0.95 quantile. */
#define MSKF_CHI2_Q05 0
#define MSKF_CHI2_Q95 1

#define MSKF_MAX_LEVELS 8

/* Everything the reference reads from camchain-imucam-euroc.yaml, app_imgproc.yaml and
 * app_msckfvio.yaml (image_processor.cpp:52-124, msckf_vio.cpp:58-162), plus the values
 * the reference hard-codes made explicit. */
typedef struct mskf_config {
    /* image geometry */
    int32_t img_rows, img_cols;
    /* front end: image_processor.cpp:232 (levels), :410/:569 (win, iters), :132 (detector) */
    int32_t pyramid_levels;       /* number of levels incl. level 0; reference hard-codes 4 */
    int32_t klt_win;              /* odd window edge; reference literal 15 */
    int32_t klt_max_iters;        /* reference literal 30 */
    double  klt_eps;              /* track_precision 0.01 */
    double  klt_min_eig;          /* minimum normalised eigenvalue, 1e-4 */
    int32_t grid_row, grid_col;
    int32_t grid_min_feature_num, grid_max_feature_num;
    int32_t det_rows, det_cols;   /* CornerDetector(30, 47, thr) fine occupancy grid */
    int32_t fast_threshold;
    double  detection_threshold;  /* minimum Shi-Tomasi response of a detected corner */
    double  stereo_threshold;     /* epipolar gate, pixels (image_processor.cpp:615) */
    double  ransac_threshold;
    int32_t use_ransac;           /* twoPointRansac is dead code in the reference: default 0 */
    int32_t compat_stale_features;/* reproduce SURVEY F4 (message never cleared): default 1 */
    /* calibration (Kalibr camchain) */
    int32_t cam0_model, cam1_model;
    double  cam0_intrinsics[4], cam0_distortion[4];
    double  cam1_intrinsics[4], cam1_distortion[4];
    double  T_cam0_imu[16];       /* cam0/T_cam_imu, row-major 4x4 (config_io.h:66-79) */
    double  T_cn_cnm1[16];        /* cam1/T_cn_cnm1 */
    double  T_imu_body[16];
    /* back end (app_msckfvio.yaml) */
    double  frame_rate;
    int32_t max_cam_state_size;   /* 5..31 (app_msckfvio.yaml: 20; BASELINE bench config: 30) */
    int32_t chi2_mode;
    double  position_std_threshold;
    double  rotation_threshold, translation_threshold, tracking_rate_threshold;
    double  feature_translation_threshold;
    double  noise_gyro, noise_acc, noise_gyro_bias, noise_acc_bias, noise_feature; /* std devs */
    double  initial_velocity[3];
    double  cov_velocity, cov_gyro_bias, cov_acc_bias, cov_ext_rot, cov_ext_trans;
    int32_t max_jacobian_rows;    /* row cap of removeLostFeatures, msckf_vio.cpp:1009 (1500) */
    /* SURVEY-style defect F6: image_processor.cpp:192 aliases cam0_prev_img_ptr to
     * cam0_curr_img_ptr (never re-allocated), so from the second frame on prev and curr
     * time stamps are equal and integrateImuData (:881) always yields dt = 0, R = I.
     * 0 = mirror the reference (default), 1 = use the real previous time stamp. */
    int32_t fix_prev_image_alias;
} mskf_config;

/* FeatureMeasurement, include/common/data_msg.h:30-37 */
typedef struct mskf_feature {
    uint32_t id;
    uint32_t pad;
    double u0, v0, u1, v1;
} mskf_feature;

/* TrackingInfo, include/common/data_msg.h:47-54 */
typedef struct mskf_tracking_info {
    double  time_stamp;
    int32_t before_tracking, after_tracking, after_matching, after_ransac;
} mskf_tracking_info;

/* One row of the front end's grid (FeatureMetaData, image_processor.h:75-81), pixel units */
typedef struct mskf_grid_feature {
    uint64_t id;
    float    response;
    int32_t  lifetime;
    float    cam0_x, cam0_y, cam1_x, cam1_y;
    int32_t  cell;
    int32_t  pad;
} mskf_grid_feature;

/* IMUState + filter bookkeeping (include/common/imu_state.h:28-88) */
typedef struct mskf_state {
    double  time;
    int64_t id;
    double  orientation[4];   /* JPL quaternion x y z w, world -> imu */
    double  position[3];
    double  velocity[3];
    double  gyro_bias[3];
    double  acc_bias[3];
    double  R_imu_cam0[9];    /* row-major */
    double  t_cam0_imu[3];
    double  gravity[3];
    int32_t n_cam_states;
    int32_t cov_dim;          /* 21 + 6 * n_cam_states */
    int32_t is_gravity_set;
    int32_t n_map_features;
    double  tracking_rate;
    double  T_b_w[16];        /* body pose published by MsckfVio::publish, msckf_vio.cpp:1242-1246 */
    int64_t n_updates;        /* measurementUpdate calls so far */
    int64_t n_resets;         /* onlineReset count */
} mskf_state;

typedef struct mskf_cam_state {
    int64_t id;
    double  time;
    double  orientation[4];
    double  position[3];
} mskf_cam_state;

typedef struct mskf_handle mskf_handle;

/* Fill `cfg` with preset "ref" (exactly what the reference code runs: L=4, win 15, 30
 * iterations, grid 4x5 min 3 max 4, N<=20, EuRoC calibration) or "bench"
 * (BASELINE.json configs 2-4: win 21, grid cap 15 => ~300 features, N=30).
 * Replaces: YAML loading in image_processor.cpp:52-124 and msckf_vio.cpp:58-162. */
int mskf_default_config(mskf_config *cfg, const char *preset);

/* Replaces: System::System (system.cpp:12-34) = ImageProcessor ctor+initialize
 * (image_processor.cpp:32-42,126-137) and MsckfVio ctor+initialize (msckf_vio.cpp:51-56,164-188),
 * once per stream.  `device` is the CUDA ordinal. */
int mskf_create(const mskf_config *cfg, int n_streams, int device, mskf_handle **out);
void mskf_destroy(mskf_handle *h);
const char *mskf_last_error(const mskf_handle *h);

/* Launch all work of this handle on `cuda_stream` (a cudaStream_t); default: a private stream. */
int mskf_set_cuda_stream(mskf_handle *h, void *cuda_stream);

/* Replaces: System::imu_callback (system.cpp:45-48) -> ImageProcessor::imuCallback
 * (image_processor.cpp:205-211) + MsckfVio::imuCallback (msckf_vio.cpp:190-207). */
int mskf_push_imu(mskf_handle *h, int stream, double t, const double w[3], const double a[3]);
/* One half only, for callers that drive the two classes separately as the reference allows:
 * MSKF_IMU_FRONTEND = ImageProcessor::imuCallback (image_processor.cpp:205-211: buffered for
 * integrateImuData, ignored before the first image), MSKF_IMU_BACKEND = MsckfVio::imuCallback
 * (msckf_vio.cpp:190-207: buffered for batchImuProcessing; the 200th sample initialises gravity and the
 * gyro bias, :198-204).  mskf_push_imu is both. */
#define MSKF_IMU_FRONTEND 1
#define MSKF_IMU_BACKEND 2
int mskf_push_imu_to(mskf_handle *h, int stream, int halves, double t, const double w[3], const double a[3]);

/* Stage one stereo pair for `stream` from HOST memory (copied before return; the caller
 * may free its buffers, as with the by-value copy at image_processor.cpp:144-145). */
int mskf_push_stereo(mskf_handle *h, int stream, double t, const uint8_t *cam0,
                     const uint8_t *cam1, int rows, int cols, int stride);
/* Same, but the images already live in DEVICE memory (tightly packed rows*cols). */
int mskf_push_stereo_device(mskf_handle *h, int stream, double t, const uint8_t *d_cam0,
                            const uint8_t *d_cam1);

/* Fleet variants of the three calls above (same semantics, fewer host calls): `samples` holds n rows
 * {t, wx, wy, wz, ax, ay, az} for `stream`, or, with stream == -1, [n_streams][n][7]; the stereo batches
 * take one image per stream at cam0 + s * stream_stride (bytes), time stamps t[n_streams]. */
int mskf_push_imu_batch(mskf_handle *h, int stream, int n, const double *samples);
/* The host -> device copy of mskf_push_stereo_batch is ASYNCHRONOUS (the engine's copy stream; one
 * contiguous copy when cam1 == cam0 + rows*cols and stream_stride == 2*rows*cols, the landing area's own
 * layout): the images should be page-locked (mskf_host_alloc) and must stay unchanged until
 * mskf_wait_uploads (or mskf_sync) returns. */
int mskf_push_stereo_batch(mskf_handle *h, const double *t, const uint8_t *cam0, const uint8_t *cam1, size_t stream_stride);
/* Blocks until every upload issued by mskf_push_stereo* on this handle has left the caller's buffers. */
int mskf_wait_uploads(mskf_handle *h);
/* Page-locked host memory for the image ring of a fleet loader (examples/run_euroc_fleet.cpp), so that a host
 * application needs no CUDA headers; no handle: valid for every engine of the process. */
int mskf_host_alloc(void **out, size_t bytes);
/* Same, write-combined (cudaHostAllocWriteCombined): for buffers the host only WRITES and the GPU's copy engine
 * reads (frame rings).  The DMA reads do not snoop the CPU caches, which raises the upload rate when several GPUs
 * pull from the same host at once; CPU reads of such memory are slow. */
int mskf_host_alloc_wc(void **out, size_t bytes);
void mskf_host_free(void *p);
int mskf_push_stereo_device_batch(mskf_handle *h, const double *t, const uint8_t *d_cam0, const uint8_t *d_cam1,
                                  size_t stream_stride);

/* Replaces: System::stereo_callback (system.cpp:40-43) -> ImageProcessor::stereoCallback
 * (image_processor.cpp:139-203) for every stream with a staged pair. */
int mskf_frontend_step(mskf_handle *h);
/* Replaces: System::backend_callback (system.cpp:50-54) -> MsckfVio::featureCallback
 * (msckf_vio.cpp:306-375) for every stream whose front end published this step. */
int mskf_backend_step(mskf_handle *h);
/* frontend_step + backend_step */
int mskf_step(mskf_handle *h);
/* Block until all launched work of this handle has finished. */
int mskf_sync(mskf_handle *h);

/* Split-phase back end: run featureCallback for one stream on caller-supplied
 * measurements ("identical feature inputs" for the 1e-9 EKF parity test). */
int mskf_backend_step_features(mskf_handle *h, int stream, double t, const mskf_feature *f, int n);

/* Replaces: reading ImageProcessor::feature_msg_ptr_ (image_processor.h:42).  Returns the
 * message of the last front-end step for `stream`; in compat_stale_features mode the
 * message carries the stale tail exactly as the reference's never-cleared vector does. */
int mskf_get_features(mskf_handle *h, int stream, mskf_feature *out, int cap, int *n, double *t);
/* The same message without materialising its value-initialised tail: entries [0, *n_head) are every entry
 * that is not {id 0, zeros} (this frame's measurements followed by stale leftovers of longer earlier
 * frames); *n_total is the length of the reference's vector, which grows by the frame's feature count every
 * frame (image_processor.cpp:1157-1164) and so needs 64 bits and O(run length) memory if copied out whole. */
int mskf_get_features_head(mskf_handle *h, int stream, mskf_feature *out, int cap, int *n_head,
                           long long *n_total, double *t);
int mskf_get_tracking_info(mskf_handle *h, int stream, mskf_tracking_info *out);
/* The front end's grid after the step (what became prev_features_ptr), publish order. */
int mskf_get_grid(mskf_handle *h, int stream, mskf_grid_feature *out, int cap, int *n);
/* Pyramid level `level` of cam 0/1 of the last processed frame, tightly packed. */
int mskf_get_pyramid(mskf_handle *h, int stream, int cam, int level, uint8_t *out, int cap,
                     int *rows, int *cols);

/* Replaces: MsckfVio state access / publish (msckf_vio.cpp:1238-1305). */
int mskf_get_state(mskf_handle *h, int stream, mskf_state *out);
int mskf_get_cam_states(mskf_handle *h, int stream, mskf_cam_state *out, int cap, int *n);
/* Row-major cov_dim x cov_dim copy of state_server.state_cov. */
int mskf_get_covariance(mskf_handle *h, int stream, double *out, int cap, int *dim);
/* Replaces: MsckfVio::resetCallback (msckf_vio.cpp:243-304). */
int mskf_reset(mskf_handle *h, int stream);

/* ---- stand-alone operators (the vikit_cg primitives of the path), device-side, batched.
 * Used by the parity tests and the per-kernel benches; host pointers in, host pointers out. */

/* cg::pyr_down chain (image_processor.cpp:229-244): level 0 given, levels 1..L-1 returned
 * concatenated in `out` (sizes ((c+1)/2, (r+1)/2) per level). */
int mskf_op_pyramid(mskf_handle *h, const uint8_t *img, int n_images, int rows, int cols,
                    int levels, uint8_t *out);
/* CornerDetector::set_grid_position + detect_features (image_processor.cpp:647,657). */
int mskf_op_detect(mskf_handle *h, const uint8_t *img, int rows, int cols, const float *occupied_xy,
                   int n_occupied, float *out_xy, double *out_response, int cap, int *n);
/* cg::optical_flow_multi_level (image_processor.cpp:410,569): pyramids are built on the
 * device from the two level-0 images. pts_b holds the initial guess on entry. */
int mskf_op_klt(mskf_handle *h, const uint8_t *img_a, const uint8_t *img_b, int rows, int cols,
                const float *pts_a, float *pts_b, uint8_t *status, int n);

/* measurementUpdate (msckf_vio.cpp:778-907: QR compression, S, gain, delta_x, covariance update)
 * as a stand-alone operator for the "identical inputs" parity test: H is m x n row-major with
 * n = 21 + 6 n_cam (its 21 IMU columns are zero, msckf_vio.cpp:709-712), r has m entries, P is
 * n x n; observation noise comes from the configuration.  Outputs delta_x (n) and the posterior P. */
int mskf_op_ekf_update(mskf_handle *h, int n_cam, int m, const double *H, const double *r, const double *P,
                       double *out_delta_x, double *out_P);

/* Feature::checkMotion + Feature::initializePosition (feature.hpp:257-287, 289-450) as a stand-alone
 * operator: n_cam camera states in ascending state id (orientation JPL [x y z w], position), n_feat
 * features, each observed by the camera states whose bit is set in obs_mask[f]; obs is
 * [n_feat][n_cam][4] = (u0, v0, u1, v1), normalised coordinates.  out_ok[f] = 1 iff checkMotion passed
 * and initializePosition returned true; out_position[f] is what initializePosition leaves in
 * Feature::position (written whenever it ran, valid or not; 0 otherwise).  The arithmetic and the
 * summation order are the reference's, so results are comparable bit for bit. */
int mskf_op_triangulate(mskf_handle *h, int n_cam, const double *cam_orientation, const double *cam_position,
                        int n_feat, const unsigned *obs_mask, const double *obs, double *out_position,
                        int *out_ok);

/* ---- instrumentation (bench.py / tests; no reference counterpart) ------------------------ */
/* Kernels launched by this handle so far. */
long long mskf_launch_count(const mskf_handle *h);
/* Number of measurements the last front-end step published for `stream` (the reference's
 * curr_ids.size() in publish(), image_processor.cpp:1146-1164), without the stale tail. */
int mskf_get_n_published(mskf_handle *h, int stream, int *n);
/* Poses T_b_w (row-major 4x4, msckf_vio.cpp:1242-1246) of all streams in one device->host copy. */
int mskf_get_poses(mskf_handle *h, double *out_T_b_w, int cap_streams);
/* Poses of the back-end step before the latest one (already on the host while the latest is in
 * flight): lets a fleet driver read every step's result without draining the pipeline. */
int mskf_get_poses_prev(mskf_handle *h, double *out_T_b_w, int cap_streams);
/* The back end runs on its own CUDA stream and overlaps the next frame's front end.  mskf_join makes
 * the handle's (front-end) stream wait for the back-end work launched so far, so that an event
 * recorded on that stream afterwards brackets both; mskf_set_overlap(h, 0) serialises the two
 * halves (used to time kernels in isolation). */
int mskf_join(mskf_handle *h);
int mskf_set_overlap(mskf_handle *h, int on);
/* CUDA-event time per kernel class, recorded on the launching stream while enabled.
 * mskf_profile_read returns 1 when `tag` is past the last class. */
int mskf_profile_enable(mskf_handle *h, int on);
int mskf_profile_read(mskf_handle *h, int tag, const char **name, double *ms, long long *count);
/* Algorithmic work (bytes for the front-end classes, flops for the EKF classes) done by kernel class
 * `tag` since mskf_profile_enable(h, 1): the numerator of bench.py's roofline figures. */
int mskf_get_work(mskf_handle *h, int tag, double *total);
/* The back end's feature map (MsckfVio::map_server) in ascending feature id. */
int mskf_debug_get_map(mskf_handle *h, int stream, long long *ids, int *is_initialized, double *position,
                       int *n_observations, int cap, int *n);
/* Per stream {rows m, active columns k, listed features} of the lost-feature update and of the prune
 * update of the last back-end step: out[n_streams][2][3]. */
int mskf_debug_update_dims(mskf_handle *h, int *out);
/* The stacked system of the latest measurementUpdate of a stream (msckf_vio.cpp:778), i.e. the output of
 * measurementJacobian + featureJacobian (null-space projection) + gatingTest + stacking (:610-775, :909-1024),
 * in the basis-independent form the engine keeps: G = [H r]^T [H r] over the k active camera columns
 * ((k+1) x (k+1), row-major, symmetric), m = stacked rows, cam_ids[g] = id of the camera state behind columns
 * 6g .. 6g+5 (at most 32 groups).  *valid = 0 when m <= k (no compression ran: G is not formed). */
int mskf_debug_last_gram(mskf_handle *h, int stream, double *G, int cap, int *m, int *k, long long *cam_ids,
                         int *valid);
/* mskf_op_detect that also returns the per-pixel FAST score map (0 = not a corner). */
int mskf_debug_detect_scores(mskf_handle *h, const uint8_t *img, int rows, int cols, float *out_xy,
                             double *out_response, int cap, int *n, uint8_t *score_map);

#ifdef __cplusplus
}
#endif
#endif /* MSCKF_B200_H */
