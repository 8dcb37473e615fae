/*
 * msckf_b200_presets.h — the two named parameter sets (header-only data).
 *
 *  "ref"   exactly what the reference code runs: 4 pyramid levels
 *          (image_processor.cpp:232), KLT window 15 / 30 iterations (:410,:569),
 *          CornerDetector(30, 47, fast_threshold) (:132), config/app_imgproc.yaml,
 *          config/app_msckfvio.yaml and config/camchain-imucam-euroc.yaml values.
 *  "bench" BASELINE.json configs 2-4: KLT window 21, grid 4x5 with up to 15 features per
 *          cell (~300 features), max_cam_state_size 30.
 *  "stress" BASELINE.json config 5: 1280x1024, 6 levels, up to 1000 features.
 */
#ifndef MSCKF_B200_PRESETS_H
#define MSCKF_B200_PRESETS_H

#include <string.h>

#include "msckf_b200.h"

static inline int mskf_fill_preset(mskf_config *c, const char *preset) {
    static const double T_cam0_imu[16] = {
        0.014865542981794, 0.999557249008346, -0.025774436697440, 0.065222909535531,
        -0.999880929698575, 0.014967213324719, 0.003756188357967, -0.020706385492719,
        0.004140296794224, 0.025715529947966, 0.999660727177902, -0.008054602460030,
        0, 0, 0, 1.0};
    static const double T_cn_cnm1[16] = {
        0.999997256477881, 0.002312067192424, 0.000376008102415, -0.110073808127187,
        -0.002317135723281, 0.999898048506644, 0.014089835846648, 0.000399121547014,
        -0.000343393120525, -0.014090668452714, 0.999900662637729, -0.000853702503357,
        0, 0, 0, 1.0};
    static const double I4[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    static const double k0[4] = {458.654, 457.296, 367.215, 248.375};
    static const double d0[4] = {-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05};
    static const double k1[4] = {457.587, 456.134, 379.999, 255.238};
    static const double d1[4] = {-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05};
    if (!c || !preset) return MSKF_ERR_ARG;
    memset(c, 0, sizeof(*c));
    c->img_rows = 480; c->img_cols = 752;
    c->pyramid_levels = 4;
    c->klt_win = 15; c->klt_max_iters = 30; c->klt_eps = 0.01; c->klt_min_eig = 1e-4;
    c->grid_row = 4; c->grid_col = 5; c->grid_min_feature_num = 3; c->grid_max_feature_num = 4;
    c->det_rows = 30; c->det_cols = 47;
    c->fast_threshold = 10; c->detection_threshold = 10.0;
    c->stereo_threshold = 5.0; c->ransac_threshold = 3.0;
    c->use_ransac = 0; c->compat_stale_features = 1; c->fix_prev_image_alias = 0;
    c->cam0_model = MSKF_MODEL_RADTAN; c->cam1_model = MSKF_MODEL_RADTAN;
    memcpy(c->cam0_intrinsics, k0, sizeof k0); memcpy(c->cam0_distortion, d0, sizeof d0);
    memcpy(c->cam1_intrinsics, k1, sizeof k1); memcpy(c->cam1_distortion, d1, sizeof d1);
    memcpy(c->T_cam0_imu, T_cam0_imu, sizeof T_cam0_imu);
    memcpy(c->T_cn_cnm1, T_cn_cnm1, sizeof T_cn_cnm1);
    memcpy(c->T_imu_body, I4, sizeof I4);
    c->frame_rate = 20; c->max_cam_state_size = 20; c->chi2_mode = MSKF_CHI2_Q05;
    c->position_std_threshold = 8.0; c->rotation_threshold = 0.2618; c->translation_threshold = 0.4;
    c->tracking_rate_threshold = 0.5; c->feature_translation_threshold = -1.0;
    c->noise_gyro = 0.005; c->noise_acc = 0.05; c->noise_gyro_bias = 0.001; c->noise_acc_bias = 0.01;
    c->noise_feature = 0.035;
    c->cov_velocity = 0.25; c->cov_gyro_bias = 0.01; c->cov_acc_bias = 0.01;
    c->cov_ext_rot = 3.0462e-4; c->cov_ext_trans = 2.5e-5;
    c->max_jacobian_rows = 1500;
    if (strcmp(preset, "ref") == 0) return MSKF_OK;
    if (strcmp(preset, "bench") == 0) {
        c->klt_win = 21;
        c->grid_min_feature_num = 12; c->grid_max_feature_num = 15;
        c->max_cam_state_size = 30;
        return MSKF_OK;
    }
    if (strcmp(preset, "stress") == 0) {
        double sx = 1280.0 / 752.0, sy = 1024.0 / 480.0;
        c->img_rows = 1024; c->img_cols = 1280;
        c->pyramid_levels = 6;
        c->klt_win = 21;
        c->grid_row = 8; c->grid_col = 10;
        c->grid_min_feature_num = 10; c->grid_max_feature_num = 12;  /* 80 cells -> <= 960 */
        c->det_rows = 64; c->det_cols = 80;
        c->max_cam_state_size = 30;
        c->cam0_intrinsics[0] *= sx; c->cam0_intrinsics[2] *= sx; c->cam0_intrinsics[1] *= sy; c->cam0_intrinsics[3] *= sy;
        c->cam1_intrinsics[0] *= sx; c->cam1_intrinsics[2] *= sx; c->cam1_intrinsics[1] *= sy; c->cam1_intrinsics[3] *= sy;
        return MSKF_OK;
    }
    return MSKF_ERR_ARG;
}

#endif /* MSCKF_B200_PRESETS_H */
