"""ctypes mirror of include/msckf_b200.h (the C-ABI structs)."""
import ctypes as C


class Config(C.Structure):
    _fields_ = [
        ("img_rows", C.c_int32), ("img_cols", C.c_int32),
        ("pyramid_levels", C.c_int32), ("klt_win", C.c_int32), ("klt_max_iters", C.c_int32),
        ("klt_eps", C.c_double), ("klt_min_eig", C.c_double),
        ("grid_row", C.c_int32), ("grid_col", C.c_int32),
        ("grid_min_feature_num", C.c_int32), ("grid_max_feature_num", C.c_int32),
        ("det_rows", C.c_int32), ("det_cols", C.c_int32),
        ("fast_threshold", C.c_int32),
        ("detection_threshold", C.c_double), ("stereo_threshold", C.c_double), ("ransac_threshold", C.c_double),
        ("use_ransac", C.c_int32), ("compat_stale_features", C.c_int32),
        ("cam0_model", C.c_int32), ("cam1_model", C.c_int32),
        ("cam0_intrinsics", C.c_double * 4), ("cam0_distortion", C.c_double * 4),
        ("cam1_intrinsics", C.c_double * 4), ("cam1_distortion", C.c_double * 4),
        ("T_cam0_imu", C.c_double * 16), ("T_cn_cnm1", C.c_double * 16), ("T_imu_body", C.c_double * 16),
        ("frame_rate", C.c_double),
        ("max_cam_state_size", C.c_int32), ("chi2_mode", C.c_int32),
        ("position_std_threshold", C.c_double),
        ("rotation_threshold", C.c_double), ("translation_threshold", C.c_double),
        ("tracking_rate_threshold", C.c_double), ("feature_translation_threshold", C.c_double),
        ("noise_gyro", C.c_double), ("noise_acc", C.c_double), ("noise_gyro_bias", C.c_double),
        ("noise_acc_bias", C.c_double), ("noise_feature", C.c_double),
        ("initial_velocity", C.c_double * 3),
        ("cov_velocity", C.c_double), ("cov_gyro_bias", C.c_double), ("cov_acc_bias", C.c_double),
        ("cov_ext_rot", C.c_double), ("cov_ext_trans", C.c_double),
        ("max_jacobian_rows", C.c_int32), ("fix_prev_image_alias", C.c_int32),
    ]


class Feature(C.Structure):
    _fields_ = [("id", C.c_uint32), ("pad", C.c_uint32), ("u0", C.c_double), ("v0", C.c_double),
                ("u1", C.c_double), ("v1", C.c_double)]


class TrackingInfo(C.Structure):
    _fields_ = [("time_stamp", C.c_double), ("before_tracking", C.c_int32), ("after_tracking", C.c_int32),
                ("after_matching", C.c_int32), ("after_ransac", C.c_int32)]


class GridFeature(C.Structure):
    _fields_ = [("id", C.c_uint64), ("response", C.c_float), ("lifetime", C.c_int32),
                ("cam0_x", C.c_float), ("cam0_y", C.c_float), ("cam1_x", C.c_float), ("cam1_y", C.c_float),
                ("cell", C.c_int32), ("pad", C.c_int32)]


class State(C.Structure):
    _fields_ = [("time", C.c_double), ("id", C.c_int64), ("orientation", C.c_double * 4),
                ("position", C.c_double * 3), ("velocity", C.c_double * 3), ("gyro_bias", C.c_double * 3),
                ("acc_bias", C.c_double * 3), ("R_imu_cam0", C.c_double * 9), ("t_cam0_imu", C.c_double * 3),
                ("gravity", C.c_double * 3), ("n_cam_states", C.c_int32), ("cov_dim", C.c_int32),
                ("is_gravity_set", C.c_int32), ("n_map_features", C.c_int32), ("tracking_rate", C.c_double),
                ("T_b_w", C.c_double * 16), ("n_updates", C.c_int64), ("n_resets", C.c_int64)]


class CamState(C.Structure):
    _fields_ = [("id", C.c_int64), ("time", C.c_double), ("orientation", C.c_double * 4),
                ("position", C.c_double * 3)]
