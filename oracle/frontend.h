// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg.h).
//
// frontend.h: statement-by-statement CPU restatement of ImageProcessor
// (msckf_core/src/image_processor.cpp, include/image_processor.h).  Each method cites the
// lines it follows.  Known reference defects are mirrored, not fixed:
//   F3 twoPointRansac is never called (image_processor.cpp:482-493) -> off unless use_ransac
//   F4 publish() never clears the message (image_processor.cpp:1157-1164) -> compat_stale_features
//   next_feature_id is never initialised (image_processor.h:314) -> treated as 0
//   response index mix-up in addNewFeatures (image_processor.cpp:668-677 vs :698) -> mirrored
//   F6 cam0_prev_img_ptr aliases cam0_curr_img_ptr (image_processor.cpp:192) -> dt = 0 in :881
//   std::sort is unstable -> SPEC fixes stable order (ties keep insertion order)
#pragma once
#include <array>
#include <map>
#include <memory>
#include <string>

#include "../include/msckf_b200.h"
#include "kin.h"
#include "spec_cv.h"

namespace orc {

struct ImuMsg {
    double t;
    V3 w, a;
};
struct FeatureMeasurement {  // data_msg.h:30-37
    unsigned int id = 0;
    double u0 = 0, v0 = 0, u1 = 0, v1 = 0;
};
struct CameraMeasurement {  // data_msg.h:40-43
    double time_stamp = 0;
    std::vector<FeatureMeasurement> features;
};

class ImageProcessor {
public:
    typedef unsigned long long FeatureIDType;
    struct FeatureMetaData {  // image_processor.h:75-81
        FeatureIDType id = 0;
        float response = 0;
        int lifetime = 0;
        Pt cam0_point, cam1_point;
    };
    typedef std::map<int, std::vector<FeatureMetaData>> GridFeatures;

    explicit ImageProcessor(const mskf_config &c) : cfg(c) {
        feature_msg.reset(new CameraMeasurement);
        prev_features.reset(new GridFeatures);
        curr_features.reset(new GridFeatures);
        loadParameters();
    }

    // image_processor.cpp:52-124
    void loadParameters() {
        SE3 m4_cam0_imu = SE3::from16(cfg.T_cam0_imu);
        R_cam0_imu = m4_cam0_imu.R.t();
        t_cam0_imu = -(R_cam0_imu * m4_cam0_imu.t);
        SE3 m4_cam1_cam0 = SE3::from16(cfg.T_cn_cnm1);
        SE3 T_cam1_imu = m4_cam1_cam0 * m4_cam0_imu;
        R_cam1_imu = T_cam1_imu.R.t();
        t_cam1_imu = -(R_cam1_imu * T_cam1_imu.t);
        // image_processor.cpp:132  detector_ = CornerDetector(30, 47, fast_threshold)
        detector.n_rows = cfg.det_rows;
        detector.n_cols = cfg.det_cols;
        detector.fast_threshold = cfg.fast_threshold;
        detector.detection_threshold = cfg.detection_threshold;
        detector.configure(cfg.img_rows, cfg.img_cols);
        klt.win = cfg.klt_win;
        klt.max_iters = cfg.klt_max_iters;
        klt.eps = cfg.klt_eps;
        klt.min_eig = cfg.klt_min_eig;
        grid_height = cfg.img_rows / cfg.grid_row;  // :250-251 (function-local statics)
        grid_width = cfg.img_cols / cfg.grid_col;
    }

    // image_processor.cpp:205-211
    void imuCallback(const ImuMsg &m) {
        if (is_first_img) return;
        imu_msg_buffer.push_back(m);
    }

    // image_processor.cpp:139-203
    void stereoCallback(double t, const uint8_t *cam0, const uint8_t *cam1) {
        curr_time = t;
        cam0_img = Img(cfg.img_rows, cfg.img_cols);
        cam1_img = Img(cfg.img_rows, cfg.img_cols);
        std::memcpy(cam0_img.d.data(), cam0, cam0_img.d.size());
        std::memcpy(cam1_img.d.data(), cam1, cam1_img.d.size());
        createImagePyramids();
        if (is_first_img) {
            initializeFirstFrame();
            is_first_img = false;
        } else {
            trackFeatures();
            addNewFeatures();
            pruneGridFeatures();
        }
        publish();
        prev_time = curr_time;
        prev_features = curr_features;
        std::swap(prev_cam0_pyramid, curr_cam0_pyramid);
        curr_features.reset(new GridFeatures);
        for (int code = 0; code < cfg.grid_row * cfg.grid_col; ++code) (*curr_features)[code] = {};
    }

    // image_processor.cpp:213-245 (4 levels hard-coded there; cfg.pyramid_levels here)
    void createImagePyramids() {
        curr_cam0_pyramid.clear();
        curr_cam1_pyramid.clear();
        for (int i = 0; i < cfg.pyramid_levels; ++i) {
            if (i == 0) {
                curr_cam0_pyramid.push_back(cam0_img);
                curr_cam1_pyramid.push_back(cam1_img);
                continue;
            }
            Img t1, t2;
            pyr_down(curr_cam0_pyramid[i - 1], t1);
            curr_cam0_pyramid.push_back(t1);
            pyr_down(curr_cam1_pyramid[i - 1], t2);
            curr_cam1_pyramid.push_back(t2);
        }
    }

    static void stable_by_response(std::vector<FeatureMetaData> &v) {
        std::stable_sort(v.begin(), v.end(),
                         [](const FeatureMetaData &a, const FeatureMetaData &b) { return a.response > b.response; });
    }

    // image_processor.cpp:247-319
    void initializeFirstFrame() {
        std::vector<Pt> new_features;
        std::vector<double> new_features_responses;
        detector.detect_features(cam0_img, new_features, new_features_responses);
        std::vector<Pt> cam0_points = new_features, cam1_points;
        std::vector<uint8_t> inlier_markers;
        stereoMatch(cam0_points, cam1_points, inlier_markers);
        std::vector<Pt> cam0_inliers, cam1_inliers;
        std::vector<float> response_inliers;
        for (size_t i = 0; i < inlier_markers.size(); ++i) {
            if (inlier_markers[i] == 0) continue;
            cam0_inliers.push_back(cam0_points[i]);
            cam1_inliers.push_back(cam1_points[i]);
            response_inliers.push_back((float)new_features_responses[i]);
        }
        GridFeatures grid_new_features;
        for (int code = 0; code < cfg.grid_row * cfg.grid_col; ++code) grid_new_features[code] = {};
        for (size_t i = 0; i < cam0_inliers.size(); ++i) {
            int row = (int)(cam0_inliers[i].y / grid_height);
            int col = (int)(cam0_inliers[i].x / grid_width);
            int code = row * cfg.grid_col + col;
            FeatureMetaData f;
            f.response = response_inliers[i];
            f.cam0_point = cam0_inliers[i];
            f.cam1_point = cam1_inliers[i];
            grid_new_features[code].push_back(f);
        }
        for (auto &item : grid_new_features) stable_by_response(item.second);
        for (int code = 0; code < cfg.grid_row * cfg.grid_col; ++code) {
            auto &features_this_grid = (*curr_features)[code];
            auto &new_features_this_grid = grid_new_features[code];
            for (int k = 0; k < cfg.grid_min_feature_num && k < (int)new_features_this_grid.size(); ++k) {
                features_this_grid.push_back(new_features_this_grid[k]);
                features_this_grid.back().id = next_feature_id++;
                features_this_grid.back().lifetime = 1;
            }
        }
    }

    // image_processor.cpp:321-350
    void predictFeatureTracking(const std::vector<Pt> &input_pts, const M3 &R_p_c, const double intr[4],
                                std::vector<Pt> &compensated_pts) {
        if (input_pts.empty()) {
            compensated_pts.clear();
            return;
        }
        compensated_pts.resize(input_pts.size());
        M3 K;
        K(0, 0) = intr[0]; K(0, 2) = intr[2];
        K(1, 1) = intr[1]; K(1, 2) = intr[3];
        K(2, 2) = 1.0;
        M3 H = K * R_p_c * inv3(K);
        for (size_t i = 0; i < input_pts.size(); ++i) {
            V3 p1((double)input_pts[i].x, (double)input_pts[i].y, 1.0);
            V3 p2 = H * p1;
            compensated_pts[i].x = (float)(p2[0] / p2[2]);
            compensated_pts[i].y = (float)(p2[1] / p2[2]);
        }
    }

    template <typename T>
    static void removeUnmarkedElements(const std::vector<T> &raw, const std::vector<uint8_t> &markers,
                                       std::vector<T> &refined) {  // image_processor.h:292-306
        for (size_t i = 0; i < markers.size(); ++i)
            if (markers[i]) refined.push_back(raw[i]);
    }

    // image_processor.cpp:352-532
    void trackFeatures() {
        M3 cam0_R_p_c, cam1_R_p_c;
        integrateImuData(cam0_R_p_c, cam1_R_p_c);
        last_cam0_R_p_c = cam0_R_p_c;
        std::vector<FeatureIDType> prev_ids;
        std::vector<int> prev_lifetime;
        std::vector<Pt> prev_cam0_points, prev_cam1_points;
        for (const auto &item : *prev_features)
            for (const auto &f : item.second) {
                prev_ids.push_back(f.id);
                prev_lifetime.push_back(f.lifetime);
                prev_cam0_points.push_back(f.cam0_point);
                prev_cam1_points.push_back(f.cam1_point);
            }
        before_tracking = (int)prev_cam0_points.size();
        if (prev_ids.empty()) return;
        std::vector<Pt> curr_cam0_points;
        std::vector<uint8_t> track_inliers;
        predictFeatureTracking(prev_cam0_points, cam0_R_p_c, cfg.cam0_intrinsics, curr_cam0_points);
        optical_flow_multi_level(prev_cam0_pyramid, curr_cam0_pyramid, prev_cam0_points, curr_cam0_points,
                                 track_inliers, klt);
        for (size_t i = 0; i < curr_cam0_points.size(); ++i) {
            if (track_inliers[i] == 0) continue;
            if (curr_cam0_points[i].y < 0 || curr_cam0_points[i].y > cfg.img_rows - 1 ||
                curr_cam0_points[i].x < 0 || curr_cam0_points[i].x > cfg.img_cols - 1)
                track_inliers[i] = 0;
        }
        std::vector<FeatureIDType> prev_tracked_ids;
        std::vector<int> prev_tracked_lifetime;
        std::vector<Pt> prev_tracked_cam0_points, prev_tracked_cam1_points, curr_tracked_cam0_points;
        removeUnmarkedElements(prev_ids, track_inliers, prev_tracked_ids);
        removeUnmarkedElements(prev_lifetime, track_inliers, prev_tracked_lifetime);
        removeUnmarkedElements(prev_cam0_points, track_inliers, prev_tracked_cam0_points);
        removeUnmarkedElements(prev_cam1_points, track_inliers, prev_tracked_cam1_points);
        removeUnmarkedElements(curr_cam0_points, track_inliers, curr_tracked_cam0_points);
        after_tracking = (int)curr_tracked_cam0_points.size();

        std::vector<Pt> curr_cam1_points;
        std::vector<uint8_t> match_inliers;
        stereoMatch(curr_tracked_cam0_points, curr_cam1_points, match_inliers);
        std::vector<FeatureIDType> prev_matched_ids;
        std::vector<int> prev_matched_lifetime;
        std::vector<Pt> prev_matched_cam0_points, prev_matched_cam1_points, curr_matched_cam0_points,
            curr_matched_cam1_points;
        removeUnmarkedElements(prev_tracked_ids, match_inliers, prev_matched_ids);
        removeUnmarkedElements(prev_tracked_lifetime, match_inliers, prev_matched_lifetime);
        removeUnmarkedElements(prev_tracked_cam0_points, match_inliers, prev_matched_cam0_points);
        removeUnmarkedElements(prev_tracked_cam1_points, match_inliers, prev_matched_cam1_points);
        removeUnmarkedElements(curr_tracked_cam0_points, match_inliers, curr_matched_cam0_points);
        removeUnmarkedElements(curr_cam1_points, match_inliers, curr_matched_cam1_points);
        after_matching = (int)curr_matched_cam0_points.size();

        // :482-493 the two twoPointRansac calls are commented out in the reference; with
        // use_ransac they run as the commented code reads
        std::vector<int> cam0_ransac_inliers, cam1_ransac_inliers;
        if (cfg.use_ransac) {
            twoPointRansac(prev_matched_cam0_points, curr_matched_cam0_points, cam0_R_p_c, cfg.cam0_intrinsics, cfg.cam0_model,
                           cfg.cam0_distortion, cfg.ransac_threshold, 0.99, cam0_ransac_inliers, 0);
            twoPointRansac(prev_matched_cam1_points, curr_matched_cam1_points, cam1_R_p_c, cfg.cam1_intrinsics, cfg.cam1_model,
                           cfg.cam1_distortion, cfg.ransac_threshold, 0.99, cam1_ransac_inliers, 1);
        }
        ++track_calls;
        after_ransac = 0;
        for (size_t i = 0; i < curr_matched_cam0_points.size(); ++i) {
            if (cfg.use_ransac && (cam0_ransac_inliers[i] == 0 || cam1_ransac_inliers[i] == 0)) continue;
            int row = (int)(curr_matched_cam0_points[i].y / grid_height);
            int col = (int)(curr_matched_cam0_points[i].x / grid_width);
            int code = row * cfg.grid_col + col;
            (*curr_features)[code].push_back(FeatureMetaData());
            FeatureMetaData &g = (*curr_features)[code].back();
            g.id = prev_matched_ids[i];
            g.lifetime = ++prev_matched_lifetime[i];
            g.cam0_point = curr_matched_cam0_points[i];
            g.cam1_point = curr_matched_cam1_points[i];
            ++after_ransac;
        }
    }

    // image_processor.cpp:534-620
    void stereoMatch(const std::vector<Pt> &cam0_points, std::vector<Pt> &cam1_points,
                     std::vector<uint8_t> &inlier_markers) {
        if (cam0_points.empty()) return;
        if (cam1_points.empty()) {
            const M3 R_cam0_cam1 = R_cam1_imu.t() * R_cam0_imu;
            std::vector<Pt> cam0_points_undistorted;
            undistortPoints(cam0_points, cfg.cam0_intrinsics, cfg.cam0_model, cfg.cam0_distortion,
                            cam0_points_undistorted, R_cam0_cam1);
            distort_points(cam0_points_undistorted, cam1_points, cfg.cam1_intrinsics, cfg.cam1_model,
                           cfg.cam1_distortion);
        }
        optical_flow_multi_level(curr_cam0_pyramid, curr_cam1_pyramid, cam0_points, cam1_points, inlier_markers,
                                 klt);
        for (size_t i = 0; i < cam1_points.size(); ++i) {
            if (inlier_markers[i] == 0) continue;
            if (cam1_points[i].y < 0 || cam1_points[i].y > cfg.img_rows - 1 || cam1_points[i].x < 0 ||
                cam1_points[i].x > cfg.img_cols - 1)
                inlier_markers[i] = 0;
        }
        const M3 R_cam0_cam1 = R_cam1_imu.t() * R_cam0_imu;
        const V3 t_cam0_cam1 = R_cam1_imu.t() * (t_cam0_imu - t_cam1_imu);
        const M3 E = skew(t_cam0_cam1) * R_cam0_cam1;
        std::vector<Pt> cam0_points_undistorted, cam1_points_undistorted;
        undistortPoints(cam0_points, cfg.cam0_intrinsics, cfg.cam0_model, cfg.cam0_distortion,
                        cam0_points_undistorted);
        undistortPoints(cam1_points, cfg.cam1_intrinsics, cfg.cam1_model, cfg.cam1_distortion,
                        cam1_points_undistorted);
        double norm_pixel_unit = 4.0 / (cfg.cam0_intrinsics[0] + cfg.cam0_intrinsics[1] + cfg.cam1_intrinsics[0] +
                                        cfg.cam1_intrinsics[1]);
        for (size_t i = 0; i < cam0_points_undistorted.size(); ++i) {
            if (inlier_markers[i] == 0) continue;
            V3 pt0((double)cam0_points_undistorted[i].x, (double)cam0_points_undistorted[i].y, 1.0);
            V3 pt1((double)cam1_points_undistorted[i].x, (double)cam1_points_undistorted[i].y, 1.0);
            V3 epipolar_line = E * pt0;
            double error = std::fabs(pt1.dot(epipolar_line)) /
                           std::sqrt(epipolar_line[0] * epipolar_line[0] + epipolar_line[1] * epipolar_line[1]);
            if (error > cfg.stereo_threshold * norm_pixel_unit) inlier_markers[i] = 0;
        }
    }

    // image_processor.cpp:622-756
    void addNewFeatures() {
        for (const auto &features : *curr_features)
            for (const auto &feature : features.second) {
                const int y = (int)feature.cam0_point.y;
                const int x = (int)feature.cam0_point.x;
                detector.set_grid_position(Pt((float)x, (float)y));
            }
        std::vector<Pt> new_features;
        std::vector<double> new_features_responses;
        detector.detect_features(cam0_img, new_features, new_features_responses);
        last_detected = new_features;
        std::vector<std::vector<std::pair<Pt, double>>> sieve(cfg.grid_row * cfg.grid_col);
        for (size_t i = 0; i < new_features.size(); ++i) {
            int row = (int)(new_features[i].y / grid_height);
            int col = (int)(new_features[i].x / grid_width);
            sieve[row * cfg.grid_col + col].push_back(std::make_pair(new_features[i], new_features_responses[i]));
        }
        new_features.clear();
        for (auto &item : sieve) {
            if ((int)item.size() > cfg.grid_max_feature_num) {
                std::stable_sort(item.begin(), item.end(),
                                 [](const std::pair<Pt, double> &a, const std::pair<Pt, double> &b) {
                                     return a.second > b.second;
                                 });
                item.erase(item.begin() + cfg.grid_max_feature_num, item.end());
            }
            for (auto &pt : item) new_features.push_back(pt.first);
        }
        std::vector<Pt> cam0_points = new_features, cam1_points;
        std::vector<uint8_t> inlier_markers;
        stereoMatch(cam0_points, cam1_points, inlier_markers);
        std::vector<Pt> cam0_inliers, cam1_inliers;
        std::vector<float> response_inliers;
        for (size_t i = 0; i < inlier_markers.size(); ++i) {
            if (inlier_markers[i] == 0) continue;
            cam0_inliers.push_back(cam0_points[i]);
            cam1_inliers.push_back(cam1_points[i]);
            // :698 indexes the pre-sieve response array with the post-sieve index (mirrored)
            response_inliers.push_back((float)new_features_responses[i]);
        }
        GridFeatures grid_new_features;
        for (int code = 0; code < cfg.grid_row * cfg.grid_col; ++code) grid_new_features[code] = {};
        for (size_t i = 0; i < cam0_inliers.size(); ++i) {
            int row = (int)(cam0_inliers[i].y / grid_height);
            int col = (int)(cam0_inliers[i].x / grid_width);
            int code = row * cfg.grid_col + col;
            FeatureMetaData f;
            f.response = response_inliers[i];
            f.cam0_point = cam0_inliers[i];
            f.cam1_point = cam1_inliers[i];
            grid_new_features[code].push_back(f);
        }
        for (auto &item : grid_new_features) stable_by_response(item.second);
        for (int code = 0; code < cfg.grid_row * cfg.grid_col; ++code) {
            auto &features_this_grid = (*curr_features)[code];
            auto &new_features_this_grid = grid_new_features[code];
            if ((int)features_this_grid.size() >= cfg.grid_min_feature_num) continue;
            int vacancy_num = cfg.grid_min_feature_num - (int)features_this_grid.size();
            for (int k = 0; k < vacancy_num && k < (int)new_features_this_grid.size(); ++k) {
                features_this_grid.push_back(new_features_this_grid[k]);
                features_this_grid.back().id = next_feature_id++;
                features_this_grid.back().lifetime = 1;
            }
        }
    }

    // cg::uniform_integer(lo, hi) lives in vikit_cg and is unseeded (image_processor.cpp:1026-1027).
    // SPEC: a counter-based hash of (trackFeatures call, camera, iteration, draw) so that the CPU
    // oracle and the CUDA engine draw the same pairs; only statistical parity with the reference.
    static uint32_t ransac_hash(uint32_t a, uint32_t b) {
        uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
        h ^= h >> 15; h *= 0x2C1B3C6Du;
        h ^= h >> 12; h *= 0x297A2D39u;
        h ^= h >> 15;
        return h;
    }
    static int uniform_integer(uint32_t call, int cam, int iter, int draw, int lo, int hi) {
        uint32_t h = ransac_hash(call * 2u + (uint32_t)cam, (uint32_t)(iter * 2 + draw));
        return lo + (int)(h % (uint32_t)(hi - lo + 1));
    }

    // image_processor.cpp:891-909
    static void rescalePoints(std::vector<Pt> &pts1, std::vector<Pt> &pts2, float &scaling_factor) {
        scaling_factor = 0.0f;
        for (size_t i = 0; i < pts1.size(); ++i) {
            scaling_factor += std::sqrt(pts1[i].x * pts1[i].x + pts1[i].y * pts1[i].y);
            scaling_factor += std::sqrt(pts2[i].x * pts2[i].x + pts2[i].y * pts2[i].y);
        }
        scaling_factor = (float)(pts1.size() + pts2.size()) / scaling_factor * std::sqrt(2.0f);
        for (size_t i = 0; i < pts1.size(); ++i) {
            pts1[i].x *= scaling_factor; pts1[i].y *= scaling_factor;
            pts2[i].x *= scaling_factor; pts2[i].y *= scaling_factor;
        }
    }

    // image_processor.cpp:911-1135
    void twoPointRansac(const std::vector<Pt> &pts1, const std::vector<Pt> &pts2, const M3 &R_p_c, const double intr[4], int model,
                        const double dist[4], double inlier_error, double success_probability, std::vector<int> &inlier_markers,
                        int cam) {
        double norm_pixel_unit = 2.0 / (intr[0] + intr[1]);
        int iter_num = (int)std::ceil(std::log(1 - success_probability) / std::log(1 - 0.7 * 0.7));
        inlier_markers.assign(pts1.size(), 1);
        std::vector<Pt> p1(pts1.size()), p2(pts2.size());
        undistortPoints(pts1, intr, model, dist, p1);
        undistortPoints(pts2, intr, model, dist, p2);
        for (auto &pt : p1) {
            V3 pt_hc = R_p_c * V3((double)pt.x, (double)pt.y, 1.0);
            pt.x = (float)pt_hc[0];
            pt.y = (float)pt_hc[1];
        }
        float scaling_factor = 0.0f;
        rescalePoints(p1, p2, scaling_factor);
        norm_pixel_unit *= (double)scaling_factor;
        const size_t n = p1.size();
        std::vector<Pt> diff(n);
        for (size_t i = 0; i < n; ++i) diff[i] = Pt(p1[i].x - p2[i].x, p1[i].y - p2[i].y);
        double mean_pt_distance = 0.0;
        int raw_inlier_cntr = 0;
        for (size_t i = 0; i < n; ++i) {
            double distance = (double)std::sqrt(diff[i].x * diff[i].x + diff[i].y * diff[i].y);
            if (distance > 50.0 * norm_pixel_unit) {
                inlier_markers[i] = 0;
            } else {
                mean_pt_distance += distance;
                ++raw_inlier_cntr;
            }
        }
        mean_pt_distance /= raw_inlier_cntr;
        if (raw_inlier_cntr < 3) {
            for (auto &m : inlier_markers) m = 0;
            return;
        }
        if (mean_pt_distance < norm_pixel_unit) {  // degenerate (no translation)
            for (size_t i = 0; i < n; ++i) {
                if (inlier_markers[i] == 0) continue;
                if ((double)std::sqrt(diff[i].x * diff[i].x + diff[i].y * diff[i].y) > inlier_error * norm_pixel_unit) inlier_markers[i] = 0;
            }
            return;
        }
        std::vector<std::array<double, 3>> coeff_t(n);
        for (size_t i = 0; i < n; ++i) {
            coeff_t[i][0] = (double)diff[i].y;
            coeff_t[i][1] = (double)(-diff[i].x);
            coeff_t[i][2] = (double)(p1[i].x * p2[i].y - p1[i].y * p2[i].x);
        }
        std::vector<int> raw_inlier_idx;
        for (size_t i = 0; i < n; ++i)
            if (inlier_markers[i] != 0) raw_inlier_idx.push_back((int)i);
        std::vector<int> best_inlier_set;
        const int nraw = (int)raw_inlier_idx.size();
        for (int iter_idx = 0; iter_idx < iter_num; ++iter_idx) {
            int select_idx1 = uniform_integer(track_calls, cam, iter_idx, 0, 0, nraw - 1);
            int select_idx_diff = uniform_integer(track_calls, cam, iter_idx, 1, 1, nraw - 1);
            int select_idx2 = select_idx1 + select_idx_diff < nraw ? select_idx1 + select_idx_diff : select_idx1 + select_idx_diff - nraw;
            int pair_idx1 = raw_inlier_idx[select_idx1], pair_idx2 = raw_inlier_idx[select_idx2];
            double c[3][2];
            for (int k = 0; k < 3; ++k) {
                c[k][0] = coeff_t[pair_idx1][k];
                c[k][1] = coeff_t[pair_idx2][k];
            }
            double l1[3];
            for (int k = 0; k < 3; ++k) l1[k] = std::fabs(c[k][0]) + std::fabs(c[k][1]);
            int base = 0;
            for (int k = 1; k < 3; ++k)
                if (l1[k] < l1[base]) base = k;
            const int ia = base == 0 ? 1 : 0, ib = base == 2 ? 1 : 2;  // the two non-base coefficient columns
            // A = [c[ia] c[ib]] (2x2, columns), solution = A^-1 (-c[base])
            double mdl[3];
            {
                double a00 = c[ia][0], a01 = c[ib][0], a10 = c[ia][1], a11 = c[ib][1];
                double det = a00 * a11 - a01 * a10, id = 1.0 / det;
                double i00 = a11 * id, i01 = -a01 * id, i10 = -a10 * id, i11 = a00 * id;
                double b0 = -c[base][0], b1 = -c[base][1];
                mdl[base] = 1.0;
                mdl[ia] = i00 * b0 + i01 * b1;
                mdl[ib] = i10 * b0 + i11 * b1;
            }
            std::vector<int> inlier_set;
            for (size_t i = 0; i < n; ++i) {
                if (inlier_markers[i] == 0) continue;
                double e = coeff_t[i][0] * mdl[0];
                e += coeff_t[i][1] * mdl[1];
                e += coeff_t[i][2] * mdl[2];
                if (std::fabs(e) < inlier_error * norm_pixel_unit) inlier_set.push_back((int)i);
            }
            if ((double)inlier_set.size() < 0.2 * (double)n) continue;
            // refit: solution = ((A^T A)^-1 A^T) (-b) with A = [coeff ia, coeff ib] over the inlier set
            double ata00 = 0, ata01 = 0, ata11 = 0;
            for (int idx : inlier_set) {
                ata00 += coeff_t[idx][ia] * coeff_t[idx][ia];
                ata01 += coeff_t[idx][ia] * coeff_t[idx][ib];
                ata11 += coeff_t[idx][ib] * coeff_t[idx][ib];
            }
            double det = ata00 * ata11 - ata01 * ata01, id = 1.0 / det;
            double m00 = ata11 * id, m01 = -ata01 * id, m11 = ata00 * id;
            double s0 = 0, s1 = 0;
            for (int idx : inlier_set) {
                double nb = -coeff_t[idx][base];
                s0 += (m00 * coeff_t[idx][ia] + m01 * coeff_t[idx][ib]) * nb;
                s1 += (m01 * coeff_t[idx][ia] + m11 * coeff_t[idx][ib]) * nb;
            }
            (void)s0;
            (void)s1;  // model_better only feeds this_error / best_error, which never influence the result (:1117-1121)
            if (inlier_set.size() > best_inlier_set.size()) best_inlier_set = inlier_set;
        }
        inlier_markers.assign(pts1.size(), 0);
        for (int idx : best_inlier_set) inlier_markers[idx] = 1;
    }

    // image_processor.cpp:758-768
    void pruneGridFeatures() {
        for (auto &item : *curr_features) {
            auto &g = item.second;
            if ((int)g.size() <= cfg.grid_max_feature_num) continue;
            std::stable_sort(g.begin(), g.end(), [](const FeatureMetaData &a, const FeatureMetaData &b) {
                return a.lifetime > b.lifetime;
            });
            g.erase(g.begin() + cfg.grid_max_feature_num, g.end());
        }
    }

    // image_processor.cpp:770-821 (rectification defaults to identity, new intrinsics to {1,1,0,0})
    void undistortPoints(const std::vector<Pt> &pts_in, const double intr[4], int model, const double dist[4],
                         std::vector<Pt> &pts_out, const M3 &rect = M3::eye()) {
        if (pts_in.empty()) return;
        const double Kn[4] = {1, 1, 0, 0};
        undistort_points(pts_in, pts_out, intr, model, dist, rect, Kn);
    }

    // image_processor.cpp:850-889
    void integrateImuData(M3 &cam0_R_p_c, M3 &cam1_R_p_c) {
        // :192 aliases the prev image object to the curr one, so prev time == curr time here
        const double prev_t = cfg.fix_prev_image_alias ? prev_time : curr_time;
        size_t begin = 0;
        while (begin < imu_msg_buffer.size()) {
            if (imu_msg_buffer[begin].t - prev_t < -0.01) ++begin;
            else break;
        }
        size_t end = begin;
        while (end < imu_msg_buffer.size()) {
            if (imu_msg_buffer[end].t - curr_time < 0.005) ++end;
            else break;
        }
        V3 mean_ang_vel;
        for (size_t i = begin; i < end; ++i) mean_ang_vel = mean_ang_vel + imu_msg_buffer[i].w;
        if (end > begin) mean_ang_vel = mean_ang_vel * (double)(1.0f / (float)(end - begin));
        V3 cam0_mean_ang_vel = R_cam0_imu.t() * mean_ang_vel;
        V3 cam1_mean_ang_vel = R_cam1_imu.t() * mean_ang_vel;
        double dtime = curr_time - prev_t;
        cam0_R_p_c = rodrigues(cam0_mean_ang_vel * dtime).t();
        cam1_R_p_c = rodrigues(cam1_mean_ang_vel * dtime).t();
        imu_msg_buffer.erase(imu_msg_buffer.begin(), imu_msg_buffer.begin() + end);
    }

    // image_processor.cpp:1137-1182
    void publish() {
        feature_msg->time_stamp = curr_time;
        std::vector<FeatureIDType> curr_ids;
        std::vector<Pt> curr_cam0_points, curr_cam1_points;
        for (const auto &g : *curr_features)
            for (const auto &f : g.second) {
                curr_ids.push_back(f.id);
                curr_cam0_points.push_back(f.cam0_point);
                curr_cam1_points.push_back(f.cam1_point);
            }
        std::vector<Pt> u0, u1;
        undistortPoints(curr_cam0_points, cfg.cam0_intrinsics, cfg.cam0_model, cfg.cam0_distortion, u0);
        undistortPoints(curr_cam1_points, cfg.cam1_intrinsics, cfg.cam1_model, cfg.cam1_distortion, u1);
        if (!cfg.compat_stale_features) feature_msg->features.clear();  // the F4 fix
        for (size_t i = 0; i < curr_ids.size(); ++i) {
            feature_msg->features.push_back(FeatureMeasurement());
            feature_msg->features[i].id = (unsigned int)curr_ids[i];
            feature_msg->features[i].u0 = u0[i].x;
            feature_msg->features[i].v0 = u0[i].y;
            feature_msg->features[i].u1 = u1[i].x;
            feature_msg->features[i].v1 = u1[i].y;
        }
        n_published = (int)curr_ids.size();
    }

    mskf_config cfg;
    std::shared_ptr<CameraMeasurement> feature_msg;
    bool is_first_img = true;
    FeatureIDType next_feature_id = 0;
    CornerDetector detector;
    KltParams klt;
    std::vector<ImuMsg> imu_msg_buffer;
    M3 R_cam0_imu, R_cam1_imu, last_cam0_R_p_c;
    V3 t_cam0_imu, t_cam1_imu;
    Img cam0_img, cam1_img;
    double prev_time = 0, curr_time = 0;
    std::vector<Img> prev_cam0_pyramid, curr_cam0_pyramid, curr_cam1_pyramid;
    std::shared_ptr<GridFeatures> prev_features, curr_features;
    int before_tracking = 0, after_tracking = 0, after_matching = 0, after_ransac = 0;
    int grid_height = 1, grid_width = 1;
    int n_published = 0;
    std::vector<Pt> last_detected;
    uint32_t track_calls = 0;  // trackFeatures calls so far (seeds the RANSAC sampler)
};

}  // namespace orc
