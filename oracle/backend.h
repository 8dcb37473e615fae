// ORACLE — TEST INFRASTRUCTURE ONLY (see linalg.h).
//
// backend.h: statement-by-statement CPU restatement of MsckfVio
// (msckf_core/src/msckf_vio.cpp, include/msckf_vio.h), Feature (include/feature.hpp) and the
// state structs (include/common/{imu_state,cam_state}.h).  Each method cites the lines it
// follows.  Class statics of the reference (IMUState::gravity, next_id, noise, T_imu_body,
// CAMState::T_cam0_cam1, Feature::observation_noise; msckf_vio.cpp:33-47) are instance
// members here so that several oracles can live in one process.
// PARITY UNPINNED for svd_fulluv / SPQR / Eigen LDLT (absent libraries): see linalg.h.
#pragma once
#include <array>
#include <map>

#include "chi2_table.h"
#include "frontend.h"

namespace orc {

typedef long long StateIDType;
typedef long long FeatureIDType;

struct IMUState {  // imu_state.h:28-88
    StateIDType id = 0;
    double time = 0;
    Quat orientation;
    V3 position, velocity, gyro_bias, acc_bias;
    M3 R_imu_cam0 = M3::eye();
    V3 t_cam0_imu;
    Quat orientation_null;
    V3 position_null, velocity_null;
};
struct CAMState {  // cam_state.h:25-60
    StateIDType id = 0;
    double time = 0;
    Quat orientation;
    V3 position;
    Quat orientation_null;
    V3 position_null;
};
typedef std::map<StateIDType, CAMState> CamStateServer;

struct Obs4 { double v[4]; };

struct OptimizationConfig {  // feature.hpp:38-55
    double translation_threshold = 0.2, huber_epsilon = 0.01, estimation_precision = 5e-7, initial_damping = 1e-3;
    int outer_loop_max_iteration = 10, inner_loop_max_iteration = 10;
};

struct Feature {  // feature.hpp:31-163
    FeatureIDType id = 0;
    std::map<StateIDType, Obs4> observations;
    V3 position;
    bool is_initialized = false;

    // feature.hpp:171-190
    static void cost(const SE3 &T_c0_ci, const V3 &x, const double z[2], double &e) {
        V3 h = T_c0_ci.R * V3(x[0], x[1], 1.0) + x[2] * T_c0_ci.t;
        double zx = h[0] / h[2] - z[0], zy = h[1] / h[2] - z[1];
        e = zx * zx + zy * zy;
    }
    // feature.hpp:192-229
    static void jacobian(const SE3 &T_c0_ci, const V3 &x, const double z[2], double J[6], double r[2], double &w,
                         const OptimizationConfig &oc) {
        V3 h = T_c0_ci.R * V3(x[0], x[1], 1.0) + x[2] * T_c0_ci.t;
        double h1 = h[0], h2 = h[1], h3 = h[2];
        double W[9];
        for (int i = 0; i < 3; ++i) {
            W[i * 3 + 0] = T_c0_ci.R(i, 0);
            W[i * 3 + 1] = T_c0_ci.R(i, 1);
            W[i * 3 + 2] = T_c0_ci.t[i];
        }
        for (int j = 0; j < 3; ++j) {
            J[j] = 1 / h3 * W[j] - h1 / (h3 * h3) * W[6 + j];
            J[3 + j] = 1 / h3 * W[3 + j] - h2 / (h3 * h3) * W[6 + j];
        }
        r[0] = h1 / h3 - z[0];
        r[1] = h2 / h3 - z[1];
        double e = std::sqrt(r[0] * r[0] + r[1] * r[1]);
        if (e <= oc.huber_epsilon) w = 1.0;
        else w = std::sqrt(2.0 * oc.huber_epsilon / e);
    }
    // feature.hpp:231-255
    static void generateInitialGuess(const SE3 &T_c1_c2, const double z1[2], const double z2[2], V3 &p) {
        V3 m = T_c1_c2.R * V3(z1[0], z1[1], 1.0);
        double A[2] = {m[0] - z2[0] * m[2], m[1] - z2[1] * m[2]};
        double b[2] = {z2[0] * T_c1_c2.t[2] - T_c1_c2.t[0], z2[1] * T_c1_c2.t[2] - T_c1_c2.t[1]};
        double depth = (1.0 / (A[0] * A[0] + A[1] * A[1])) * (A[0] * b[0] + A[1] * b[1]);
        p = V3(z1[0] * depth, z1[1] * depth, depth);
    }
    // feature.hpp:257-287
    bool checkMotion(const CamStateServer &cam_states, const OptimizationConfig &oc) const {
        StateIDType first_cam_id = observations.begin()->first;
        StateIDType last_cam_id = (--observations.end())->first;
        const CAMState &c0 = cam_states.find(first_cam_id)->second, &c1 = cam_states.find(last_cam_id)->second;
        SE3 first_pose(quat_to_rot(c0.orientation).t(), c0.position);
        SE3 last_pose(quat_to_rot(c1.orientation).t(), c1.position);
        V3 dir(observations.begin()->second.v[0], observations.begin()->second.v[1], 1.0);
        dir = dir / dir.norm();
        dir = first_pose.R * dir;
        V3 translation = last_pose.t - first_pose.t;
        double parallel = translation.dot(dir);
        V3 orth = translation - parallel * dir;
        return orth.norm() > oc.translation_threshold;
    }
    // feature.hpp:289-450
    bool initializePosition(const CamStateServer &cam_states, const SE3 &T_cam0_cam1, const OptimizationConfig &oc) {
        std::vector<SE3> cam_poses;
        std::vector<std::array<double, 2>> measurements;
        for (auto &m : observations) {
            auto it = cam_states.find(m.first);
            if (it == cam_states.end()) continue;
            measurements.push_back({m.second.v[0], m.second.v[1]});
            measurements.push_back({m.second.v[2], m.second.v[3]});
            SE3 cam0_pose(quat_to_rot(it->second.orientation).t(), it->second.position);
            SE3 cam1_pose = cam0_pose * T_cam0_cam1.inv();
            cam_poses.push_back(cam0_pose);
            cam_poses.push_back(cam1_pose);
        }
        SE3 T_c0_w = cam_poses[0];
        for (auto &pose : cam_poses) pose = pose.inv() * T_c0_w;
        V3 initial_position;
        generateInitialGuess(cam_poses.back(), measurements[0].data(), measurements.back().data(), initial_position);
        V3 solution(initial_position[0] / initial_position[2], initial_position[1] / initial_position[2],
                    1.0 / initial_position[2]);
        double lambda = oc.initial_damping;
        int inner_loop_cntr = 0, outer_loop_cntr = 0;
        bool is_cost_reduced = false;
        double delta_norm = 0;
        double total_cost = 0.0;
        for (size_t i = 0; i < cam_poses.size(); ++i) {
            double c = 0;
            cost(cam_poses[i], solution, measurements[i].data(), c);
            total_cost += c;
        }
        do {
            M3 A;
            V3 b;
            for (size_t i = 0; i < cam_poses.size(); ++i) {
                double J[6], r[2], w;
                jacobian(cam_poses[i], solution, measurements[i].data(), J, r, w, oc);
                double ws = (w == 1) ? 1.0 : w * w;
                for (int a = 0; a < 3; ++a) {
                    for (int c = 0; c < 3; ++c) {
                        double jtj = J[a] * J[c] + J[3 + a] * J[3 + c];
                        A(a, c) += (w == 1) ? jtj : ws * jtj;
                    }
                    double jtr = J[a] * r[0] + J[3 + a] * r[1];
                    b[a] += (w == 1) ? jtr : ws * jtr;
                }
            }
            do {
                M3 At = A;
                for (int i = 0; i < 3; ++i) At(i, i) += lambda;
                Mat X = ldlt_solve(toMat(At), toMat(b));  // Eigen Matrix3d::ldlt().solve, feature.hpp:395
                V3 delta(X.d[0], X.d[1], X.d[2]);
                V3 new_solution = solution - delta;
                delta_norm = delta.norm();
                double new_cost = 0.0;
                for (size_t i = 0; i < cam_poses.size(); ++i) {
                    double c = 0;
                    cost(cam_poses[i], new_solution, measurements[i].data(), c);
                    new_cost += c;
                }
                if (new_cost < total_cost) {
                    is_cost_reduced = true;
                    solution = new_solution;
                    total_cost = new_cost;
                    lambda = lambda / 10 > 1e-10 ? lambda / 10 : 1e-10;
                } else {
                    is_cost_reduced = false;
                    lambda = lambda * 10 < 1e12 ? lambda * 10 : 1e12;
                }
            } while (inner_loop_cntr++ < oc.inner_loop_max_iteration && !is_cost_reduced);
            inner_loop_cntr = 0;
        } while (outer_loop_cntr++ < oc.outer_loop_max_iteration && delta_norm > oc.estimation_precision);
        V3 final_position(solution[0] / solution[2], solution[1] / solution[2], 1.0 / solution[2]);
        bool is_valid_solution = true;
        for (const auto &pose : cam_poses) {
            V3 p = pose.R * final_position + pose.t;
            if (p[2] <= 0) {
                is_valid_solution = false;
                break;
            }
        }
        position = T_c0_w.R * final_position + T_c0_w.t;
        if (is_valid_solution) is_initialized = true;
        return is_valid_solution;
    }
};
typedef std::map<FeatureIDType, Feature> MapServer;

class MsckfVio {
public:
    explicit MsckfVio(const mskf_config &c) : cfg(c) {
        loadParameters();
        initialize();
    }

    // msckf_vio.cpp:58-162
    void loadParameters() {
        opt_cfg.translation_threshold = cfg.feature_translation_threshold;
        gyro_noise = cfg.noise_gyro * cfg.noise_gyro;
        acc_noise = cfg.noise_acc * cfg.noise_acc;
        gyro_bias_noise = cfg.noise_gyro_bias * cfg.noise_gyro_bias;
        acc_bias_noise = cfg.noise_acc_bias * cfg.noise_acc_bias;
        observation_noise = cfg.noise_feature * cfg.noise_feature;
        imu_state.velocity = V3(cfg.initial_velocity[0], cfg.initial_velocity[1], cfg.initial_velocity[2]);
        resetCovariance();
        SE3 T_cam0_imu = SE3::from16(cfg.T_cam0_imu).inv();
        imu_state.R_imu_cam0 = T_cam0_imu.R.t();
        imu_state.t_cam0_imu = T_cam0_imu.t;
        T_cam0_cam1 = SE3::from16(cfg.T_cn_cnm1);
        T_imu_body = SE3::from16(cfg.T_imu_body).inv();
        max_cam_state_size = cfg.max_cam_state_size;
    }
    void resetCovariance() {  // msckf_vio.cpp:102-112 (= :275-285, :1222-1232)
        state_cov = Mat(21, 21);
        for (int i = 3; i < 6; ++i) state_cov(i, i) = cfg.cov_gyro_bias;
        for (int i = 6; i < 9; ++i) state_cov(i, i) = cfg.cov_velocity;
        for (int i = 9; i < 12; ++i) state_cov(i, i) = cfg.cov_acc_bias;
        for (int i = 15; i < 18; ++i) state_cov(i, i) = cfg.cov_ext_rot;
        for (int i = 18; i < 21; ++i) state_cov(i, i) = cfg.cov_ext_trans;
    }
    // msckf_vio.cpp:164-188
    void initialize() {
        continuous_noise_cov = Mat(12, 12);
        for (int i = 0; i < 3; ++i) {
            continuous_noise_cov(i, i) = gyro_noise;
            continuous_noise_cov(3 + i, 3 + i) = gyro_bias_noise;
            continuous_noise_cov(6 + i, 6 + i) = acc_noise;
            continuous_noise_cov(9 + i, 9 + i) = acc_bias_noise;
        }
        gravity = V3(0, 0, -9.81);  // imu_state.h:22, msckf_vio.cpp:38
    }
    double chi2(int dof) const {  // msckf_vio.cpp:181-185; std::map operator[] yields 0 off-table
        if (dof < 1 || dof > 99) return 0.0;
        return cfg.chi2_mode == MSKF_CHI2_Q95 ? kChi2Q95[dof - 1] : kChi2Q05[dof - 1];
    }

    // msckf_vio.cpp:190-207
    void imuCallback(const ImuMsg &m) {
        imu_msg_buffer.push_back(m);
        if (!is_gravity_set) {
            if (imu_msg_buffer.size() < 200) return;
            initializeGravityAndBias();
            is_gravity_set = true;
        }
    }
    // msckf_vio.cpp:209-241
    void initializeGravityAndBias() {
        V3 sum_w, sum_a;
        for (const auto &m : imu_msg_buffer) {
            sum_w = sum_w + m.w;
            sum_a = sum_a + m.a;
        }
        imu_state.gyro_bias = sum_w / (double)imu_msg_buffer.size();
        V3 gravity_imu = sum_a / (double)imu_msg_buffer.size();
        double gravity_norm = gravity_imu.norm();
        gravity = V3(0.0, 0.0, -gravity_norm);
        imu_state.orientation = rot_to_quat(from_two_vector(gravity_imu, -gravity).t());
    }
    // msckf_vio.cpp:243-304
    void resetCallback() {
        imu_state.time = 0.0;
        imu_state.orientation = Quat();
        imu_state.position = V3();
        imu_state.velocity = V3();
        imu_state.gyro_bias = V3();
        imu_state.acc_bias = V3();
        imu_state.orientation_null = Quat();
        imu_state.position_null = V3();
        imu_state.velocity_null = V3();
        cam_states.clear();
        resetCovariance();
        map_server.clear();
        imu_msg_buffer.clear();
        is_gravity_set = false;
        is_first_img = true;
    }

    // msckf_vio.cpp:306-375
    void featureCallback(const CameraMeasurement &msg) {
        if (!is_gravity_set) return;
        if (is_first_img) {
            is_first_img = false;
            imu_state.time = msg.time_stamp;
        }
        batchImuProcessing(msg.time_stamp);
        stateAugmentation(msg.time_stamp);
        addFeatureObservations(msg);
        removeLostFeatures();
        pruneCamStateBuffer();
        publish(msg.time_stamp);
        n_pub++;
        onlineReset();
    }

    // msckf_vio.cpp:377-407
    void batchImuProcessing(double time_bound) {
        int used = 0;
        for (const auto &m : imu_msg_buffer) {
            if (m.t < imu_state.time) {
                ++used;
                continue;
            }
            if (m.t > time_bound) break;
            processModel(m.t, m.w, m.a);
            ++used;
        }
        imu_state.id = next_state_id++;
        imu_msg_buffer.erase(imu_msg_buffer.begin(), imu_msg_buffer.begin() + used);
    }

    // msckf_vio.cpp:409-480
    void processModel(double time, const V3 &m_gyro, const V3 &m_acc) {
        V3 gyro = m_gyro - imu_state.gyro_bias;
        V3 acc = m_acc - imu_state.acc_bias;
        double dtime = time - imu_state.time;
        Mat F(21, 21), G(21, 12);
        M3 Rt = quat_to_rot(imu_state.orientation).t();
        F.set(0, 0, toMat(-skew(gyro)));
        F.set(0, 3, toMat(-M3::eye()));
        F.set(6, 0, toMat(-(Rt * skew(acc))));
        F.set(6, 9, toMat(-Rt));
        F.set(12, 6, toMat(M3::eye()));
        G.set(0, 0, toMat(-M3::eye()));
        G.set(3, 3, toMat(M3::eye()));
        G.set(6, 6, toMat(-Rt));
        G.set(9, 9, toMat(M3::eye()));
        Mat Fdt = F * dtime;
        Mat Fdt_square = Fdt * Fdt;
        Mat Fdt_cube = Fdt_square * Fdt;
        Mat Phi = Mat::eye(21) + Fdt + 0.5 * Fdt_square + (1.0 / 6.0) * Fdt_cube;
        predictNewState(dtime, gyro, acc);
        M3 R_kk_1 = quat_to_rot(imu_state.orientation_null);
        Phi.set(0, 0, toMat(quat_to_rot(imu_state.orientation) * R_kk_1.t()));
        V3 u = R_kk_1 * gravity;
        V3 s = (1.0 / u.dot(u)) * u;
        auto block3 = [&](int r0, int c0) {
            M3 m;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) m(i, j) = Phi(r0 + i, c0 + j);
            return m;
        };
        M3 A1 = block3(6, 0);
        V3 w1 = skew(imu_state.velocity_null - imu_state.velocity) * gravity;
        Phi.set(6, 0, toMat(A1 - outer(A1 * u - w1, s)));
        M3 A2 = block3(12, 0);
        V3 w2 = skew(dtime * imu_state.velocity_null + imu_state.position_null - imu_state.position) * gravity;
        Phi.set(12, 0, toMat(A2 - outer(A2 * u - w2, s)));
        Mat Q = Phi * G * continuous_noise_cov * G.t() * Phi.t() * dtime;
        state_cov.set(0, 0, Phi * state_cov.block(0, 0, 21, 21) * Phi.t() + Q);
        if (cam_states.size() > 0) {
            state_cov.set(0, 21, Phi * state_cov.block(0, 21, 21, state_cov.c - 21));
            state_cov.set(21, 0, state_cov.block(21, 0, state_cov.r - 21, 21) * Phi.t());
        }
        state_cov = (state_cov + state_cov.t()) * 0.5;
        imu_state.orientation_null = imu_state.orientation;
        imu_state.position_null = imu_state.position;
        imu_state.velocity_null = imu_state.velocity;
        imu_state.time = time;
    }

    // msckf_vio.cpp:482-531
    void predictNewState(double dt, const V3 &gyro, const V3 &acc) {
        double gyro_norm = gyro.norm();
        double Om[16] = {0};
        M3 ms = -skew(gyro);
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) Om[i * 4 + j] = ms(i, j);
            Om[i * 4 + 3] = gyro[i];
            Om[12 + i] = -gyro[i];
        }
        Quat &q = imu_state.orientation;
        V3 &v = imu_state.velocity;
        V3 &p = imu_state.position;
        double q4[4] = {q.x, q.y, q.z, q.w};
        auto apply = [&](double cI, double cO, double post) {
            double r[4];
            for (int i = 0; i < 4; ++i) {
                double s = 0;
                for (int j = 0; j < 4; ++j) s += ((i == j ? cI : 0.0) + cO * Om[i * 4 + j]) * post * q4[j];
                r[i] = s;
            }
            return Quat(r[0], r[1], r[2], r[3]);
        };
        Quat dq_dt, dq_dt2;
        if (gyro_norm > 1e-5) {
            dq_dt = apply(std::cos(gyro_norm * dt * 0.5), 1 / gyro_norm * std::sin(gyro_norm * dt * 0.5), 1.0);
            dq_dt2 = apply(std::cos(gyro_norm * dt * 0.25), 1 / gyro_norm * std::sin(gyro_norm * dt * 0.25), 1.0);
        } else {
            dq_dt = apply(1.0, 0.5 * dt, std::cos(gyro_norm * dt * 0.5));
            dq_dt2 = apply(1.0, 0.25 * dt, std::cos(gyro_norm * dt * 0.25));
        }
        M3 dR_dt_transpose = quat_to_rot(dq_dt).t();
        M3 dR_dt2_transpose = quat_to_rot(dq_dt2).t();
        V3 k1_v_dot = quat_to_rot(q).t() * acc + gravity;
        V3 k1_p_dot = v;
        V3 k1_v = v + k1_v_dot * dt / 2;
        V3 k2_v_dot = dR_dt2_transpose * acc + gravity;
        V3 k2_p_dot = k1_v;
        V3 k2_v = v + k2_v_dot * dt / 2;
        V3 k3_v_dot = dR_dt2_transpose * acc + gravity;
        V3 k3_p_dot = k2_v;
        V3 k3_v = v + k3_v_dot * dt;
        V3 k4_v_dot = dR_dt_transpose * acc + gravity;
        V3 k4_p_dot = k3_v;
        q = dq_dt.normalized();
        v = v + dt / 6 * (k1_v_dot + 2 * k2_v_dot + 2 * k3_v_dot + k4_v_dot);
        p = p + dt / 6 * (k1_p_dot + 2 * k2_p_dot + 2 * k3_p_dot + k4_p_dot);
    }

    // msckf_vio.cpp:533-585
    void stateAugmentation(double time) {
        const M3 &R_i_c = imu_state.R_imu_cam0;
        const V3 &t_c_i = imu_state.t_cam0_imu;
        M3 R_w_i = quat_to_rot(imu_state.orientation);
        M3 R_w_c = R_i_c * R_w_i;
        V3 t_c_w = imu_state.position + R_w_i.t() * t_c_i;
        CAMState &cam_state = cam_states[imu_state.id];
        cam_state = CAMState();
        cam_state.id = imu_state.id;
        cam_state.time = time;
        cam_state.orientation = rot_to_quat(R_w_c);
        cam_state.position = t_c_w;
        cam_state.orientation_null = cam_state.orientation;
        cam_state.position_null = cam_state.position;
        Mat J(6, 21);
        J.set(0, 0, toMat(R_i_c));
        J.set(0, 15, toMat(M3::eye()));
        J.set(3, 0, toMat(skew(R_w_i.t() * t_c_i)));
        J.set(3, 12, toMat(M3::eye()));
        J.set(3, 18, toMat(M3::eye()));
        int old_rows = state_cov.r, old_cols = state_cov.c;
        state_cov.conservative_resize(old_rows + 6, old_cols + 6);
        Mat P11 = state_cov.block(0, 0, 21, 21);
        Mat P12 = state_cov.block(0, 21, 21, old_cols - 21);
        state_cov.set(old_rows, 0, J * P11);
        state_cov.set(old_rows, 21, J * P12);
        state_cov.set(0, old_cols, state_cov.block(old_rows, 0, 6, old_cols).t());
        state_cov.set(old_rows, old_cols, J * P11 * J.t());
        state_cov = (state_cov + state_cov.t()) * 0.5;
    }

    // msckf_vio.cpp:587-608
    void addFeatureObservations(const CameraMeasurement &msg) {
        StateIDType state_id = imu_state.id;
        int curr_feature_num = (int)map_server.size();
        int tracked_feature_num = 0;
        for (const auto &f : msg.features) {
            FeatureIDType fid = (FeatureIDType)f.id;
            Obs4 o = {{f.u0, f.v0, f.u1, f.v1}};
            if (map_server.find(fid) == map_server.end()) {
                map_server[fid] = Feature();
                map_server[fid].id = fid;
                map_server[fid].observations[state_id] = o;
            } else {
                map_server[fid].observations[state_id] = o;
                ++tracked_feature_num;
            }
        }
        tracking_rate = (double)tracked_feature_num / (double)curr_feature_num;
    }

    // msckf_vio.cpp:610-677
    void measurementJacobian(StateIDType cam_state_id, FeatureIDType feature_id, Mat &H_x, Mat &H_f, double r[4]) {
        const CAMState &cam_state = cam_states[cam_state_id];
        const Feature &feature = map_server[feature_id];
        M3 R_w_c0 = quat_to_rot(cam_state.orientation);
        const V3 &t_c0_w = cam_state.position;
        M3 R_c0_c1 = T_cam0_cam1.R;
        M3 R_w_c1 = T_cam0_cam1.R * R_w_c0;
        V3 t_c1_w = t_c0_w - R_w_c1.t() * T_cam0_cam1.t;
        const V3 &p_w = feature.position;
        const Obs4 &z = feature.observations.find(cam_state_id)->second;
        V3 p_c0 = R_w_c0 * (p_w - t_c0_w);
        V3 p_c1 = R_w_c1 * (p_w - t_c1_w);
        Mat dz_dpc0(4, 3), dz_dpc1(4, 3);
        dz_dpc0(0, 0) = 1 / p_c0[2];
        dz_dpc0(1, 1) = 1 / p_c0[2];
        dz_dpc0(0, 2) = -p_c0[0] / (p_c0[2] * p_c0[2]);
        dz_dpc0(1, 2) = -p_c0[1] / (p_c0[2] * p_c0[2]);
        dz_dpc1(2, 0) = 1 / p_c1[2];
        dz_dpc1(3, 1) = 1 / p_c1[2];
        dz_dpc1(2, 2) = -p_c1[0] / (p_c1[2] * p_c1[2]);
        dz_dpc1(3, 2) = -p_c1[1] / (p_c1[2] * p_c1[2]);
        Mat dpc0_dxc(3, 6), dpc1_dxc(3, 6);
        dpc0_dxc.set(0, 0, toMat(skew(p_c0)));
        dpc0_dxc.set(0, 3, toMat(-R_w_c0));
        dpc1_dxc.set(0, 0, toMat(R_c0_c1 * skew(p_c0)));
        dpc1_dxc.set(0, 3, toMat(-R_w_c1));
        H_x = dz_dpc0 * dpc0_dxc + dz_dpc1 * dpc1_dxc;
        H_f = dz_dpc0 * toMat(R_w_c0) + dz_dpc1 * toMat(R_w_c1);
        Mat A = H_x;
        Mat u(6, 1);
        u.set(0, 0, toMat(quat_to_rot(cam_state.orientation_null) * gravity));
        u.set(3, 0, toMat(skew(p_w - cam_state.position_null) * gravity));
        double utu = 0;
        for (int i = 0; i < 6; ++i) utu += u(i, 0) * u(i, 0);
        H_x = A - A * u * (1.0 / utu) * u.t();
        H_f = -H_x.block(0, 3, 4, 3);
        r[0] = z.v[0] - p_c0[0] / p_c0[2];
        r[1] = z.v[1] - p_c0[1] / p_c0[2];
        r[2] = z.v[2] - p_c1[0] / p_c1[2];
        r[3] = z.v[3] - p_c1[1] / p_c1[2];
    }

    // msckf_vio.cpp:679-775
    void featureJacobian(FeatureIDType feature_id, const std::vector<StateIDType> &cam_state_ids, Mat &H_x, Mat &r) {
        const Feature &feature = map_server[feature_id];
        std::vector<StateIDType> valid_cam_state_ids;
        for (const auto &cam_id : cam_state_ids) {
            if (feature.observations.find(cam_id) == feature.observations.end()) continue;
            valid_cam_state_ids.push_back(cam_id);
        }
        int jacobian_row_size = 4 * (int)valid_cam_state_ids.size();
        Mat H_xj(jacobian_row_size, 21 + (int)cam_states.size() * 6);
        Mat H_fj(jacobian_row_size, 3);
        Mat r_j(jacobian_row_size, 1);
        int stack_cntr = 0;
        for (const auto &cam_id : valid_cam_state_ids) {
            Mat H_xi(4, 6), H_fi(4, 3);
            double r_i[4];
            measurementJacobian(cam_id, feature.id, H_xi, H_fi, r_i);
            int cam_state_cntr = (int)std::distance(cam_states.begin(), cam_states.find(cam_id));
            H_xj.set(stack_cntr, 21 + 6 * cam_state_cntr, H_xi);
            H_fj.set(stack_cntr, 0, H_fi);
            for (int i = 0; i < 4; ++i) r_j(stack_cntr + i, 0) = r_i[i];
            stack_cntr += 4;
        }
        // :757-763 A = last 4M-3 columns of U from svd_fulluv(H_fj): an orthonormal basis of
        // the left null space of H_fj.  Householder QR gives an equivalent basis (linalg.h).
        Mat QR = H_fj;
        std::vector<double> tau;
        householder_qr(QR, tau);
        Mat Hq = H_xj, rq = r_j;
        apply_qt(QR, tau, Hq);
        apply_qt(QR, tau, rq);
        H_x = Hq.block(3, 0, jacobian_row_size - 3, Hq.c);
        r = rq.block(3, 0, jacobian_row_size - 3, 1);
    }

    // msckf_vio.cpp:795-857,897-904: the linear algebra of measurementUpdate on (H, r, P)
    static void update_math(const Mat &H, const Mat &r, const Mat &P, double obs_noise, Mat &delta_x, Mat &P_new) {
        Mat H_thin, r_thin;
        if (H.r > H.c) {
            // :795-810 SPQR (natural ordering): H_thin = (Q^T H)[0:n], r_thin = (Q^T r)[0:n]
            Mat QR = H;
            std::vector<double> tau;
            householder_qr(QR, tau);
            Mat rq = r;
            apply_qt(QR, tau, rq);
            int n = H.c;
            H_thin = Mat(n, H.c);
            for (int i = 0; i < n; ++i)
                for (int j = i; j < H.c; ++j) H_thin(i, j) = QR(i, j);
            r_thin = rq.block(0, 0, n, 1);
        } else {
            H_thin = H;
            r_thin = r;
        }
        Mat HP = H_thin * P;
        Mat S = HP * H_thin.t();
        for (int i = 0; i < S.r; ++i) S(i, i) += obs_noise;
        Mat K_transpose = ldlt_solve(S, HP);  // :850
        Mat K = K_transpose.t();
        delta_x = K * r_thin;
        Mat I_KH = Mat::eye(K.r) - K * H_thin;
        P_new = I_KH * P;
        P_new = (P_new + P_new.t()) * 0.5;
    }

    // msckf_vio.cpp:778-907
    void measurementUpdate(const Mat &H, const Mat &r) {
        if (H.r == 0 || r.r == 0) return;
        ++n_updates;
        if (keep_last_update) {
            last_H = H;
            last_r = r;
            last_P_prior = state_cov;
            last_cam_ids.clear();
            for (const auto &kv : cam_states) last_cam_ids.push_back((long long)kv.first);
        }
        Mat delta_x, P_new;
        update_math(H, r, state_cov, observation_noise, delta_x, P_new);
        last_delta_x = delta_x;
        auto seg3 = [&](int o) { return V3(delta_x(o, 0), delta_x(o + 1, 0), delta_x(o + 2, 0)); };
        const Quat dq_imu = small_angle_quat(seg3(0));
        imu_state.orientation = quat_mul(dq_imu, imu_state.orientation);
        imu_state.gyro_bias = imu_state.gyro_bias + seg3(3);
        imu_state.velocity = imu_state.velocity + seg3(6);
        imu_state.acc_bias = imu_state.acc_bias + seg3(9);
        imu_state.position = imu_state.position + seg3(12);
        const Quat dq_extrinsic = small_angle_quat(seg3(15));
        imu_state.R_imu_cam0 = quat_to_rot(dq_extrinsic) * imu_state.R_imu_cam0;
        imu_state.t_cam0_imu = imu_state.t_cam0_imu + seg3(18);
        int i = 0;
        for (auto it = cam_states.begin(); it != cam_states.end(); ++it, ++i) {
            const Quat dq_cam = small_angle_quat(seg3(21 + i * 6));
            it->second.orientation = quat_mul(dq_cam, it->second.orientation);
            it->second.position = it->second.position + seg3(21 + i * 6 + 3);
        }
        state_cov = P_new;
    }

    // msckf_vio.cpp:909-935
    bool gatingTest(const Mat &H, const Mat &r, int dof) {
        Mat P = H * state_cov * H.t();
        for (int i = 0; i < P.r; ++i) P(i, i) += observation_noise;
        Mat x = ldlt_solve(P, r);
        double gamma = 0;
        for (int i = 0; i < r.r; ++i) gamma += r(i, 0) * x(i, 0);
        last_gamma = gamma;
        return gamma < chi2(dof);
    }

    // msckf_vio.cpp:937-1024
    void removeLostFeatures() {
        int jacobian_row_size = 0;
        std::vector<FeatureIDType> invalid_feature_ids, processed_feature_ids;
        for (auto iter = map_server.begin(); iter != map_server.end(); ++iter) {
            auto &feature = iter->second;
            if (feature.observations.find(imu_state.id) != feature.observations.end()) continue;
            if (feature.observations.size() < 3) {
                invalid_feature_ids.push_back(feature.id);
                continue;
            }
            if (!feature.is_initialized) {
                if (!feature.checkMotion(cam_states, opt_cfg)) {
                    invalid_feature_ids.push_back(feature.id);
                    continue;
                } else if (!feature.initializePosition(cam_states, T_cam0_cam1, opt_cfg)) {
                    invalid_feature_ids.push_back(feature.id);
                    continue;
                }
            }
            jacobian_row_size += 4 * (int)feature.observations.size() - 3;
            processed_feature_ids.push_back(feature.id);
        }
        for (const auto &fid : invalid_feature_ids) map_server.erase(fid);
        if (processed_feature_ids.empty()) return;
        Mat H_x(jacobian_row_size, 21 + 6 * (int)cam_states.size());
        Mat r(jacobian_row_size, 1);
        int stack_cntr = 0;
        for (const auto &fid : processed_feature_ids) {
            auto &feature = map_server[fid];
            std::vector<StateIDType> cam_state_ids;
            for (const auto &m : feature.observations) cam_state_ids.push_back(m.first);
            Mat H_xj, r_j;
            featureJacobian(feature.id, cam_state_ids, H_xj, r_j);
            if (gatingTest(H_xj, r_j, (int)cam_state_ids.size() - 1)) {
                H_x.set(stack_cntr, 0, H_xj);
                r.set(stack_cntr, 0, r_j);
                stack_cntr += H_xj.r;
            }
            if (stack_cntr > cfg.max_jacobian_rows) break;
        }
        H_x.conservative_resize(stack_cntr, H_x.c);
        r.conservative_resize(stack_cntr, 1);
        measurementUpdate(H_x, r);
        for (const auto &fid : processed_feature_ids) map_server.erase(fid);
    }

    // msckf_vio.cpp:1026-1071
    void findRedundantCamStates(std::vector<StateIDType> &rm_cam_state_ids) {
        auto key_cam_state_iter = cam_states.end();
        for (int i = 0; i < 4; ++i) --key_cam_state_iter;
        auto cam_state_iter = key_cam_state_iter;
        ++cam_state_iter;
        auto first_cam_state_iter = cam_states.begin();
        const V3 key_position = key_cam_state_iter->second.position;
        const M3 key_rotation = quat_to_rot(key_cam_state_iter->second.orientation);
        for (int i = 0; i < 2; ++i) {
            const V3 position = cam_state_iter->second.position;
            const M3 rotation = quat_to_rot(cam_state_iter->second.orientation);
            double distance = (position - key_position).norm();
            double angle = rotation_angle(rotation * key_rotation.t());
            if (angle < cfg.rotation_threshold && distance < cfg.translation_threshold &&
                tracking_rate > cfg.tracking_rate_threshold) {
                rm_cam_state_ids.push_back(cam_state_iter->first);
                ++cam_state_iter;
            } else {
                rm_cam_state_ids.push_back(first_cam_state_iter->first);
                ++first_cam_state_iter;
            }
        }
        std::sort(rm_cam_state_ids.begin(), rm_cam_state_ids.end());
    }

    // msckf_vio.cpp:1073-1184
    void pruneCamStateBuffer() {
        if ((int)cam_states.size() < max_cam_state_size) return;
        std::vector<StateIDType> rm_cam_state_ids;
        findRedundantCamStates(rm_cam_state_ids);
        int jacobian_row_size = 0;
        for (auto &item : map_server) {
            auto &feature = item.second;
            std::vector<StateIDType> involved;
            for (const auto &cam_id : rm_cam_state_ids)
                if (feature.observations.find(cam_id) != feature.observations.end()) involved.push_back(cam_id);
            if (involved.size() == 0) continue;
            if (involved.size() == 1) {
                feature.observations.erase(involved[0]);
                continue;
            }
            if (!feature.is_initialized) {
                if (!feature.checkMotion(cam_states, opt_cfg)) {
                    for (const auto &cam_id : involved) feature.observations.erase(cam_id);
                    continue;
                } else if (!feature.initializePosition(cam_states, T_cam0_cam1, opt_cfg)) {
                    for (const auto &cam_id : involved) feature.observations.erase(cam_id);
                    continue;
                }
            }
            jacobian_row_size += 4 * (int)involved.size() - 3;
        }
        Mat H_x(jacobian_row_size, 21 + 6 * (int)cam_states.size());
        Mat r(jacobian_row_size, 1);
        int stack_cntr = 0;
        for (auto &item : map_server) {
            auto &feature = item.second;
            std::vector<StateIDType> involved;
            for (const auto &cam_id : rm_cam_state_ids)
                if (feature.observations.find(cam_id) != feature.observations.end()) involved.push_back(cam_id);
            if (involved.size() == 0) continue;
            Mat H_xj, r_j;
            featureJacobian(feature.id, involved, H_xj, r_j);
            if (gatingTest(H_xj, r_j, (int)involved.size())) {
                H_x.set(stack_cntr, 0, H_xj);
                r.set(stack_cntr, 0, r_j);
                stack_cntr += H_xj.r;
            }
            for (const auto &cam_id : involved) feature.observations.erase(cam_id);
        }
        H_x.conservative_resize(stack_cntr, H_x.c);
        r.conservative_resize(stack_cntr, 1);
        measurementUpdate(H_x, r);
        for (const auto &cam_id : rm_cam_state_ids) {
            int cam_sequence = (int)std::distance(cam_states.begin(), cam_states.find(cam_id));
            int cam_state_start = 21 + 6 * cam_sequence;
            int cam_state_end = cam_state_start + 6;
            if (cam_state_end < state_cov.r) {
                state_cov.set(cam_state_start, 0,
                              state_cov.block(cam_state_end, 0, state_cov.r - cam_state_end, state_cov.c));
                state_cov.set(0, cam_state_start,
                              state_cov.block(0, cam_state_end, state_cov.r, state_cov.c - cam_state_end));
            }
            state_cov.conservative_resize(state_cov.r - 6, state_cov.c - 6);
            cam_states.erase(cam_id);
        }
    }

    // msckf_vio.cpp:1186-1236
    void onlineReset() {
        if (cfg.position_std_threshold <= 0) return;
        double sx = std::sqrt(state_cov(12, 12)), sy = std::sqrt(state_cov(13, 13)), sz = std::sqrt(state_cov(14, 14));
        if (sx < cfg.position_std_threshold && sy < cfg.position_std_threshold && sz < cfg.position_std_threshold)
            return;
        ++n_resets;
        cam_states.clear();
        map_server.clear();
        resetCovariance();
    }

    // msckf_vio.cpp:1238-1305 (pose part; the unbounded points3d_ list is not reproduced)
    void publish(double) {
        SE3 T_i_w(quat_to_rot(imu_state.orientation).t(), imu_state.position);
        T_b_w = T_imu_body * T_i_w * T_imu_body.inv();
    }

    mskf_config cfg;
    IMUState imu_state;
    CamStateServer cam_states;
    Mat state_cov, continuous_noise_cov;
    MapServer map_server;
    std::vector<ImuMsg> imu_msg_buffer;
    OptimizationConfig opt_cfg;
    double gyro_noise, acc_noise, gyro_bias_noise, acc_bias_noise, observation_noise;
    V3 gravity;
    SE3 T_cam0_cam1, T_imu_body, T_b_w;
    StateIDType next_state_id = 0;
    bool is_gravity_set = false, is_first_img = true;
    int max_cam_state_size = 20;
    double tracking_rate = 0;
    long long n_pub = 0, n_updates = 0, n_resets = 0;
    Mat last_delta_x;
    double last_gamma = 0;
    bool keep_last_update = false;  // test hook: keep (H, r, P-) of the latest measurementUpdate
    Mat last_H, last_r, last_P_prior;
    std::vector<long long> last_cam_ids;  // camera-state ids behind the column groups of last_H
};

}  // namespace orc
