// msckf_b200.hpp — C++ host façade over the C ABI of msckf_b200.h.
//
// Mirrors the reference's host-side API name for name so an application written against
// mfkiwl/msckf_stereo_c switches by changing the include and the constructor arguments:
//   cg::Image / Imu / FeatureMeasurement / CameraMeasurement / TrackingInfo   include/common/data_msg.h:15-55
//   cg::ImageProcessor {initialize, stereoCallback, imuCallback, feature_msg_ptr_}   include/image_processor.h:30-58
//   cg::MsckfVio {initialize, resetCallback, imuCallback, featureCallback, get_path}  include/msckf_vio.h:39-80
//   cg::System {stereo_callback, imu_callback, backend_callback, path_to_draw_}       include/system.h:18-31
// Differences, all forced by the absent vikit_cg / yaml-cpp types: cg::YImg8 becomes
// {data, rows, cols, stride}; cg::Vector3 becomes double[3]; YAML::Node becomes mskf_config
// (mskf_default_config fills the values of the three reference YAML files).  One Engine owns
// `n_streams` independent filters on one GPU; ImageProcessor / MsckfVio / System are views of
// one stream of it, so a fleet is `n_streams` System objects sharing an Engine and one
// Engine::step() per frame.  Header-only; link with libmsckf_b200.so.  No CPU fallback:
// every failure of the C ABI (including "no CUDA device") throws std::runtime_error.
#pragma once
#include <array>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "msckf_b200.h"

namespace mskf_b200 {

struct Image {  // cg::Image
    double time_stamp = 0;
    const uint8_t *data = nullptr;  // cg::YImg8: rows x cols, 8-bit luminance
    int rows = 0, cols = 0, stride = 0;
};
struct Imu {  // cg::Imu
    double time_stamp = 0;
    double angular_velocity[3] = {0, 0, 0};
    double linear_acceleration[3] = {0, 0, 0};
};
typedef std::shared_ptr<const Imu> ImuConstPtr;
struct FeatureMeasurement {  // cg::FeatureMeasurement
    unsigned int id = 0;
    double u0 = 0, v0 = 0, u1 = 0, v1 = 0;
};
struct CameraMeasurement {  // cg::CameraMeasurement
    double time_stamp = 0;
    std::vector<FeatureMeasurement> features;
};
typedef std::shared_ptr<CameraMeasurement> CameraMeasurementPtr;
typedef std::shared_ptr<const CameraMeasurement> CameraMeasurementConstPtr;
typedef mskf_tracking_info TrackingInfo;  // cg::TrackingInfo
typedef std::array<double, 16> Mat4;      // row-major T_b_w

inline mskf_config default_config(const std::string &preset = "ref") {
    mskf_config c;
    if (mskf_default_config(&c, preset.c_str()) != MSKF_OK) throw std::invalid_argument("unknown preset " + preset);
    return c;
}

class Engine {
public:
    Engine(const mskf_config &cfg, int n_streams = 1, int device = 0) : n_streams_(n_streams) {
        int rc = mskf_create(&cfg, n_streams, device, &h_);
        if (rc != MSKF_OK) {
            std::string msg = h_ ? mskf_last_error(h_) : "no CUDA device";
            if (h_) mskf_destroy(h_);
            h_ = nullptr;
            throw std::runtime_error("mskf_create failed (" + std::to_string(rc) + "): " + msg);
        }
    }
    ~Engine() {
        if (h_) mskf_destroy(h_);
    }
    Engine(const Engine &) = delete;
    Engine &operator=(const Engine &) = delete;
    mskf_handle *handle() const { return h_; }
    int n_streams() const { return n_streams_; }
    void check(int rc) const {
        if (rc != MSKF_OK) throw std::runtime_error(std::string("msckf_b200: ") + mskf_last_error(h_));
    }
    // stereoCallback + featureCallback of every stream with a staged pair
    void frontend_step() { check(mskf_frontend_step(h_)); }
    void backend_step() { check(mskf_backend_step(h_)); }
    void step() { check(mskf_step(h_)); }
    void sync() { check(mskf_sync(h_)); }

private:
    mskf_handle *h_ = nullptr;
    int n_streams_;
};
typedef std::shared_ptr<Engine> EnginePtr;

class ImageProcessor {  // cg::ImageProcessor, image_processor.h:30-58
public:
    ImageProcessor(EnginePtr e, int stream = 0) : feature_msg_ptr_(new CameraMeasurement), e_(e), s_(stream) {}
    bool initialize() { return true; }  // parameters were loaded by mskf_create
    // Stages the pair and, like the reference, runs the front end before returning.  In a fleet use
    // stageStereo() on every stream and one Engine::frontend_step().
    void stereoCallback(const Image &cam0_img, const Image &cam1_img, bool /*is_draw*/ = false) {
        stageStereo(cam0_img, cam1_img);
        e_->frontend_step();
        fetchFeatures();
    }
    void stageStereo(const Image &cam0_img, const Image &cam1_img) {
        e_->check(mskf_push_stereo(e_->handle(), s_, cam0_img.time_stamp, cam0_img.data, cam1_img.data, cam0_img.rows,
                                   cam0_img.cols, cam0_img.stride ? cam0_img.stride : cam0_img.cols));
    }
    // image_processor.cpp:205-211: the front end's own IMU buffer only (System::imu_callback feeds both halves)
    void imuCallback(const ImuConstPtr &msg) {
        e_->check(mskf_push_imu_to(e_->handle(), s_, MSKF_IMU_FRONTEND, msg->time_stamp, msg->angular_velocity,
                                   msg->linear_acceleration));
    }
    // Copies the device-side message into feature_msg_ptr_.  Like the reference's vector it is never cleared
    // (SURVEY F4): it grows to the reference's length, but only the head that can hold measurements is fetched
    // and rewritten each frame - the tail beyond it is value-initialised once, when the vector grows.
    void fetchFeatures() {
        int n_head = 0;
        long long n_total = 0;
        double t = 0;
        e_->check(mskf_get_features_head(e_->handle(), s_, nullptr, 0, &n_head, &n_total, &t));
        buf_.resize(n_head);
        if (n_head) e_->check(mskf_get_features_head(e_->handle(), s_, buf_.data(), n_head, &n_head, &n_total, &t));
        feature_msg_ptr_->time_stamp = t;
        std::vector<FeatureMeasurement> &v = feature_msg_ptr_->features;
        if ((long long)v.size() != n_total) v.resize((size_t)n_total);  // shrinks only in fixed mode (fresh message per frame)
        for (int i = 0; i < n_head; ++i) {
            FeatureMeasurement &f = v[i];
            f.id = buf_[i].id; f.u0 = buf_[i].u0; f.v0 = buf_[i].v0; f.u1 = buf_[i].u1; f.v1 = buf_[i].v1;
        }
    }
    TrackingInfo trackingInfo() const {
        TrackingInfo ti;
        e_->check(mskf_get_tracking_info(e_->handle(), s_, &ti));
        return ti;
    }
    std::shared_ptr<CameraMeasurement> feature_msg_ptr_;

private:
    EnginePtr e_;
    int s_;
    std::vector<mskf_feature> buf_;
};

class MsckfVio {  // cg::MsckfVio, msckf_vio.h:39-80
public:
    MsckfVio(EnginePtr e, int stream = 0) : e_(e), s_(stream) {}
    bool initialize() { return true; }
    bool resetCallback() {
        e_->check(mskf_reset(e_->handle(), s_));
        return true;
    }
    // msckf_vio.cpp:190-207: the filter's IMU buffer; the 200th sample initialises gravity and the gyro bias.
    // A caller that drives MsckfVio alone (imuCallback + featureCallback(msg)) needs nothing else.
    void imuCallback(const ImuConstPtr &msg) {
        e_->check(mskf_push_imu_to(e_->handle(), s_, MSKF_IMU_BACKEND, msg->time_stamp, msg->angular_velocity,
                                   msg->linear_acceleration));
    }
    // The device-resident message of the front end is consumed directly ...
    void featureCallback() {
        e_->backend_step();
        publish();
    }
    // ... or a caller-supplied one (the reference signature)
    void featureCallback(const CameraMeasurementConstPtr &msg) {
        std::vector<mskf_feature> buf(msg->features.size());
        for (size_t i = 0; i < buf.size(); ++i) {
            const FeatureMeasurement &f = msg->features[i];
            buf[i].id = f.id; buf[i].pad = 0; buf[i].u0 = f.u0; buf[i].v0 = f.v0; buf[i].u1 = f.u1; buf[i].v1 = f.v1;
        }
        e_->check(mskf_backend_step_features(e_->handle(), s_, msg->time_stamp, buf.data(), (int)buf.size()));
        publish();
    }
    mskf_state state() const {
        mskf_state st;
        e_->check(mskf_get_state(e_->handle(), s_, &st));
        return st;
    }
    std::vector<double> covariance() const {
        int dim = 0;
        e_->check(mskf_get_covariance(e_->handle(), s_, nullptr, 0, &dim));
        std::vector<double> P((size_t)dim * dim);
        e_->check(mskf_get_covariance(e_->handle(), s_, P.data(), (int)P.size(), &dim));
        return P;
    }
    const std::vector<std::pair<double, Mat4>> &get_path() const { return path_; }  // (time, T_b_w), msckf_vio.cpp:1262-1266

private:
    void publish() {
        mskf_state st = state();
        if (!st.is_gravity_set) return;
        Mat4 T;
        for (int i = 0; i < 16; ++i) T[i] = st.T_b_w[i];
        path_.push_back(std::make_pair(st.time, T));
    }
    EnginePtr e_;
    int s_;
    std::vector<std::pair<double, Mat4>> path_;
};

class System {  // cg::System, system.h:18-31, system.cpp:12-54
public:
    explicit System(const mskf_config &cfg, int device = 0) : System(std::make_shared<Engine>(cfg, 1, device), 0) {}
    System(EnginePtr e, int stream)
        : imgproc_ptr_(new ImageProcessor(e, stream)), msckfvio_ptr_(new MsckfVio(e, stream)),
          feature_msg_ptr_(imgproc_ptr_->feature_msg_ptr_), e_(e) {
        imgproc_ptr_->initialize();
        msckfvio_ptr_->initialize();
    }
    void stereo_callback(const Image &cam0_img, const Image &cam1_img, bool is_draw = false) {
        imgproc_ptr_->stereoCallback(cam0_img, cam1_img, is_draw);
        feature_msg_ptr_ = imgproc_ptr_->feature_msg_ptr_;
    }
    void imu_callback(const ImuConstPtr &msg) {
        imgproc_ptr_->imuCallback(msg);
        msckfvio_ptr_->imuCallback(msg);
    }
    void backend_callback() {
        msckfvio_ptr_->featureCallback();
        // system.cpp:52 copies the whole path every frame (O(frames^2) over a run); only the new poses are appended here
        const std::vector<std::pair<double, Mat4>> &p = msckfvio_ptr_->get_path();
        if (p.size() < path_to_draw_.size()) path_to_draw_.clear();
        path_to_draw_.insert(path_to_draw_.end(), p.begin() + path_to_draw_.size(), p.end());
    }
    std::shared_ptr<ImageProcessor> imgproc_ptr_;
    std::shared_ptr<MsckfVio> msckfvio_ptr_;
    std::shared_ptr<CameraMeasurement> feature_msg_ptr_;
    std::vector<std::pair<double, Mat4>> path_to_draw_;

private:
    EnginePtr e_;
};

}  // namespace mskf_b200
