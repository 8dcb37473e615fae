// run_stream.cpp — the reference's run_euroc_single_thread feed loop
// (apps/run_euroc_single_thread.cpp:189-254) against the C++ façade, on a raw dump of a stereo+IMU
// stream:  run_stream <dump> [preset] [vio]
//   with `vio` the printed filter is a second, stand-alone MsckfVio (its own engine) driven the way the
//   reference allows: MsckfVio::imuCallback + featureCallback(msg) with the front end's CameraMeasurement
//   dump = int32 n_frames, rows, cols; then per frame: int32 n_imu, n_imu x {t, w[3], a[3]} doubles,
//          double t_img, rows*cols bytes cam0, rows*cols bytes cam1   (written by tests/test_cpp_facade.py)
// Prints one TUM line per frame (time tx ty tz qx qy qz qw, msckf_vio.cpp:1255-1258) at full precision.
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "../include/msckf_b200.hpp"

using namespace mskf_b200;

int main(int argc, char **argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s dump [preset]\n", argv[0]);
        return 2;
    }
    FILE *f = std::fopen(argv[1], "rb");
    if (!f) return 2;
    int hdr[3];
    if (std::fread(hdr, 4, 3, f) != 3) return 2;
    const int n_frames = hdr[0], rows = hdr[1], cols = hdr[2];
    mskf_config cfg = default_config(argc > 2 ? argv[2] : "ref");
    cfg.img_rows = rows;
    cfg.img_cols = cols;
    const bool vio_alone = argc > 3 && std::string(argv[3]) == "vio";
    try {
        System sys(cfg, 0);
        std::shared_ptr<MsckfVio> vio;
        if (vio_alone) {
            vio.reset(new MsckfVio(std::make_shared<Engine>(cfg, 1, 0), 0));
            vio->initialize();
        }
        std::vector<uint8_t> im0((size_t)rows * cols), im1((size_t)rows * cols);
        for (int k = 0; k < n_frames; ++k) {
            int n_imu = 0;
            if (std::fread(&n_imu, 4, 1, f) != 1) return 2;
            for (int i = 0; i < n_imu; ++i) {
                double v[7];
                if (std::fread(v, 8, 7, f) != 7) return 2;
                std::shared_ptr<Imu> m(new Imu);
                m->time_stamp = v[0];
                for (int j = 0; j < 3; ++j) {
                    m->angular_velocity[j] = v[1 + j];
                    m->linear_acceleration[j] = v[4 + j];
                }
                sys.imu_callback(m);
                if (vio) vio->imuCallback(m);
            }
            Image a, b;
            if (std::fread(&a.time_stamp, 8, 1, f) != 1) return 2;
            if (std::fread(im0.data(), 1, im0.size(), f) != im0.size() || std::fread(im1.data(), 1, im1.size(), f) != im1.size()) return 2;
            b.time_stamp = a.time_stamp;
            a.data = im0.data(); b.data = im1.data();
            a.rows = b.rows = rows; a.cols = b.cols = cols; a.stride = b.stride = cols;
            sys.stereo_callback(a, b);
            sys.backend_callback();
            if (vio) vio->featureCallback(sys.feature_msg_ptr_);
            mskf_state st = vio ? vio->state() : sys.msckfvio_ptr_->state();
            if (sys.path_to_draw_.size() != sys.msckfvio_ptr_->get_path().size()) return 3;
            std::printf("%.9f %.17g %.17g %.17g %.17g %.17g %.17g %.17g %d %zu\n", a.time_stamp, st.position[0], st.position[1],
                        st.position[2], st.orientation[0], st.orientation[1], st.orientation[2], st.orientation[3], st.n_cam_states,
                        sys.feature_msg_ptr_->features.size());
        }
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    std::fclose(f);
    return 0;
}
