// run_euroc.cpp — headless re-host of apps/run_euroc_single_thread.cpp:117-254 on the B200 engine:
//   run_euroc <mav0 dir> [preset=ref] [pose_out.txt] [decimals=6]
// Feed order as the reference: IMU rows until one has t > t_img, stereo_callback, backend_callback.
// Writes the TUM trajectory of MsckfVio::publish (msckf_vio.cpp:1255-1258; std::fixed => 6 decimals in
// the reference, which loses the sub-microsecond part of the stamp: `decimals` can raise it).
//   run_euroc --dump-image <file> <out.raw>   decodes one image (PNG/PGM) for the decoder test.
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../include/msckf_b200.hpp"
#include "euroc_io.hpp"

using namespace mskf_b200;

int main(int argc, char **argv) {
    if (argc >= 4 && std::string(argv[1]) == "--dump-image") {
        euroc::Gray8 g = euroc::load_gray8(argv[2]);
        if (g.empty()) return 1;
        FILE *o = std::fopen(argv[3], "wb");
        std::fwrite(g.data.data(), 1, g.data.size(), o);
        std::fclose(o);
        std::printf("%d %d\n", g.rows, g.cols);
        return 0;
    }
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <mav0 dir> [preset] [pose_out.txt] [decimals]\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1], preset = argc > 2 ? argv[2] : "ref", out_path = argc > 3 ? argv[3] : "pose_out.txt";
    const int decimals = argc > 4 ? std::atoi(argv[4]) : 6;
    try {
        std::vector<euroc::Stamped> cams[2] = {euroc::read_cam_csv(dir + "/cam0/data.csv"), euroc::read_cam_csv(dir + "/cam1/data.csv")};
        if (cams[0].empty() || cams[0].size() != cams[1].size()) throw std::runtime_error("cam0/cam1 csv mismatch");
        std::ifstream imu_file(dir + "/imu0/data.csv");
        if (!imu_file.good()) throw std::runtime_error("no imu file found");
        std::string line;
        std::getline(imu_file, line);
        euroc::Gray8 first = euroc::load_gray8(dir + "/cam0/data/" + cams[0][0].name);
        if (first.empty()) throw std::runtime_error("cannot read " + cams[0][0].name);
        mskf_config cfg = default_config(preset);
        cfg.img_rows = first.rows;
        cfg.img_cols = first.cols;
        System sys(cfg, 0);
        FILE *out = std::fopen(out_path.c_str(), "w");
        if (!out) throw std::runtime_error("cannot write " + out_path);
        double t_decode = 0, t_path = 0;
        size_t n_frames = 0;
        for (size_t k = 0; k < cams[0].size(); ++k) {
            auto c0 = std::chrono::steady_clock::now();
            euroc::Gray8 im[2];
            for (int j = 0; j < 2; ++j) im[j] = euroc::load_gray8(dir + "/cam" + std::to_string(j) + "/data/" + cams[j][k].name);
            if (im[0].empty() || im[1].empty()) {
                std::fprintf(stderr, "ERROR: img is empty !!!\n");
                continue;
            }
            auto c1 = std::chrono::steady_clock::now();
            const double t_img = cams[0][k].t;
            double t_imu = 0.0;
            do {
                if (!std::getline(imu_file, line)) break;
                euroc::ImuRow row;
                if (!euroc::parse_imu_row(line, row)) continue;
                std::shared_ptr<Imu> m(new Imu);
                m->time_stamp = row.t;
                for (int j = 0; j < 3; ++j) {
                    m->angular_velocity[j] = row.w[j];
                    m->linear_acceleration[j] = row.a[j];
                }
                sys.imu_callback(m);
                t_imu = row.t;
            } while (t_imu <= t_img);
            Image a, b;
            a.time_stamp = b.time_stamp = t_img;
            a.data = im[0].data.data(); b.data = im[1].data.data();
            a.rows = b.rows = im[0].rows; a.cols = b.cols = im[0].cols; a.stride = b.stride = im[0].cols;
            sys.stereo_callback(a, b, false);
            sys.backend_callback();
            mskf_state st = sys.msckfvio_ptr_->state();
            if (st.is_gravity_set) {
                double R[9], q[4];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) R[i * 3 + j] = st.T_b_w[i * 4 + j];
                euroc::rot_to_quat_hamilton(R, q);
                std::fprintf(out, "%.*f %.*f %.*f %.*f %.*f %.*f %.*f %.*f\n", decimals, t_img, decimals, st.T_b_w[3], decimals, st.T_b_w[7],
                             decimals, st.T_b_w[11], decimals, q[0], decimals, q[1], decimals, q[2], decimals, q[3]);
            }
            auto c2 = std::chrono::steady_clock::now();
            t_decode += std::chrono::duration<double>(c1 - c0).count();
            t_path += std::chrono::duration<double>(c2 - c1).count();
            ++n_frames;
        }
        std::fclose(out);
        std::fprintf(stderr, "%zu frames: image decode %.1f ms/frame (excluded from the path), callbacks %.2f ms/frame\n", n_frames,
                     1e3 * t_decode / n_frames, 1e3 * t_path / n_frames);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
