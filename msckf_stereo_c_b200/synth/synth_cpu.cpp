// synth_cpu.cpp — host build of the synthetic stereo+IMU generator (see scene.h).
// C API, loaded with ctypes by msckf_stereo_c_b200/synth.py.
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/msckf_b200_presets.h"
#include "scene.h"

struct synth_ctx {
    SynthTraj traj;
    SynthCam cam[2];
    std::vector<float> rays[2];
};

static void mat4_mul(const double *a, const double *b, double *c) {
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            c[i * 4 + j] = s;
        }
}

// pixel -> normalised undistorted direction (inverse of the radtan / equidistant model)
static void pixel_to_ray(double u, double v, const double K[4], int model, const double D[4], float *out) {
    double x = (u - K[2]) / K[0], y = (v - K[3]) / K[1];
    if (model == MSKF_MODEL_RADTAN) {
        double x0 = x, y0 = y;
        for (int it = 0; it < 20; ++it) {
            double r2 = x * x + y * y;
            double icd = 1.0 / (1.0 + (D[1] * r2 + D[0]) * r2);
            double dx = 2.0 * D[2] * x * y + D[3] * (r2 + 2.0 * x * x);
            double dy = D[2] * (r2 + 2.0 * y * y) + 2.0 * D[3] * x * y;
            x = (x0 - dx) * icd;
            y = (y0 - dy) * icd;
        }
    } else {
        double thd = sqrt(x * x + y * y);
        if (thd > 1e-9) {
            double th = thd;
            for (int it = 0; it < 20; ++it) {
                double t2 = th * th;
                double f = th * (1 + t2 * (D[0] + t2 * (D[1] + t2 * (D[2] + t2 * D[3])))) - thd;
                double df = 1 + t2 * (3 * D[0] + t2 * (5 * D[1] + t2 * (7 * D[2] + t2 * 9 * D[3])));
                th -= f / df;
            }
            double s = tan(th) / thd;
            x *= s;
            y *= s;
        }
    }
    out[0] = (float)x;
    out[1] = (float)y;
}

extern "C" {

int synth_default_config(mskf_config *cfg, const char *preset) { return mskf_fill_preset(cfg, preset); }

synth_ctx *synth_create(const mskf_config *cfg, uint32_t seed, double t0) {
    synth_ctx *c = new synth_ctx;
    synth_make_traj(&c->traj, seed, t0);
    double T1[16];
    mat4_mul(cfg->T_cn_cnm1, cfg->T_cam0_imu, T1);  // T_cam1_imu
    const double *Tc[2] = {cfg->T_cam0_imu, T1};
    const double *K[2] = {cfg->cam0_intrinsics, cfg->cam1_intrinsics};
    const double *D[2] = {cfg->cam0_distortion, cfg->cam1_distortion};
    const int model[2] = {cfg->cam0_model, cfg->cam1_model};
    for (int cam = 0; cam < 2; ++cam) {
        SynthCam &sc = c->cam[cam];
        sc.rows = cfg->img_rows;
        sc.cols = cfg->img_cols;
        // T_cam_imu maps imu -> cam; camera->imu rotation is its transpose, centre = -R^T t
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) sc.R_ci[i * 3 + j] = (float)Tc[cam][j * 4 + i];
            double s = 0;
            for (int k = 0; k < 3; ++k) s -= Tc[cam][k * 4 + i] * Tc[cam][k * 4 + 3];
            sc.t_ci[i] = (float)s;
        }
        c->rays[cam].resize((size_t)sc.rows * sc.cols * 8);
        static const double off[4][2] = {{-0.25, -0.25}, {0.25, -0.25}, {-0.25, 0.25}, {0.25, 0.25}};
        for (int r = 0; r < sc.rows; ++r)
            for (int q = 0; q < sc.cols; ++q)
                for (int s = 0; s < 4; ++s)
                    pixel_to_ray(q + off[s][0], r + off[s][1], K[cam], model[cam], D[cam],
                                 &c->rays[cam][((size_t)r * sc.cols + q) * 8 + s * 2]);
    }
    return c;
}

void synth_destroy(synth_ctx *c) { delete c; }

void synth_get_pose(const synth_ctx *c, double t, double R[9], double p[3]) { synth_pose(&c->traj, t, R, p); }

// One IMU row as the EuRoC runner would read it: values go through float (std::stof,
// apps/run_euroc_single_thread.cpp:220,225).  noise_scale scales the configured
// continuous-time noise densities (0 = noise free); sample index seeds the noise.
void synth_get_imu(const synth_ctx *c, double t, uint32_t sample_idx, double noise_gyro, double noise_acc,
                   double w[3], double a[3]) {
    synth_imu(&c->traj, t, w, a);
    for (int i = 0; i < 3; ++i) {
        // sum of 4 uniforms ~ gaussian enough for a synthetic sensor
        double ng = 0, na = 0;
        for (int k = 0; k < 4; ++k) {
            ng += synth_u01(c->traj.seed ^ 0xABCDu, sample_idx * 32u + i * 8u + k) - 0.5;
            na += synth_u01(c->traj.seed ^ 0x1234u, sample_idx * 32u + i * 8u + k) - 0.5;
        }
        w[i] = (double)(float)(w[i] + noise_gyro * ng * 1.7320508);
        a[i] = (double)(float)(a[i] + noise_acc * na * 1.7320508);
    }
}

void synth_render(const synth_ctx *c, double t, int cam, uint8_t *out, int n_threads) {
    double R[9], p[3];
    synth_pose(&c->traj, t, R, p);
    float Rwc[9], o[3];
    const SynthCam *sc = &c->cam[cam];
    synth_cam_pose(R, p, sc, Rwc, o);
    const float *rays = c->rays[cam].data();
    if (n_threads < 1) n_threads = 1;
    auto work = [&](int r0, int r1) {
        for (int r = r0; r < r1; ++r)
            for (int q = 0; q < sc->cols; ++q) out[(size_t)r * sc->cols + q] = synth_pixel(&c->traj, sc, rays, Rwc, o, r, q);
    };
    if (n_threads == 1) {
        work(0, sc->rows);
    } else {
        std::vector<std::thread> th;
        int per = (sc->rows + n_threads - 1) / n_threads;
        for (int i = 0; i < n_threads; ++i) {
            int r0 = i * per, r1 = std::min(sc->rows, r0 + per);
            if (r0 < r1) th.emplace_back(work, r0, r1);
        }
        for (auto &x : th) x.join();
    }
}

// ---- fleet helpers: many trajectories sharing one calibration (bench.py) ------------------
int synth_traj_size() { return (int)sizeof(SynthTraj); }
int synth_cam_size() { return (int)sizeof(SynthCam); }
void synth_make_trajs(const uint32_t *seeds, int n, double t0, SynthTraj *out) {
    for (int i = 0; i < n; ++i) synth_make_traj(&out[i], seeds[i], t0);
}
// IMU rows j0..j1-1 of every trajectory, rows {t, w, a} as synth_get_imu makes them; out [n][j1-j0][7]
void synth_imu_block(const SynthTraj *trajs, int n, int j0, int j1, double imu_dt, double noise_gyro, double noise_acc,
                     double *out) {
    synth_ctx tmp;
    for (int s = 0; s < n; ++s) {
        tmp.traj = trajs[s];
        for (int j = j0; j < j1; ++j) {
            double *o = out + ((size_t)s * (j1 - j0) + (j - j0)) * 7;
            o[0] = trajs[s].t0 + j * imu_dt;
            synth_get_imu(&tmp, o[0], (uint32_t)j, noise_gyro, noise_acc, o + 1, o + 4);
        }
    }
}
void synth_traj_pose(const SynthTraj *traj, double t, double R[9], double p[3]) { synth_pose(traj, t, R, p); }

const float *synth_rays(const synth_ctx *c, int cam) { return c->rays[cam].data(); }
const SynthTraj *synth_traj(const synth_ctx *c) { return &c->traj; }
const SynthCam *synth_cam(const synth_ctx *c, int cam) { return &c->cam[cam]; }

}  // extern "C"
