// scene.h — seeded synthetic EuRoC-style stereo+IMU generator ("textured room").
//
// Not part of the hot path and not part of the oracle: it only makes the inputs that both
// are fed (BASELINE.json config 1: 752x480 stereo @20 Hz, IMU @200 Hz, textured room,
// >=1.2 s static start so MsckfVio::initializeGravityAndBias (msckf_vio.cpp:198) sees 200
// static samples).  The same header is compiled for the host (synth_cpu.cpp, used by the
// CPU tests and the CPU baseline) and for the device (synth.cu, used by bench.py), so the
// engine and the oracle can always be handed byte-identical images.
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD inline
#endif

struct SynthTraj {
    double t0;          // time stamp of the first sample (s)
    double t_static;    // motion starts at t0 + t_static
    double amp_p[3], frq_p[3], phs_p[3];
    double amp_r[3], frq_r[3], phs_r[3];
    double p0[3];
    double room[3];     // half extents of the box (m)
    uint32_t seed;
    uint32_t pad;
};

struct SynthCam {
    float R_ci[9];      // camera -> imu rotation (row-major)
    float t_ci[3];      // camera centre in the imu frame
    int rows, cols;
};

SYNTH_HD uint32_t synth_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    uint32_t h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u ^ (c + 0x165667B1u) * 0xC2B2AE3Du ^ d * 0x27D4EB2Fu;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}
SYNTH_HD double synth_u01(uint32_t seed, uint32_t k) { return (double)(synth_hash(seed, k, 0x51u, 7u) >> 8) / 16777216.0; }

inline void synth_make_traj(SynthTraj *tr, uint32_t seed, double t0) {
    tr->t0 = t0;
    tr->t_static = 1.5;
    tr->seed = seed;
    tr->pad = 0;
    tr->room[0] = 3.0 + 1.0 * synth_u01(seed, 1);
    tr->room[1] = 3.0 + 1.0 * synth_u01(seed, 2);
    tr->room[2] = 2.0 + 0.5 * synth_u01(seed, 3);
    for (int i = 0; i < 3; ++i) {
        tr->amp_p[i] = (i == 2 ? 0.25 : 0.45) * (0.6 + 0.4 * synth_u01(seed, 10 + i));
        tr->frq_p[i] = 0.08 + 0.09 * synth_u01(seed, 20 + i);
        tr->phs_p[i] = 6.283185307179586 * synth_u01(seed, 30 + i);
        tr->amp_r[i] = (i == 0 ? 0.30 : 0.12) * (0.6 + 0.4 * synth_u01(seed, 40 + i));
        tr->frq_r[i] = 0.07 + 0.08 * synth_u01(seed, 50 + i);
        tr->phs_r[i] = 6.283185307179586 * synth_u01(seed, 60 + i);
        tr->p0[i] = (i == 2 ? 0.2 : 0.5) * (synth_u01(seed, 70 + i) - 0.5);
    }
}

// imu pose in the world: R_wi (imu -> world, row-major) and p (imu origin in world)
SYNTH_HD void synth_pose(const SynthTraj *tr, double t, double R[9], double p[3]) {
    double s = t - tr->t0 - tr->t_static;
    double ramp = 0.0;
    if (s > 0) {
        double x = s / 2.0;
        ramp = x >= 1.0 ? 1.0 : x * x * x * (x * (6.0 * x - 15.0) + 10.0);
    } else {
        s = 0;
    }
    double ang[3];
    for (int i = 0; i < 3; ++i) {
        p[i] = tr->p0[i] + ramp * tr->amp_p[i] * (sin(6.283185307179586 * tr->frq_p[i] * s + tr->phs_p[i]) - sin(tr->phs_p[i]));
        ang[i] = ramp * tr->amp_r[i] * (sin(6.283185307179586 * tr->frq_r[i] * s + tr->phs_r[i]) - sin(tr->phs_r[i]));
    }
    // local perturbation Rz(yaw about imu x = world up) * Ry * Rx expressed in the imu frame
    double cy = cos(ang[0]), sy = sin(ang[0]), cp = cos(ang[1]), sp = sin(ang[1]), cr = cos(ang[2]), sr = sin(ang[2]);
    // rotation about imu x (up) by yaw, then imu y by pitch, then imu z by roll
    double Rx[9] = {1, 0, 0, 0, cy, -sy, 0, sy, cy};
    double Ry[9] = {cp, 0, sp, 0, 1, 0, -sp, 0, cp};
    double Rz[9] = {cr, -sr, 0, sr, cr, 0, 0, 0, 1};
    // base: imu x = world z (up), imu y = -world y, imu z = world x (forward)
    const double B[9] = {0, 0, 1, 0, -1, 0, 1, 0, 0};
    double T1[9], T2[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += Rx[i * 3 + k] * Ry[k * 3 + j];
            T1[i * 3 + j] = a;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += T1[i * 3 + k] * Rz[k * 3 + j];
            T2[i * 3 + j] = a;
        }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double a = 0;
            for (int k = 0; k < 3; ++k) a += B[i * 3 + k] * T2[k * 3 + j];
            R[i * 3 + j] = a;
        }
}

// noise-free IMU sample at time t: body rates and specific force in the imu frame
inline void synth_imu(const SynthTraj *tr, double t, double w[3], double a[3]) {
    const double h = 1e-3;
    double R0[9], R1[9], R2[9], p0[3], p1[3], p2[3];
    synth_pose(tr, t - h, R0, p0);
    synth_pose(tr, t, R1, p1);
    synth_pose(tr, t + h, R2, p2);
    // dR = R0^T R2 ~ exp([w] 2h)
    double dR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += R0[k * 3 + i] * R2[k * 3 + j];
            dR[i * 3 + j] = s;
        }
    w[0] = (dR[7] - dR[5]) / (4 * h);
    w[1] = (dR[2] - dR[6]) / (4 * h);
    w[2] = (dR[3] - dR[1]) / (4 * h);
    double aw[3];
    for (int i = 0; i < 3; ++i) aw[i] = (p2[i] - 2 * p1[i] + p0[i]) / (h * h);
    aw[2] += 9.81;  // specific force = a - g, g = (0,0,-9.81)
    for (int i = 0; i < 3; ++i) a[i] = R1[0 * 3 + i] * aw[0] + R1[1 * 3 + i] * aw[1] + R1[2 * 3 + i] * aw[2];
}

SYNTH_HD float synth_texture(uint32_t seed, int wall, float a, float b) {
    float v = 0.f;
    const float cell[3] = {0.31f, 0.097f, 0.031f};
    const float wgt[3] = {0.45f, 0.35f, 0.20f};
#pragma unroll
    for (int o = 0; o < 3; ++o) {
        int ia = (int)floorf(a / cell[o]), ib = (int)floorf(b / cell[o]);
        uint32_t h = synth_hash(seed + 977u * (uint32_t)wall, (uint32_t)ia, (uint32_t)ib, (uint32_t)o + 1u);
        v += wgt[o] * (float)(h >> 8) * (1.0f / 16777216.0f);
    }
    return 20.f + 215.f * v;
}

// One pixel of camera `cam`: rays holds, per pixel, 4 sub-sample normalised undistorted
// directions (x, y) laid out [row][col][4][2].
SYNTH_HD uint8_t synth_pixel(const SynthTraj *tr, const SynthCam *cam, const float *rays, const float Rwc[9],
                             const float o[3], int row, int col) {
    const float *r = rays + ((size_t)row * cam->cols + col) * 8;
    float acc = 0.f;
#pragma unroll
    for (int sidx = 0; sidx < 4; ++sidx) {
        float dx = r[sidx * 2], dy = r[sidx * 2 + 1];
        float d[3];
        for (int i = 0; i < 3; ++i) d[i] = Rwc[i * 3] * dx + Rwc[i * 3 + 1] * dy + Rwc[i * 3 + 2];
        float tbest = 1e30f;
        int wall = 0;
        for (int ax = 0; ax < 3; ++ax) {
            if (fabsf(d[ax]) < 1e-9f) continue;
            float plane = d[ax] > 0 ? (float)tr->room[ax] : -(float)tr->room[ax];
            float tt = (plane - o[ax]) / d[ax];
            if (tt > 0 && tt < tbest) {
                tbest = tt;
                wall = ax * 2 + (d[ax] > 0 ? 1 : 0);
            }
        }
        float hit[3] = {o[0] + tbest * d[0], o[1] + tbest * d[1], o[2] + tbest * d[2]};
        int ax = wall >> 1;
        float a = hit[(ax + 1) % 3], b = hit[(ax + 2) % 3];
        acc += synth_texture(tr->seed, wall, a, b);
    }
    float v = acc * 0.25f + 0.5f;
    return (uint8_t)(v < 0.f ? 0.f : (v > 255.f ? 255.f : v));
}

// camera pose in the world from the imu pose
SYNTH_HD void synth_cam_pose(const double Rwi[9], const double pwi[3], const SynthCam *cam, float Rwc[9], float o[3]) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += Rwi[i * 3 + k] * (double)cam->R_ci[k * 3 + j];
            Rwc[i * 3 + j] = (float)s;
        }
        double s = pwi[i];
        for (int k = 0; k < 3; ++k) s += Rwi[i * 3 + k] * (double)cam->t_ci[k];
        o[i] = (float)s;
    }
}
