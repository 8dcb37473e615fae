// euroc_io.hpp — EuRoC "mav0" reader for the host examples: the csv parsing of the reference runner
// (apps/run_euroc_single_thread.cpp:150-238: nanosecond stamps split as seconds + 9 digits, file
// name with the trailing CR stripped, IMU values through std::stof) and an 8-bit grayscale image
// loader (non-interlaced PNG through zlib, or binary PGM) standing in for cv::imread(path, 0) (:194).
#pragma once
#include <zlib.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace euroc {

struct Gray8 {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    bool empty() const { return data.empty(); }
};

inline uint32_t be32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

inline Gray8 decode_png_gray8(const std::vector<uint8_t> &f) {
    static const uint8_t sig[8] = {137, 80, 78, 71, 13, 10, 26, 10};
    if (f.size() < 33 || std::memcmp(f.data(), sig, 8) != 0) throw std::runtime_error("not a PNG");
    Gray8 img;
    std::vector<uint8_t> z;
    size_t pos = 8;
    bool have_hdr = false;
    while (pos + 12 <= f.size()) {
        const uint32_t len = be32(&f[pos]);
        const char *type = (const char *)&f[pos + 4];
        if (pos + 12 + len > f.size()) throw std::runtime_error("truncated PNG");
        if (!std::strncmp(type, "IHDR", 4)) {
            img.cols = (int)be32(&f[pos + 8]);
            img.rows = (int)be32(&f[pos + 12]);
            const int depth = f[pos + 16], colour = f[pos + 17], interlace = f[pos + 20];
            if (depth != 8 || colour != 0 || interlace != 0) throw std::runtime_error("PNG must be 8-bit grayscale, non-interlaced");
            have_hdr = true;
        } else if (!std::strncmp(type, "IDAT", 4)) {
            z.insert(z.end(), f.begin() + pos + 8, f.begin() + pos + 8 + len);
        } else if (!std::strncmp(type, "IEND", 4)) {
            break;
        }
        pos += 12 + len;
    }
    if (!have_hdr) throw std::runtime_error("PNG without IHDR");
    const size_t stride = (size_t)img.cols + 1;
    std::vector<uint8_t> raw(stride * img.rows);
    uLongf out_len = (uLongf)raw.size();
    if (uncompress(raw.data(), &out_len, z.data(), (uLong)z.size()) != Z_OK || out_len != raw.size())
        throw std::runtime_error("PNG inflate failed");
    img.data.resize((size_t)img.rows * img.cols);
    for (int y = 0; y < img.rows; ++y) {
        const uint8_t ft = raw[y * stride];
        const uint8_t *in = &raw[y * stride + 1];
        uint8_t *cur = &img.data[(size_t)y * img.cols];
        const uint8_t *up = y ? cur - img.cols : nullptr;
        for (int x = 0; x < img.cols; ++x) {
            const int a = x ? cur[x - 1] : 0, b = up ? up[x] : 0, c = (x && up) ? up[x - 1] : 0;
            int pred = 0;
            switch (ft) {
                case 0: pred = 0; break;
                case 1: pred = a; break;
                case 2: pred = b; break;
                case 3: pred = (a + b) >> 1; break;
                case 4: {
                    const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
                    pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    break;
                }
                default: throw std::runtime_error("bad PNG filter");
            }
            cur[x] = (uint8_t)(in[x] + pred);
        }
    }
    return img;
}

inline Gray8 load_gray8(const std::string &path) {
    std::ifstream fs(path, std::ios::binary);
    if (!fs.good()) return Gray8();
    std::vector<uint8_t> f((std::istreambuf_iterator<char>(fs)), std::istreambuf_iterator<char>());
    if (f.size() > 2 && f[0] == 'P' && f[1] == '5') {  // binary PGM, maxval 255
        std::istringstream hs(std::string(f.begin(), f.begin() + std::min<size_t>(f.size(), 64)));
        std::string magic;
        int w = 0, h = 0, maxv = 0;
        hs >> magic >> w >> h >> maxv;
        const size_t off = (size_t)hs.tellg() + 1;
        Gray8 img;
        if (maxv != 255 || off + (size_t)w * h > f.size()) return img;
        img.rows = h;
        img.cols = w;
        img.data.assign(f.begin() + off, f.begin() + off + (size_t)w * h);
        return img;
    }
    return decode_png_gray8(f);
}

struct Stamped {
    double t;  // seconds
    std::string name;
};
struct ImuRow {
    double t, w[3], a[3];
};

// cam<n>/data.csv, run_euroc_single_thread.cpp:151-172
inline std::vector<Stamped> read_cam_csv(const std::string &path) {
    std::ifstream file(path);
    if (!file.good()) throw std::runtime_error("no cam file found: " + path);
    std::vector<Stamped> out;
    std::string line;
    std::getline(file, line);
    while (std::getline(file, line)) {
        if (line.size() < 11) continue;
        std::stringstream stream(line);
        std::string s;
        std::getline(stream, s, ',');
        const std::string nanoseconds = s.substr(s.size() - 9, 9), seconds = s.substr(0, s.size() - 9);
        const double stamp_ns = std::stoi(seconds) * 1e9 + std::stoi(nanoseconds);
        std::getline(stream, s, ',');
        out.push_back({stamp_ns * 1e-9, s.substr(0, s.size() - 1)});  // the reference drops the last char (CR of CRLF)
    }
    return out;
}

// one row of imu0/data.csv, run_euroc_single_thread.cpp:211-232 (values go through float)
inline bool parse_imu_row(const std::string &line, ImuRow &row) {
    if (line.size() < 11) return false;
    std::stringstream stream(line);
    std::string s;
    std::getline(stream, s, ',');
    const std::string nanoseconds = s.substr(s.size() - 9, 9), seconds = s.substr(0, s.size() - 9);
    row.t = (std::stoi(seconds) * 1e9 + std::stoi(nanoseconds)) * 1e-9;
    for (int j = 0; j < 3; ++j) {
        std::getline(stream, s, ',');
        row.w[j] = std::stof(s);
    }
    for (int j = 0; j < 3; ++j) {
        std::getline(stream, s, ',');
        row.a[j] = std::stof(s);
    }
    return true;
}

// cg::RotationMatrix::quarternion_hamilton() of a row-major 3x3 (msckf_vio.cpp:1251), order x y z w
inline void rot_to_quat_hamilton(const double R[9], double q[4]) {
    const double tr = R[0] + R[4] + R[8];
    double x, y, z, w;
    if (tr > 0) {
        double s = std::sqrt(tr + 1.0) * 2;
        w = 0.25 * s; x = (R[7] - R[5]) / s; y = (R[2] - R[6]) / s; z = (R[3] - R[1]) / s;
    } else if (R[0] > R[4] && R[0] > R[8]) {
        double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2;
        w = (R[7] - R[5]) / s; x = 0.25 * s; y = (R[1] + R[3]) / s; z = (R[2] + R[6]) / s;
    } else if (R[4] > R[8]) {
        double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2;
        w = (R[2] - R[6]) / s; x = (R[1] + R[3]) / s; y = 0.25 * s; z = (R[5] + R[7]) / s;
    } else {
        double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2;
        w = (R[3] - R[1]) / s; x = (R[2] + R[6]) / s; y = (R[5] + R[7]) / s; z = 0.25 * s;
    }
    q[0] = x; q[1] = y; q[2] = z; q[3] = w;
}

}  // namespace euroc
