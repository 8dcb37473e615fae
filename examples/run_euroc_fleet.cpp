// run_euroc_fleet.cpp — the mav0 feed of apps/run_euroc_single_thread.cpp:150-254 for a FLEET of streams:
//   run_euroc_fleet [--threads T] [--ring R] [--repeat N] [--preset P] [--decimals D] [--out DIR] [--device G]
//                   <mav0 dir> [<mav0 dir> ...]
// One engine handle drives all streams (stream s = the s-th directory; --repeat N feeds every directory N
// times, the way a load test is built from a few recordings).  T decoder threads turn the PNG/PGM files of
// step k into slot k mod R of a PAGE-LOCKED ring laid out like the engine's landing area
// ([stream][cam][pixels]); the feeding thread pushes each stream's IMU rows up to its image stamp (same
// loop as the reference runner, :206-238), uploads the slot with ONE asynchronous copy
// (mskf_push_stereo_batch), runs mskf_step, and writes the poses of the step before (already on the host,
// mskf_get_poses_prev) as one TUM file per stream: DIR/pose_<s>.txt (msckf_vio.cpp:1255-1258).
// A slot goes back to the decoders when its upload has left it (mskf_wait_uploads), so disk, decode,
// host->device copy and the GPU step of consecutive frames overlap.  Prints disk->pose frames/s.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <thread>

#include "../include/msckf_b200.hpp"
#include "euroc_io.hpp"

using namespace mskf_b200;

namespace {

struct StreamFeed {
    std::string dir;
    std::vector<euroc::Stamped> cam[2];
    std::ifstream imu_file;
    size_t imu_rows = 0;    // rows pushed so far: the 200th sets gravity (msckf_vio.cpp:198-204)
    FILE *out = nullptr;
};

// Decode jobs of one step: (stream, cam) pairs handed out by an atomic counter.
struct Slot {
    uint8_t *base = nullptr;
    std::atomic<int> next{0}, done{0};
    long long step = -1;     // which step the decoders should fill; -1 = not yet released
    bool failed = false;
};

}  // namespace

int main(int argc, char **argv) {
    int n_threads = (int)std::max(1u, std::thread::hardware_concurrency()), ring = 3, repeat = 1, decimals = 6, device = 0;
    std::string preset = "ref", out_dir = ".";
    std::vector<std::string> dirs;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { std::fprintf(stderr, "%s needs a value\n", name); std::exit(2); }
            return argv[++i];
        };
        if (a == "--threads") n_threads = std::atoi(val("--threads"));
        else if (a == "--ring") ring = std::atoi(val("--ring"));
        else if (a == "--repeat") repeat = std::atoi(val("--repeat"));
        else if (a == "--preset") preset = val("--preset");
        else if (a == "--decimals") decimals = std::atoi(val("--decimals"));
        else if (a == "--out") out_dir = val("--out");
        else if (a == "--device") device = std::atoi(val("--device"));
        else dirs.push_back(a);
    }
    if (dirs.empty() || n_threads < 1 || ring < 2 || repeat < 1) {
        std::fprintf(stderr, "usage: %s [--threads T] [--ring R>=2] [--repeat N] [--preset P] [--decimals D] [--out DIR] [--device G] <mav0 dir>...\n", argv[0]);
        return 2;
    }
    try {
        const int S = (int)dirs.size() * repeat;
        std::vector<StreamFeed> feed(S);
        size_t n_steps = (size_t)-1;
        for (int s = 0; s < S; ++s) {
            StreamFeed &f = feed[s];
            f.dir = dirs[s / repeat];
            for (int j = 0; j < 2; ++j) f.cam[j] = euroc::read_cam_csv(f.dir + "/cam" + std::to_string(j) + "/data.csv");
            if (f.cam[0].empty() || f.cam[0].size() != f.cam[1].size()) throw std::runtime_error("cam0/cam1 csv mismatch in " + f.dir);
            f.imu_file.open(f.dir + "/imu0/data.csv");
            if (!f.imu_file.good()) throw std::runtime_error("no imu file found in " + f.dir);
            std::string header;
            std::getline(f.imu_file, header);
            n_steps = std::min(n_steps, f.cam[0].size());
            const std::string p = out_dir + "/pose_" + std::to_string(s) + ".txt";
            f.out = std::fopen(p.c_str(), "w");
            if (!f.out) throw std::runtime_error("cannot write " + p);
        }
        euroc::Gray8 first = euroc::load_gray8(feed[0].dir + "/cam0/data/" + feed[0].cam[0][0].name);
        if (first.empty()) throw std::runtime_error("cannot read " + feed[0].cam[0][0].name);
        mskf_config cfg = default_config(preset);
        cfg.img_rows = first.rows;
        cfg.img_cols = first.cols;
        const size_t img = (size_t)first.rows * first.cols, set_bytes = 2 * img * (size_t)S;

        mskf_handle *h = nullptr;
        if (mskf_create(&cfg, S, device, &h) != MSKF_OK) throw std::runtime_error(std::string("mskf_create: ") + mskf_last_error(h));
        auto check = [&](int rc, const char *what) {
            if (rc != MSKF_OK) throw std::runtime_error(std::string(what) + ": " + mskf_last_error(h));
        };

        std::vector<Slot> slots(ring);
        struct Owner {  // declared before the decoder pool: the ring and the engine outlive the threads
            std::vector<Slot> &slots; mskf_handle *&h;
            ~Owner() {
                for (Slot &sl : slots) mskf_host_free(sl.base);
                mskf_destroy(h);
            }
        } owner{slots, h};
        for (Slot &sl : slots) {
            void *p = nullptr;
            if (mskf_host_alloc(&p, set_bytes) != MSKF_OK) throw std::runtime_error("page-locked ring allocation failed");
            sl.base = (uint8_t *)p;
        }
        std::mutex mu;
        std::condition_variable cv_work, cv_done;
        bool quit = false;
        const int jobs_per_step = 2 * S;

        // decoders: take the oldest released slot that still has jobs, decode (stream, cam) into its place
        auto decoder = [&]() {
            for (;;) {
                Slot *sl = nullptr;
                int job = -1;
                {
                    std::unique_lock<std::mutex> lk(mu);
                    for (;;) {
                        if (quit) return;
                        long long best = -1;
                        for (Slot &c : slots)
                            if (c.step >= 0 && c.next.load() < jobs_per_step && (best < 0 || c.step < best)) { best = c.step; sl = &c; }
                        if (sl) { job = sl->next.fetch_add(1); if (job < jobs_per_step) break; sl = nullptr; continue; }
                        cv_work.wait(lk);
                    }
                }
                const int s = job >> 1, j = job & 1;
                bool ok = true;
                try {
                    euroc::Gray8 g = euroc::load_gray8(feed[s].dir + "/cam" + std::to_string(j) + "/data/" + feed[s].cam[j][(size_t)sl->step].name);
                    if (g.rows != cfg.img_rows || g.cols != cfg.img_cols) ok = false;
                    else std::memcpy(sl->base + ((size_t)s * 2 + j) * img, g.data.data(), img);
                } catch (const std::exception &) { ok = false; }
                {
                    std::lock_guard<std::mutex> lk(mu);
                    if (!ok) sl->failed = true;
                    if (sl->done.fetch_add(1) + 1 == jobs_per_step) cv_done.notify_all();
                }
            }
        };
        struct Pool {  // stops and joins the decoders on every way out of the block, errors included
            std::vector<std::thread> threads;
            std::mutex &mu; std::condition_variable &cv; bool &quit;
            ~Pool() {
                { std::lock_guard<std::mutex> lk(mu); quit = true; cv.notify_all(); }
                for (std::thread &t : threads) t.join();
            }
        } pool{{}, mu, cv_work, quit};
        for (int i = 0; i < n_threads; ++i) pool.threads.emplace_back(decoder);
        auto release = [&](int slot, long long step) {
            std::lock_guard<std::mutex> lk(mu);
            Slot &sl = slots[slot];
            sl.next = 0; sl.done = 0; sl.failed = false;
            sl.step = step < (long long)n_steps ? step : -1;
            cv_work.notify_all();
        };
        for (int r = 0; r < ring; ++r) release(r, r);

        std::vector<double> stamps(S), poses((size_t)16 * S);
        std::vector<double> t_hist((size_t)S * 2);            // stamps of the last two steps, per stream
        std::vector<uint8_t> grav_hist((size_t)S * 2, 0);
        auto write_poses = [&](size_t k) {                    // poses of step k sit in `poses`
            for (int s = 0; s < S; ++s) {
                if (!grav_hist[(k & 1) * S + s]) continue;    // MsckfVio::featureCallback returns before publish (msckf_vio.cpp:308)
                const double *T = &poses[(size_t)16 * s];
                double R[9], q[4];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) R[i * 3 + j] = T[i * 4 + j];
                euroc::rot_to_quat_hamilton(R, q);
                std::fprintf(feed[s].out, "%.*f %.*f %.*f %.*f %.*f %.*f %.*f %.*f\n", decimals, t_hist[(k & 1) * S + s], decimals, T[3],
                             decimals, T[7], decimals, T[11], decimals, q[0], decimals, q[1], decimals, q[2], decimals, q[3]);
            }
        };

        const auto t0 = std::chrono::steady_clock::now();
        double wait_decode = 0;
        std::vector<double> imu_rows;
        std::string line;
        for (size_t k = 0; k < n_steps; ++k) {
            Slot &sl = slots[k % ring];
            {
                const auto w0 = std::chrono::steady_clock::now();
                std::unique_lock<std::mutex> lk(mu);
                cv_done.wait(lk, [&] { return sl.done.load() == jobs_per_step; });
                wait_decode += std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
                if (sl.failed) throw std::runtime_error("ERROR: img is empty !!! (step " + std::to_string(k) + ")");
            }
            for (int s = 0; s < S; ++s) {
                StreamFeed &f = feed[s];
                const double t_img = f.cam[0][k].t;
                stamps[s] = t_img;
                imu_rows.clear();
                double t_imu = 0.0;
                do {  // run_euroc_single_thread.cpp:206-238
                    if (!std::getline(f.imu_file, line)) break;
                    euroc::ImuRow row;
                    if (!euroc::parse_imu_row(line, row)) continue;
                    const double v[7] = {row.t, row.w[0], row.w[1], row.w[2], row.a[0], row.a[1], row.a[2]};
                    imu_rows.insert(imu_rows.end(), v, v + 7);
                    t_imu = row.t;
                    ++f.imu_rows;
                } while (t_imu <= t_img);
                check(mskf_push_imu_batch(h, s, (int)(imu_rows.size() / 7), imu_rows.data()), "mskf_push_imu_batch");
                t_hist[(k & 1) * S + s] = t_img;
                grav_hist[(k & 1) * S + s] = f.imu_rows >= 200;
            }
            if (k > 0) {  // the previous upload has had a whole step to finish: its slot goes back to the decoders
                check(mskf_wait_uploads(h), "mskf_wait_uploads");
                release((int)((k - 1) % ring), (long long)(k - 1 + ring));
            }
            check(mskf_push_stereo_batch(h, stamps.data(), sl.base, sl.base + img, 2 * img), "mskf_push_stereo_batch");
            check(mskf_step(h), "mskf_step");
            if (k > 0) {
                check(mskf_get_poses_prev(h, poses.data(), S), "mskf_get_poses_prev");
                write_poses(k - 1);
            }
        }
        check(mskf_sync(h), "mskf_sync");
        check(mskf_get_poses(h, poses.data(), S), "mskf_get_poses");
        write_poses(n_steps - 1);
        const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        for (StreamFeed &f : feed) std::fclose(f.out);
        std::printf("{\"streams\": %d, \"steps\": %zu, \"decode_threads\": %d, \"ring\": %d, \"disk_to_pose_frames_per_s\": %.1f, "
                    "\"wall_s\": %.3f, \"feeder_waited_for_decode_s\": %.3f, \"h2d_bytes_per_step\": %zu}\n",
                    S, n_steps, n_threads, ring, S * (double)n_steps / wall, wall, wait_decode, set_bytes);
        std::fflush(stdout);
    } catch (const std::exception &e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
