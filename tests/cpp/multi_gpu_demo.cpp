// tests/cpp/multi_gpu_demo.cpp -- BASELINE configs[4] from a C++ host, no Python: one process, one fbus_handle per GPU, the
// filters sharded into contiguous global index ranges with no hot-path traffic, and ONE collective at the end:
// fbus_stats_allreduce (ncclAllReduce of the 8-double statistics vectors over NVLink).
//
//   multi_gpu_demo traj.bin total_filters n_gpus [seconds]
//
// traj.bin (written by tests/test_gpu_multi_handle.py): int64 N, W ; double t_imu[N], t_frames[W], base_imu[N*6], base_pose[W*7],
// truth_p[3], truth_q[4] ; uint32 win_off[W+1].  Prints the combined statistics vector and the per-GPU shard sizes.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/fbus_ekf.h"

#define CK(x)                                                                                     \
    do {                                                                                          \
        const int rc_ = (x);                                                                      \
        if (rc_ != 0) { std::fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, fbus_last_error(nullptr)); return 1; } \
    } while (0)
#define CU(x)                                                                                     \
    do {                                                                                          \
        const cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } \
    } while (0)

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: multi_gpu_demo traj.bin total_filters n_gpus [seconds]\n"); return 2; }
    FILE* fp = std::fopen(argv[1], "rb");
    if (!fp) { std::perror(argv[1]); return 2; }
    const size_t total = std::strtoull(argv[2], nullptr, 10);
    const int n_gpu = std::atoi(argv[3]);
    const int seconds = argc > 4 ? std::atoi(argv[4]) : 1;
    int64_t NW[2];
    if (std::fread(NW, sizeof(int64_t), 2, fp) != 2) return 2;
    const size_t N = (size_t)NW[0], W = (size_t)NW[1];
    std::vector<double> t_imu(N), t_frames(W), base_imu(N * 6), base_pose(W * 7), truth(7);
    std::vector<uint32_t> win_off(W + 1);
    bool ok = std::fread(t_imu.data(), 8, N, fp) == N && std::fread(t_frames.data(), 8, W, fp) == W &&
              std::fread(base_imu.data(), 8, N * 6, fp) == N * 6 && std::fread(base_pose.data(), 8, W * 7, fp) == W * 7 &&
              std::fread(truth.data(), 8, 7, fp) == 7 && std::fread(win_off.data(), 4, W + 1, fp) == W + 1;
    std::fclose(fp);
    if (!ok) { std::fprintf(stderr, "short trajectory file\n"); return 2; }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (n_gpu < 1 || n_gpu > ndev) { std::fprintf(stderr, "need %d GPUs, have %d\n", n_gpu, ndev); return 3; }
    fbus_config cfg;
    CK(fbus_config_default(&cfg));

    struct Shard {
        fbus_handle* h = nullptr;
        size_t lo = 0, n = 0;
        double *imu = nullptr, *pose = nullptr, *tp = nullptr, *tq = nullptr, *stats = nullptr;
        int32_t* id = nullptr;
    };
    std::vector<Shard> sh(n_gpu);
    const size_t base = total / n_gpu, rem = total % n_gpu;  // contiguous ranges, sizes differ by at most one (shard.shard_range)
    for (int r = 0; r < n_gpu; ++r) {
        Shard& s = sh[r];
        s.lo = r * base + ((size_t)r < rem ? r : rem);
        s.n = base + ((size_t)r < rem ? 1 : 0);
        CK(fbus_create(&s.h, &cfg, r, s.n));
        CU(cudaSetDevice(r));
        CU(cudaMalloc(&s.imu, N * 6 * s.n * 8));
        CU(cudaMalloc(&s.pose, W * 7 * s.n * 8));
        CU(cudaMalloc(&s.id, W * s.n * 4));
        CU(cudaMalloc(&s.tp, 3 * s.n * 8));
        CU(cudaMalloc(&s.tq, 4 * s.n * 8));
        CU(cudaMalloc(&s.stats, FBUS_NSTATS * 8));
        std::vector<double> tp(3 * s.n), tq(4 * s.n);
        for (size_t b = 0; b < s.n; ++b) {
            for (int c = 0; c < 3; ++c) tp[c * s.n + b] = truth[c];
            for (int c = 0; c < 4; ++c) tq[c * s.n + b] = truth[3 + c];
        }
        CU(cudaMemcpy(s.tp, tp.data(), tp.size() * 8, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s.tq, tq.data(), tq.size() * 8, cudaMemcpyHostToDevice));
        fbus_synth_spec sp = {};
        sp.n_samples = N; sp.n_frames = W; sp.base_imu = base_imu.data(); sp.base_pose = base_pose.data();
        sp.marker_id = 0;
        sp.sigma_acc = 0.015; sp.sigma_gyro = 1e-3; sp.sigma_ba = 0.05; sp.sigma_bg = 2e-3; sp.sigma_pos = 2.5e-4; sp.sigma_quat = 1.5e-3;
        sp.seed = 20260117 + 5;
        sp.filter_offset = s.lo;  // Philox streams are keyed by the GLOBAL filter index: the shards together are the single-GPU batch
        CK(fbus_synth_streams(s.h, &sp, s.imu, s.id, s.pose, nullptr));
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int k = 0; k < seconds; ++k) {  // the trajectory is periodic: every second replays the streams with advanced timestamps
        std::vector<double> ti(t_imu), tf(t_frames);
        for (auto& x : ti) x += k;
        for (auto& x : tf) x += k;
        for (int r = 0; r < n_gpu; ++r) {  // asynchronous: all GPUs work at the same time
            Shard& s = sh[r];
            fbus_imu_stream imu = {N, s.n, ti.data(), s.imu, FBUS_MEM_DEVICE, FBUS_IMU_F64_SI};
            fbus_det_frames det = {W, 1, s.n, tf.data(), s.id, s.pose, FBUS_MEM_DEVICE, 0};
            CK(fbus_step_windows(s.h, &imu, &det, win_off.data(), 0, W, nullptr, FBUS_MEM_HOST));
        }
        for (int r = 0; r < n_gpu; ++r) CK(fbus_synchronize(sh[r].h));  // ti / tf must outlive the staged copies
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<fbus_handle*> hs(n_gpu);
    std::vector<double*> vecs(n_gpu);
    for (int r = 0; r < n_gpu; ++r) {
        CK(fbus_stats(sh[r].h, sh[r].tp, sh[r].tq, FBUS_MEM_DEVICE, nullptr, sh[r].stats));
        hs[r] = sh[r].h;
        vecs[r] = sh[r].stats;
    }
    double out[FBUS_NSTATS];
    CK(fbus_stats_allreduce(hs.data(), n_gpu, vecs.data(), out));
    std::printf("stats");
    for (int i = 0; i < FBUS_NSTATS; ++i) std::printf(" %.17g", out[i]);
    std::printf("\nshards");
    for (int r = 0; r < n_gpu; ++r) std::printf(" %zu", sh[r].n);
    std::printf("\nfilter_steps_per_s %.6g\n", (double)total * (N + W) * seconds / sec);
    // every rank's device vector holds the combined result
    for (int r = 0; r < n_gpu; ++r) {
        double v[FBUS_NSTATS];
        CU(cudaSetDevice(r));
        CU(cudaMemcpy(v, sh[r].stats, sizeof v, cudaMemcpyDeviceToHost));
        for (int i = 0; i < FBUS_NSTATS; ++i)
            if (v[i] != out[i]) { std::fprintf(stderr, "rank %d entry %d differs\n", r, i); return 4; }
    }
    for (int r = 0; r < n_gpu; ++r) fbus_destroy(sh[r].h);
    return 0;
}
