// fbus_filter.hpp -- header-only C++ shim with the public surface of the reference's FBUSEKF::FILTER
// (C++/include/filter.hpp:143-178) on top of the C ABI of fbus_ekf.h, for a batch of ONE filter.
//
// A maintainer of the reference replaces `#include "filter.hpp"` by this header (and links libfbus_ekf.so); the IMU
// callback keeps calling SetImuData(), the vision thread keeps calling SetDetectionResult() +
// SetDetectionResultUpdated(), the viewer keeps calling GetCameraPose()/GetVisualPose().  What the reference's filter
// thread did on every wake-up (filter.cpp:207-235) happens inside SetDetectionResultUpdated() on the GPU.
// No Eigen dependency: poses are row-major 4x4 std::array<double,16>.
#ifndef FBUS_FILTER_HPP
#define FBUS_FILTER_HPP

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "fbus_ekf.h"

namespace FBUSB200 {

struct IMUData {  // common.hpp:176-193
    double timeStamp;
    double accel[3];
    double gyro[3];
};
struct DetectionResult {  // filter.hpp:39-56
    int markerID;
    double timeStamp;
    double positionAtCL[3];
    double quaternionM2CL[4];  // w, x, y, z
};
typedef std::vector<DetectionResult> DetectionResultList;
typedef std::array<double, 16> Matrix4d;

class FILTER {
  public:
    static constexpr size_t IMU_BUFFER_MAX_SIZE = 2000;  // filter.hpp:25

    explicit FILTER(const fbus_config& cfg, int device = 0, bool iir = true) : iir_(iir), cfg_(cfg) {
        if (fbus_abi_version() != FBUS_ABI_VERSION) throw std::runtime_error("libfbus_ekf.so was built from another version of fbus_ekf.h");
        if (fbus_create(&h_, &cfg, device, 1) != FBUS_OK) throw std::runtime_error(std::string("fbus_create: ") + fbus_last_error(nullptr));
    }
    ~FILTER() {
        StopFilterThread();
        fbus_destroy(h_);
    }
    FILTER(const FILTER&) = delete;
    FILTER& operator=(const FILTER&) = delete;

    // FILTER::InputIMUData (filter.hpp:143): the static callback a data generator / driver registers
    static void InputIMUData(IMUData imudata, void* pObject) {
        FILTER* self = static_cast<FILTER*>(pObject);
        self->SetImuData(imudata);
        self->SetImuDataUpdated();
    }

    // FILTER::SetImuData (filter.cpp:24-55): 1-pole IIR on the incoming sample, bounded buffer
    void SetImuData(const IMUData& raw) {
        std::lock_guard<std::mutex> lk(mu_);
        IMUData f = raw;
        if (iir_ && !buf_.empty()) {
            const IMUData& p = buf_.back();
            for (int i = 0; i < 3; ++i) {
                f.accel[i] = p.accel[i] * (1 - 0.1) + raw.accel[i] * 0.1;
                f.gyro[i] = p.gyro[i] * (1 - 0.1) + raw.gyro[i] * 0.1;
            }
        }
        buf_.push_back(f);
        if (buf_.size() > IMU_BUFFER_MAX_SIZE) buf_.erase(buf_.begin(), buf_.begin() + 500);
    }
    void SetImuDataUpdated() {}

    // FILTER::InitializeGravityAndBias (filter.cpp:256-285): uses and clears everything buffered so far
    void InitializeGravityAndBias() {
        std::lock_guard<std::mutex> lk(mu_);
        if (buf_.empty()) return;
        std::vector<double> t, d;
        pack(buf_.size(), t, d);
        fbus_imu_stream s{buf_.size(), 1, t.data(), d.data(), FBUS_MEM_HOST, 0};
        check(fbus_init_gravity_gyrobias(h_, &s, 0, buf_.size()));
        buf_.clear();
    }

    void SetDetectionResult(const DetectionResultList& r) {
        std::lock_guard<std::mutex> lk(mu_);
        det_ = r;
    }
    // the reference notifies the filter thread here (filter.hpp:166-170).  With a filter thread running
    // (StartFilterThread) this does the same; without one the frame body (init | reset -> propagate -> update,
    // filter.cpp:207-235) runs on the GPU before this call returns.
    void SetDetectionResultUpdated() {
        if (threaded_.load()) {  // producers never touch the std::thread object itself
            std::lock_guard<std::mutex> lk(cv_mu_);
            pending_ = true;
            cv_.notify_one();
            return;
        }
        ProcessFrame();
    }

    // FILTER::StartFilterThread / FilterThreadFunction / JoinFilterThread (filter.cpp:181-250): wait for IMU data
    // (1000 ms in the reference), InitializeGravityAndBias, then one frame body per notification.  The reference's loop
    // never ends; StopFilterThread (not in the reference) lets the thread leave it so that the object can be destroyed.
    void StartFilterThread(int init_wait_ms = 1000) {
        if (thread_.joinable()) return;
        stop_ = false;
        ready_ = false;
        threaded_.store(true);
        thread_ = std::thread(&FILTER::FilterThreadFunction, this, init_wait_ms);
    }
    void JoinFilterThread() {
        if (thread_.joinable()) thread_.join();
        threaded_.store(false);
    }
    void StopFilterThread() {
        {
            std::lock_guard<std::mutex> lk(cv_mu_);
            stop_ = true;
            cv_.notify_all();
        }
        JoinFilterThread();
    }
    void FilterThreadFunction(int init_wait_ms) {
        std::this_thread::sleep_for(std::chrono::milliseconds(init_wait_ms));
        InitializeGravityAndBias();
        {
            std::lock_guard<std::mutex> lk(cv_mu_);
            ready_ = true;
            idle_.notify_all();
        }
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(cv_mu_);
                cv_.wait(lk, [&] { return stop_ || pending_; });
                if (stop_) return;
                pending_ = false;
                busy_ = true;
            }
            try {
                ProcessFrame();
            } catch (const std::exception& e) {
                std::lock_guard<std::mutex> lk(cv_mu_);
                thread_error_ = e.what();
            }
            std::lock_guard<std::mutex> lk(cv_mu_);
            busy_ = false;
            idle_.notify_all();
        }
    }
    // blocks until the filter thread has initialised and worked off every notified frame (for tests and log replays;
    // the live callers never wait).  Rethrows an error the thread ran into.
    void WaitIdle() {
        std::unique_lock<std::mutex> lk(cv_mu_);
        idle_.wait(lk, [&] { return !threaded_.load() || (ready_ && !pending_ && !busy_); });
        if (!thread_error_.empty()) throw std::runtime_error("fbus filter thread: " + thread_error_);
    }

    // frame body of FilterThreadFunction (filter.cpp:207-235)
    void ProcessFrame() {
        std::lock_guard<std::mutex> lk(mu_);
        if (det_.empty()) return;
        const double t_det = det_[0].timeStamp;
        size_t n = 0;  // buffered samples not later than the frame; the rest stay buffered (filter.cpp:493-503)
        while (n < buf_.size() && buf_[n].timeStamp <= t_det) ++n;
        std::vector<double> t, d;
        pack(n, t, d);
        if (n == 0) { t.push_back(0.0); d.assign(6, 0.0); }
        const size_t m = det_.size();
        std::vector<int32_t> ids(m);
        std::vector<double> pose(m * 7);
        for (size_t s = 0; s < m; ++s) {
            ids[s] = det_[s].markerID;
            for (int c = 0; c < 3; ++c) pose[s * 7 + c] = det_[s].positionAtCL[c];
            for (int c = 0; c < 4; ++c) pose[s * 7 + 3 + c] = det_[s].quaternionM2CL[c];
        }
        fbus_imu_stream is{n ? n : 1, 1, t.data(), d.data(), FBUS_MEM_HOST, 0};
        fbus_det_frames df{1, m, 1, &t_det, ids.data(), pose.data(), FBUS_MEM_HOST, 0};
        const uint32_t off[2] = {0u, (uint32_t)n};
        check(fbus_step_windows(h_, &is, &df, off, 0, 1, nullptr, FBUS_MEM_HOST));
        check(fbus_synchronize(h_));
        // samples the frame consumed are erased; a frame that could not initialise leaves them (filter.cpp:390,520)
        int32_t inited = 0;
        fbus_state_soa sv{};
        sv.batch = 1;
        sv.initialised = &inited;
        check(fbus_get_state(h_, &sv));
        if (was_init_ || inited) buf_.erase(buf_.begin(), buf_.begin() + n);
        was_init_ = inited != 0;
    }

    // FILTER::GetCameraPose / GetVisualPose (filter.cpp:71-82, 128-139)
    Matrix4d GetCameraPose() { return pose(false); }
    Matrix4d GetVisualPose() { return pose(true); }

    // FILTER::GetStaticMarkerPose / GetDynamicMarkerPose (filter.cpp:88-121): the marker map as 4x4 poses, in the order of
    // the reference's std::map (ascending marker id), appended to `markerPoses`; nothing before the pose is initialised.
    // Both return the same list, as in the reference (the marker states are not part of the shipped filter).
    void GetStaticMarkerPose(std::vector<Matrix4d>& markerPoses) {
        if (initialised()) append_marker_poses(markerPoses);
    }
    void GetDynamicMarkerPose(std::vector<Matrix4d>& markerPoses) {
        if (initialised()) append_marker_poses(markerPoses);
    }
    // FILTER::GetVisualizeInfo (filter.cpp:141-175): everything the Pangolin viewer draws (visualizer.cpp:49)
    void GetVisualizeInfo(Matrix4d& cameraPose, Matrix4d& visualPose, std::vector<Matrix4d>& staticMarkerPoses,
                          std::vector<Matrix4d>& dynamicMarkerPoses) {
        cameraPose = pose(false);
        visualPose = pose(true);
        if (cameraPose[15] != 0) {
            append_marker_poses(staticMarkerPoses);
            append_marker_poses(dynamicMarkerPoses);
        }
    }

    // one row of data/fusion.txt (filter.cpp:241-246): t p(3) q(wxyz) v(3) b_a(3) b_g(3)
    std::array<double, 17> GetFusionRow() {
        std::lock_guard<std::mutex> lk(mu_);
        double t, q[4], p[3], v[3], ba[3], bg[3];
        fbus_state_soa sv{};
        sv.batch = 1;
        sv.t = &t; sv.q = q; sv.p = p; sv.v = v; sv.ba = ba; sv.bg = bg;
        check(fbus_get_state(h_, &sv));
        return {t, p[0], p[1], p[2], q[0], q[1], q[2], q[3], v[0], v[1], v[2], ba[0], ba[1], ba[2], bg[0], bg[1], bg[2]};
    }
    fbus_handle* handle() { return h_; }

  private:
    void check(int rc) {
        if (rc != FBUS_OK) throw std::runtime_error(std::string("fbus: ") + fbus_last_error(h_));
    }
    void pack(size_t n, std::vector<double>& t, std::vector<double>& d) const {
        t.resize(n);
        d.resize(n * 6);
        for (size_t i = 0; i < n; ++i) {
            t[i] = buf_[i].timeStamp;
            for (int c = 0; c < 3; ++c) { d[i * 6 + c] = buf_[i].accel[c]; d[i * 6 + 3 + c] = buf_[i].gyro[c]; }
        }
    }
    bool initialised() {
        std::lock_guard<std::mutex> lk(mu_);
        int32_t inited = 0;
        fbus_state_soa sv{};
        sv.batch = 1;
        sv.initialised = &inited;
        check(fbus_get_state(h_, &sv));
        return inited != 0;
    }
    static void quat_to_rot(const double* q, double* R) {  // Eigen toRotationMatrix
        const double w = q[0], x = q[1], y = q[2], z = q[3];
        R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z);     R[2] = 2 * (x * z + w * y);
        R[3] = 2 * (x * y + w * z);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
        R[6] = 2 * (x * z - w * y);     R[7] = 2 * (y * z + w * x);     R[8] = 1 - 2 * (x * x + y * y);
    }
    void append_marker_poses(std::vector<Matrix4d>& out) const {
        std::vector<std::pair<int, int>> order;  // (id, slot); the first entry of a duplicated id stays, as std::map::insert does (main.cpp:202)
        for (int s = 0; s < cfg_.n_markers; ++s) {
            bool dup = false;
            for (const auto& o : order) dup = dup || o.first == cfg_.marker_id[s];
            if (!dup) order.emplace_back(cfg_.marker_id[s], s);
        }
        std::sort(order.begin(), order.end());
        for (const auto& o : order) {
            const int s = o.second;
            double q[4], R[9];
            fbus_quat_from_rotmat(cfg_.marker_rot + 9 * s, q);  // main.cpp:201 stores Quaterniond(R) ...
            quat_to_rot(q, R);                                  // ... and the getters convert it back
            Matrix4d T{};
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j]; T[i * 4 + 3] = cfg_.marker_pos[3 * s + i]; }
            T[15] = 1;
            out.push_back(T);
        }
    }
    Matrix4d pose(bool visual) {
        std::lock_guard<std::mutex> lk(mu_);
        Matrix4d T{};
        int32_t inited = 0;
        double R[9], p[3], pv[3], qv[4];
        fbus_state_soa sv{};
        sv.batch = 1;
        sv.R = R; sv.p = p; sv.pv = pv; sv.qv = qv; sv.initialised = &inited;
        check(fbus_get_state(h_, &sv));
        if (!inited) return T;  // zero matrix until the pose is initialised, as the reference
        if (visual) {  // quaternionOnlyVisual.toRotationMatrix()
            const double w = qv[0], x = qv[1], y = qv[2], z = qv[3];
            const double Rv[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                                  2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                                  2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)};
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = Rv[i * 3 + j]; T[i * 4 + 3] = pv[i]; }
        } else {  // the CARRIED rotmatI2G, as GetCameraPose does
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[i * 3 + j]; T[i * 4 + 3] = p[i]; }
        }
        T[15] = 1;
        return T;
    }

    fbus_handle* h_ = nullptr;
    bool iir_;
    fbus_config cfg_;
    bool was_init_ = false;
    std::mutex mu_;  // buffers + handle (the reference's imuMutex / imgMutex / visualMutex in one)
    // filter thread (the reference's ekfMutex + ekfCondVar, with a pending flag instead of a bare wait)
    std::thread thread_;
    std::atomic<bool> threaded_{false};
    std::mutex cv_mu_;
    std::condition_variable cv_, idle_;
    bool pending_ = false, busy_ = false, ready_ = false, stop_ = false;
    std::string thread_error_;
    std::vector<IMUData> buf_;
    DetectionResultList det_;
};

}  // namespace FBUSB200
#endif
