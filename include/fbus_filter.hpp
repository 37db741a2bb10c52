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

#include <array>
#include <cstdint>
#include <mutex>
#include <stdexcept>
#include <string>
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

    explicit FILTER(const fbus_config& cfg, int device = 0, bool iir = true) : iir_(iir) {
        if (fbus_create(&h_, &cfg, device, 1) != FBUS_OK) throw std::runtime_error(std::string("fbus_create: ") + fbus_last_error(nullptr));
    }
    ~FILTER() { fbus_destroy(h_); }
    FILTER(const FILTER&) = delete;
    FILTER& operator=(const FILTER&) = delete;

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
    // the reference notifies the filter thread here (filter.hpp:166-170); the frame body (init | reset -> propagate ->
    // update, filter.cpp:207-235) runs on the GPU before this call returns
    void SetDetectionResultUpdated() {
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
    bool was_init_ = false;
    std::mutex mu_;
    std::vector<IMUData> buf_;
    DetectionResultList det_;
};

}  // namespace FBUSB200
#endif
