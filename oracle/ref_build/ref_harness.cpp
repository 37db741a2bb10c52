// ref_harness.cpp -- drives the REFERENCE'S OWN filter.cpp (compiled unmodified from /root/reference/C++/src/filter.cpp
// against the stand-in headers in stubs/) through the same C structs the oracle and the product use.
// TEST INFRASTRUCTURE, NOT PRODUCT: only tests/ and bench.py's cpu_baseline / --impl reference legs load the resulting
// oracle/_ref/libfbus_ref.so.  Nothing here re-implements filter arithmetic: every numeric statement that runs is in
// filter.cpp / filter.hpp / common.hpp / matrix_math.hpp of the reference; this file only moves data in and out of
// FBUSEKF::FILTER objects and calls the reference's methods in the order FILTER::FilterThreadFunction does
// (filter.cpp:207-235).
//
// The methods on the path are private (filter.hpp:187-212).  filter.cpp is compiled as is; THIS translation unit includes
// filter.hpp with `private` spelled `public`, which changes neither the layout of the class nor the mangled names of its
// members (access is not part of either in the Itanium ABI), so the calls below bind to the unmodified object code.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>
#include <sys/stat.h>

#define private public
#include "filter.hpp"
#undef private

#include "../../include/fbus_ekf.h"

namespace {

using FBUSEKF::FILTER;

FBUSEKF::DetectionResultList gather(const fbus_det_frames* det, size_t frame, size_t b) {
    FBUSEKF::DetectionResultList out;
    const size_t B = det->batch, m = det->max_markers;
    for (size_t s = 0; s < m; ++s) {
        const int id = det->id[(frame * m + s) * B + b];
        if (id < 0) continue;
        const double* base = det->pose + ((frame * m + s) * 7) * B + b;
        FBUSEKF::DetectionResult d;  // as VISION::VisionThreadFunction fills it (vision.cpp:96-101)
        d.markerID = id;
        d.timeStamp = det->t[frame];
        d.positionAtCL = Eigen::Vector3d(base[0], base[B], base[2 * B]);
        d.quaternionM2CL = Eigen::Quaterniond(base[3 * B], base[4 * B], base[5 * B], base[6 * B]);
        d.rotmatM2L = d.quaternionM2CL.toRotationMatrix();
        out.push_back(d);
    }
    return out;
}

FBUSEKF::IMUData sample(const fbus_imu_stream* imu, size_t i, size_t b) {
    const size_t B = imu->batch;
    const double* d = imu->data + i * 6 * B + b;
    FBUSEKF::IMUData s;
    s.timeStamp = imu->t[i];
    s.accel = Eigen::Vector3d(d[0], d[B], d[2 * B]);
    s.gyro = Eigen::Vector3d(d[3 * B], d[4 * B], d[5 * B]);
    return s;
}

template <class Fn>
void parallel_for(size_t n, int n_threads, Fn fn) {
    if (n_threads <= 1 || n < 2) { fn(0, n); return; }
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    std::vector<int> cpus;
    if (sched_getaffinity(0, sizeof allowed, &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c)
            if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
    std::vector<std::thread> th;
    const size_t T = (size_t)n_threads;
    for (size_t i = 0; i < T; ++i) {
        const size_t lo = n * i / T, hi = n * (i + 1) / T;
        if (lo >= hi) continue;
        th.emplace_back([=] { fn(lo, hi); });
        if (!cpus.empty()) {  // one thread per allowed core, 1:1
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[i % cpus.size()], &one);
            pthread_setaffinity_np(th.back().native_handle(), sizeof one, &one);
        }
    }
    for (auto& x : th) x.join();
}

}  // namespace

struct ref_handle {
    std::vector<FILTER*> f;
    std::vector<int> status;
    std::vector<size_t> pushed;  // per filter: IMU samples [.., pushed) have been handed to the filter's buffer
    const void* last_imu_data = nullptr;
    size_t last_n_samples = 0, last_w1 = 0;
    ~ref_handle() { for (FILTER* p : f) delete p; }
};

extern "C" {

int ref_abi_version(void) { return FBUS_ABI_VERSION; }

// FILTER::FILTER(leftCamera, rightCamera, imu, ekfParam, markerPoseServer) exactly as main.cpp:170-207 builds its arguments
ref_handle* ref_create(const fbus_config* cfg, size_t batch) {
    // the reference hard-codes what fbus_config makes configurable: refuse anything else rather than emulate it
    const double p0_ref[6] = {POSITION_CONV, VELOCITY_CONV, QUATERNION_CONV, ACCEL_BIAS_CONV, GYRO_BIAS_CONV, GRAVITY_CONV};
    for (int i = 0; i < 6; ++i)
        if (cfg->p0_diag[i] != p0_ref[i]) return nullptr;
    if (cfg->reset_gap != 0.1 || (cfg->flags & FBUS_FLAG_JOSEPH)) return nullptr;
    FBUSEKF::CameraInfo left, right;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            left.T_SC(i, j) = cfg->tsc_left[i * 4 + j];
            right.T_SC(i, j) = cfg->tsc_right[i * 4 + j];
        }
    FBUSEKF::IMUInfo imuInfo;
    imuInfo.g = cfg->imu_g;
    FBUSEKF::EkfParam ekf;
    ekf.accelNoiseCov = cfg->accel_n_cov;
    ekf.gyroNoiseCov = cfg->gyro_n_cov;
    ekf.accelBiasCov = cfg->accel_b_cov;
    ekf.gyroBiasCov = cfg->gyro_b_cov;
    ekf.posNoiseCov = cfg->pos_n_cov;
    ekf.quatNoiseCov = cfg->quat_n_cov;
    ekf.fcpMarkerMaxDist = cfg->marker_max_dist;
    ekf.fcpMarkerSwitchThres = cfg->marker_switch_thres;
    FBUSEKF::MarkerPoseServer server;
    for (int m = 0; m < cfg->n_markers; ++m) {
        FBUSEKF::MarkerPose mp;  // main.cpp:196-203
        mp.markerID = cfg->marker_id[m];
        mp.positionAtG = Eigen::Vector3d(cfg->marker_pos[m * 3], cfg->marker_pos[m * 3 + 1], cfg->marker_pos[m * 3 + 2]);
        Eigen::Matrix3d rot;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) rot(i, j) = cfg->marker_rot[m * 9 + i * 3 + j];
        mp.quaternionM2G = Eigen::Quaterniond(rot);
        server.insert(std::pair<FBUSEKF::MarkerID, FBUSEKF::MarkerPose>(mp.markerID, mp));
    }
    ref_handle* h = new ref_handle;
    h->f.reserve(batch);
    for (size_t b = 0; b < batch; ++b) {
        FILTER* f = new FILTER(left, right, imuInfo, ekf, server);
        f->isImuDataUpdated_ = false;  // left uninitialised by the reference's constructor; never read on this path
        f->isDetectionResultUpdated_ = false;
        h->f.push_back(f);
    }
    h->status.assign(batch, 0);
    h->pushed.assign(batch, 0);
    return h;
}
void ref_destroy(ref_handle* h) { delete h; }

// FILTER::InitializeGravityAndBias (filter.cpp:256-285) on samples [first, first+count)
int ref_init_gravity_gyrobias(ref_handle* h, const fbus_imu_stream* imu, size_t first, size_t count) {
    const size_t B = imu->batch;
    if (B != h->f.size() || first + count > imu->n_samples) return FBUS_E_BADARG;
    for (size_t b = 0; b < B; ++b) {
        FILTER& f = *h->f[b];
        f.imuMeasuementBuffer_.clear();
        // filter.cpp:261-262 read imuMeasuementBuffer_.end()->timeStamp, one element past the last sample: give that read a
        // defined home (a constructed element inside the capacity, popped again) and restore the timestamps afterwards --
        // the reference never uses them before InitializePose overwrites both
        f.imuMeasuementBuffer_.reserve(count + 1);
        for (size_t i = first; i < first + count; ++i) f.imuMeasuementBuffer_.push_back(sample(imu, i, b));
        f.imuMeasuementBuffer_.push_back(FBUSEKF::IMUData());
        f.imuMeasuementBuffer_.pop_back();
        const double t_nom = f.sysNominalState_.timeStamp, t_err = f.sysErrorState_.timeStamp;
        f.InitializeGravityAndBias();
        f.sysNominalState_.timeStamp = t_nom;
        f.sysErrorState_.timeStamp = t_err;
    }
    return FBUS_OK;
}

// FILTER::InitializePose (filter.cpp:291-399) with n_imu_before buffered samples not later than the frame
int ref_init_position_quaternion(ref_handle* h, const fbus_det_frames* det, size_t frame, size_t n_imu_before) {
    if (det->batch != h->f.size() || frame >= det->n_frames) return FBUS_E_BADARG;
    for (size_t b = 0; b < h->f.size(); ++b) {
        FILTER& f = *h->f[b];
        f.detectionResult_ = gather(det, frame, b);
        f.imuMeasuementBuffer_.clear();
        FBUSEKF::IMUData s;
        s.timeStamp = det->t[frame];
        for (size_t i = 0; i < n_imu_before; ++i) f.imuMeasuementBuffer_.push_back(s);
        if (f.InitializePose()) f.isInitializePose_ = true;  // filter.cpp:209-211
        else h->status[b] |= FBUS_ST_INIT_FAILED;
        f.imuMeasuementBuffer_.clear();
    }
    return FBUS_OK;
}

// FILTER::BatchImuProcessing (filter.cpp:483-531) over samples [first, first+count) with endTime = t_end
int ref_propagate(ref_handle* h, const fbus_imu_stream* imu, size_t first, size_t count, double t_end) {
    if (imu->batch != h->f.size() || first + count > imu->n_samples) return FBUS_E_BADARG;
    for (size_t b = 0; b < h->f.size(); ++b) {
        FILTER& f = *h->f[b];
        FBUSEKF::DetectionResult d;
        d.timeStamp = t_end;  // endTime = detectionResult_[0].timeStamp, filter.cpp:491
        f.detectionResult_.assign(1, d);
        f.imuMeasuementBuffer_.clear();
        for (size_t i = first; i < first + count; ++i) f.imuMeasuementBuffer_.push_back(sample(imu, i, b));
        f.BatchImuProcessing();
        f.imuMeasuementBuffer_.clear();
    }
    return FBUS_OK;
}

int ref_reset_state(ref_handle* h, const fbus_det_frames* det, size_t frame) {
    if (det->batch != h->f.size() || frame >= det->n_frames) return FBUS_E_BADARG;
    for (size_t b = 0; b < h->f.size(); ++b) {
        FILTER& f = *h->f[b];
        f.detectionResult_ = gather(det, frame, b);
        f.ResetSystemState();
    }
    return FBUS_OK;
}

int ref_update(ref_handle* h, const fbus_det_frames* det, size_t frame) {
    if (det->batch != h->f.size() || frame >= det->n_frames) return FBUS_E_BADARG;
    for (size_t b = 0; b < h->f.size(); ++b) {
        FILTER& f = *h->f[b];
        f.detectionResult_ = gather(det, frame, b);
        if (f.detectionResult_.empty()) h->status[b] |= FBUS_ST_NO_DETECTION;
        f.ObservationUpdate();
    }
    return FBUS_OK;
}

// The loop body of FILTER::FilterThreadFunction (filter.cpp:199-249) for frames [w0, w1).  The IMU samples
// [win_off[w], win_off[w+1]) arrive in the filter's own imuMeasuementBuffer_ before frame w's detections do; the
// reference's methods erase what they consume (filter.cpp:390,520), everything else stays buffered.  Samples are pushed
// as they are (the callers pre-filter streams themselves; FILTER::SetImuData's IIR is exercised by ref_set_imu_data).
int ref_step_windows(ref_handle* h, const fbus_imu_stream* imu, const fbus_det_frames* det, const uint32_t* win_off,
                     size_t w0, size_t w1, double* trace, int n_threads) {
    const size_t B = h->f.size();
    if (imu->batch != B || det->batch != B || w1 > det->n_frames || w0 > w1) return FBUS_E_BADARG;
    const bool resume = w0 > 0 && w0 == h->last_w1 && imu->data == h->last_imu_data && imu->n_samples == h->last_n_samples;
    h->last_imu_data = imu->data;
    h->last_n_samples = imu->n_samples;
    h->last_w1 = w1;
    parallel_for(B, n_threads, [&](size_t lo, size_t hi) {
        for (size_t b = lo; b < hi; ++b) {
            FILTER& f = *h->f[b];
            if (!resume) { f.imuMeasuementBuffer_.clear(); h->pushed[b] = win_off[w0]; }
            for (size_t w = w0; w < w1; ++w) {
                for (size_t i = h->pushed[b]; i < win_off[w + 1]; ++i) f.imuMeasuementBuffer_.push_back(sample(imu, i, b));
                h->pushed[b] = std::max<size_t>(h->pushed[b], win_off[w + 1]);
                f.detectionResult_ = gather(det, w, b);
                if (f.detectionResult_.empty()) {
                    // the filter thread is only woken by a non-empty detection list (vision.cpp:136-140)
                    h->status[b] |= FBUS_ST_NO_DETECTION;
                } else if (!f.isInitializePose_) {  // filter.cpp:207-226
                    if (f.InitializePose()) f.isInitializePose_ = true;
                    else h->status[b] |= FBUS_ST_INIT_FAILED;
                } else {  // filter.cpp:229-235
                    f.ResetSystemState();
                    f.BatchImuProcessing();
                    f.ObservationUpdate();
                }
                if (trace) {  // the data/fusion.txt row, filter.cpp:241-246
                    double* row = trace + ((w - w0) * 17) * B + b;
                    const FBUSEKF::NominalState& n = f.sysNominalState_;
                    row[0] = n.timeStamp;
                    for (int c = 0; c < 3; ++c) row[(1 + c) * B] = n.positionAtG[c];
                    row[4 * B] = n.quaternionI2G.w(); row[5 * B] = n.quaternionI2G.x();
                    row[6 * B] = n.quaternionI2G.y(); row[7 * B] = n.quaternionI2G.z();
                    for (int c = 0; c < 3; ++c) {
                        row[(8 + c) * B] = n.velocityAtG[c];
                        row[(11 + c) * B] = n.accelBias[c];
                        row[(14 + c) * B] = n.gyroBias[c];
                    }
                }
            }
        }
    });
    return FBUS_OK;
}

// FILTER::SetImuData (filter.cpp:24-55) for filter 0 of a stream: feeds samples [first, first+count) one by one into an
// EMPTY buffer and returns what the buffer holds afterwards -- the 1-pole IIR (filter.cpp:42-47) and the
// 2000-sample cap that erases the oldest 500 (filter.cpp:50-54).  out_t [cap], out [cap][6]; *n_out = buffer size.
int ref_set_imu_data(ref_handle* h, const fbus_imu_stream* imu, size_t first, size_t count, double* out_t, double* out,
                     size_t cap, size_t* n_out) {
    if (h->f.empty() || first + count > imu->n_samples) return FBUS_E_BADARG;
    struct stat st;
    if (stat("data", &st) == 0) return FBUS_E_STATE;  // OPEN_DATA_RECORDING would append to data/imu.txt (filter.cpp:28-34)
    FILTER& f = *h->f[0];
    f.imuMeasuementBuffer_.clear();
    f.imuMeasuementBuffer_.reserve(IMU_BUFFER_MAX_SIZE + 2);
    for (size_t i = first; i < first + count; ++i) f.SetImuData(sample(imu, i, 0));
    const size_t n = f.imuMeasuementBuffer_.size();
    *n_out = n;
    for (size_t i = 0; i < n && i < cap; ++i) {
        out_t[i] = f.imuMeasuementBuffer_[i].timeStamp;
        for (int c = 0; c < 3; ++c) {
            out[i * 6 + c] = f.imuMeasuementBuffer_[i].accel[c];
            out[i * 6 + 3 + c] = f.imuMeasuementBuffer_[i].gyro[c];
        }
    }
    f.imuMeasuementBuffer_.clear();
    return FBUS_OK;
}

// FILTER::GetCameraPose / GetVisualPose (filter.cpp:71-82,128-139): the 4x4 poses the viewer reads, row-major [16] each
int ref_get_poses(ref_handle* h, size_t b, double* camera_pose, double* visual_pose) {
    if (b >= h->f.size()) return FBUS_E_BADARG;
    const Eigen::Matrix4d c = h->f[b]->GetCameraPose(), v = h->f[b]->GetVisualPose();
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { camera_pose[i * 4 + j] = c(i, j); visual_pose[i * 4 + j] = v(i, j); }
    return FBUS_OK;
}

static void copy_state(ref_handle* h, fbus_state_soa* s, bool get) {
    const size_t B = h->f.size();
    for (size_t b = 0; b < B; ++b) {
        FILTER& f = *h->f[b];
        FBUSEKF::NominalState& n = f.sysNominalState_;
        auto vec3 = [&](double* dst, Eigen::Vector3d& v) {
            if (!dst) return;
            for (int c = 0; c < 3; ++c) { if (get) dst[(size_t)c * B + b] = v[c]; else v[c] = dst[(size_t)c * B + b]; }
        };
        auto quat = [&](double* dst, Eigen::Quaterniond& q) {
            if (!dst) return;
            if (get) { dst[b] = q.w(); dst[B + b] = q.x(); dst[2 * B + b] = q.y(); dst[3 * B + b] = q.z(); }
            else q = Eigen::Quaterniond(dst[b], dst[B + b], dst[2 * B + b], dst[3 * B + b]);
        };
        if (s->t) { if (get) s->t[b] = n.timeStamp; else { n.timeStamp = s->t[b]; f.sysErrorState_.timeStamp = s->t[b]; } }
        quat(s->q, n.quaternionI2G);
        if (s->R)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    if (get) s->R[(size_t)(i * 3 + j) * B + b] = n.rotmatI2G(i, j);
                    else n.rotmatI2G(i, j) = s->R[(size_t)(i * 3 + j) * B + b];
                }
        vec3(s->p, n.positionAtG);
        vec3(s->v, n.velocityAtG);
        vec3(s->ba, n.accelBias);
        vec3(s->bg, n.gyroBias);
        vec3(s->g, n.gravityAtG);
        vec3(s->pv, n.positionOnlyVisual);
        quat(s->qv, n.quaternionOnlyVisual);
        if (s->P)
            for (int i = 0; i < 18; ++i)
                for (int j = 0; j < 18; ++j) {
                    if (get) s->P[(size_t)(i * 18 + j) * B + b] = f.sysErrorState_.stateCovariance(i, j);
                    else f.sysErrorState_.stateCovariance(i, j) = s->P[(size_t)(i * 18 + j) * B + b];
                }
        if (s->prev_marker_id) { if (get) s->prev_marker_id[b] = f.preUsedMarkerID_; else f.preUsedMarkerID_ = s->prev_marker_id[b]; }
        if (s->initialised) { if (get) s->initialised[b] = f.isInitializePose_ ? 1 : 0; else f.isInitializePose_ = s->initialised[b] != 0; }
        if (s->status) { if (get) s->status[b] = h->status[b]; else h->status[b] = s->status[b]; }
    }
}

int ref_get_state(ref_handle* h, fbus_state_soa* out) {
    if (out->batch != h->f.size()) return FBUS_E_BADARG;
    copy_state(h, out, true);
    return FBUS_OK;
}
int ref_set_state(ref_handle* h, const fbus_state_soa* in) {
    if (in->batch != h->f.size()) return FBUS_E_BADARG;
    copy_state(h, const_cast<fbus_state_soa*>(in), false);
    return FBUS_OK;
}

}  // extern "C"
