/*
 * fbus_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * A dense, scalar C++ restatement of the reference's filter-and-refraction hot path, following
 * C++/src/filter.cpp and C++/src/vision.cpp of CASIA-RoboticFish/FBUS-EKF operation by operation.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (fbus_ekf_b200/) never links, imports or calls it.
 *
 * Parity status:
 *   - F1-F6 (EKF chain): PINNED to the reference's own code.  oracle/_ref is C++/src/filter.cpp compiled
 *     UNMODIFIED from /root/reference against stand-in third-party headers (oracle/ref_build/: the image
 *     has no Eigen / OpenCV / ArUco / yaml-cpp / glog) and executed here; tests/test_oracle_vs_ref.py
 *     compares this restatement with it per call on random states (1e-13 relative) and over both bundled
 *     log replays, reset frames included.  fusion.txt stays a shape-only check (older revision).
 *   - R1+R2 (refractive triangulation + marker pose): PINNED by waterdata/dataset-06
 *     corners.txt -> image.txt (1062 rows) to the 6-significant-digit precision of the logs;
 *   - R2 (marker pose): PINNED by landdata/dataset-02 corners.txt -> image.txt (1257 rows);
 *   - R3 (Gauss-Newton refinement): "parity unpinned" -- it does not exist in the reference.
 *   vision.cpp itself is not compiled (it needs ArUco's patched tracker and a dozen OpenCV calls).
 * The third-party arithmetic on the path is Eigen3 (un-vendored, version unpinned; Ubuntu 18.04
 * ships 3.3.4): quaternion product / normalize / toRotationMatrix / Quaterniond(AngleAxisd) /
 * Quaterniond(Matrix3d), AngleAxisd::matrix, LDLT::solve, EigenSolver<Matrix3d>, determinant.
 * Their published algorithms are restated below.
 */
#ifndef FBUS_ORACLE_H
#define FBUS_ORACLE_H

#include "../include/fbus_ekf.h" /* fbus_config, fbus_imu_stream, fbus_det_frames, fbus_state_soa */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_handle orc_handle;

orc_handle* orc_create(const fbus_config* cfg, size_t batch);
void orc_destroy(orc_handle* h);

int orc_init_gravity_gyrobias(orc_handle* h, const fbus_imu_stream* imu, size_t first, size_t count);
int orc_init_position_quaternion(orc_handle* h, const fbus_det_frames* det, size_t frame, size_t n_imu_before);
int orc_propagate(orc_handle* h, const fbus_imu_stream* imu, size_t first, size_t count, double t_end);
int orc_reset_state(orc_handle* h, const fbus_det_frames* det, size_t frame);
int orc_update(orc_handle* h, const fbus_det_frames* det, size_t frame);
/* n_threads independent copies of the scalar filter, one filter range per thread */
int orc_step_windows(orc_handle* h, const fbus_imu_stream* imu, const fbus_det_frames* det,
                     const uint32_t* win_off, size_t w0, size_t w1, double* trace, int n_threads);

int orc_refract_solve(const fbus_config* cfg, const float* corners, size_t n, double* pose,
                      double* corners3d, int32_t* valid, int n_threads);
int orc_inair_solve(const fbus_config* cfg, const float* corners, size_t n, double* pose, double* corners3d, int32_t* valid);
int orc_undistort_fisheye(const fbus_config* cfg, const float* pixels, size_t n, float* out);
int orc_marker_pose(const fbus_config* cfg, const double* corners3d, size_t n, double* pose);

int orc_get_state(orc_handle* h, fbus_state_soa* out);
int orc_set_state(orc_handle* h, const fbus_state_soa* in);
int orc_stats(orc_handle* h, const double* truth_p, const double* truth_q, double* out);

#ifdef __cplusplus
}
#endif
#endif
