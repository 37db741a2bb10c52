// fbus_host_consts.hpp -- host-side derivation of the per-run constants (DevConsts) from fbus_config,
// and the default configuration of the reference's bundled logs.  Plain C++ (no CUDA).
//
// Mirrors what FILTER::FILTER (C++/include/filter.hpp:63-137), the per-call prologues of
// filter.cpp:369-372/411-414/629-632, main.cpp:192-203 (marker rotation -> quaternion) and
// vision.cpp:476-481 compute from the YAML values.
#pragma once

#include <string.h>

#include "../../include/fbus_ekf.h"
#include "fbus_math.cuh"
#include "fbus_refract.cuh"

namespace fbus {

inline void host_quat_left(const double* q, double* m) {  // matrix_math.hpp:38-62
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double t[16] = {w, -x, -y, -z, x, w, -z, y, y, z, w, -x, z, -y, x, w};
    memcpy(m, t, sizeof t);
}
inline void host_quat_right(const double* q, double* m) {  // matrix_math.hpp:64-88
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double t[16] = {w, -x, -y, -z, x, w, z, -y, y, -z, w, x, z, y, -x, w};
    memcpy(m, t, sizeof t);
}

// matlab/rotmat_to_quaternion.m:23-43: the unit eigenvector of the largest eigenvalue of the symmetric 4x4 matrix K(R^T)
// (the best-fit unit quaternion of a not exactly orthonormal R), by cyclic Jacobi rotations.  MATLAB leaves the sign to
// LAPACK; no filter output except the sign of the state quaternion depends on it (see oracle/fbus_oracle_matlab.py): here the
// first component that is not ~0 is made positive.
inline void rotmat_to_quat_eig(const double* Rin, double* q) {
    double R[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i * 3 + j] = Rin[j * 3 + i];  // R = R'
    double K[16];
    K[0] = R[0] - R[4] - R[8];
    K[1] = K[4] = R[3] + R[1];
    K[2] = K[8] = R[6] + R[2];
    K[3] = K[12] = R[5] - R[7];
    K[5] = R[4] - R[0] - R[8];
    K[6] = K[9] = R[7] + R[5];
    K[7] = K[13] = R[6] - R[2];
    K[10] = R[8] - R[0] - R[4];
    K[11] = K[14] = R[1] - R[3];
    K[15] = R[0] + R[4] + R[8];
    for (int i = 0; i < 16; ++i) K[i] /= 3.0;
    double V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 4; ++p)
            for (int r = p + 1; r < 4; ++r) off += K[p * 4 + r] * K[p * 4 + r];
        if (off < 1e-34) break;
        for (int p = 0; p < 4; ++p)
            for (int r = p + 1; r < 4; ++r) {
                const double apq = K[p * 4 + r];
                if (apq == 0.0) continue;
                const double th = (K[r * 4 + r] - K[p * 4 + p]) / (2.0 * apq);
                const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int x = 0; x < 4; ++x) {  // K <- K J
                    const double kp = K[x * 4 + p], kr = K[x * 4 + r];
                    K[x * 4 + p] = cs * kp - sn * kr;
                    K[x * 4 + r] = sn * kp + cs * kr;
                }
                for (int x = 0; x < 4; ++x) {  // K <- J^T K
                    const double kp = K[p * 4 + x], kr = K[r * 4 + x];
                    K[p * 4 + x] = cs * kp - sn * kr;
                    K[r * 4 + x] = sn * kp + cs * kr;
                }
                for (int x = 0; x < 4; ++x) {  // V <- V J
                    const double vp = V[x * 4 + p], vr = V[x * 4 + r];
                    V[x * 4 + p] = cs * vp - sn * vr;
                    V[x * 4 + r] = sn * vp + cs * vr;
                }
            }
    }
    int best = 0;
    for (int i = 1; i < 4; ++i)
        if (K[i * 4 + i] > K[best * 4 + best]) best = i;
    const double v[4] = {V[0 * 4 + best], V[1 * 4 + best], V[2 * 4 + best], V[3 * 4 + best]};
    double nrm = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
    q[0] = v[3] / nrm; q[1] = v[0] / nrm; q[2] = v[1] / nrm; q[3] = v[2] / nrm;  // q = [V(4); V(1); V(2); V(3)]
    for (int i = 0; i < 4; ++i)
        if (fabs(q[i]) > 1e-12) {
            if (q[i] < 0)
                for (int j = 0; j < 4; ++j) q[j] = -q[j];
            break;
        }
}

inline int make_dev_consts(const fbus_config* c, DevConsts* k, MarkerTable* tab) {
    if (c->n_markers < 0 || c->n_markers > MAXM) return FBUS_E_BADARG;
    memset(k, 0, sizeof *k);
    memset(tab, 0, sizeof *tab);
    const double flip[3] = {-1.0, -1.0, 1.0};  // T_C_I, filter.hpp:67-69
    double P_LI[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) k->R_IL[i * 3 + j] = flip[i] * c->tsc_left[i * 4 + j];
        P_LI[i] = flip[i] * c->tsc_left[i * 4 + 3];
    }
    const bool matlab = (c->flags & FBUS_FLAG_MATLAB) != 0;
    if (matlab) rotmat_to_quat_eig(k->R_IL, k->Q_IL);  // rotmat_to_quaternion.m (unit quaternion)
    else R2q(k->R_IL, k->Q_IL);  // Quaterniond(R_I_L), NOT normalised (filter.cpp:370)
    for (int i = 0; i < 3; ++i)  // P_I_L = -R_I_L^T * P_L_I (filter.cpp:372)
        k->P_IL[i] = -(k->R_IL[i] * P_LI[0] + k->R_IL[3 + i] * P_LI[1] + k->R_IL[6 + i] * P_LI[2]);
    k->Qd[0] = c->accel_n_cov; k->Qd[1] = c->gyro_n_cov; k->Qd[2] = c->accel_b_cov; k->Qd[3] = c->gyro_b_cov;
    k->Rp = c->pos_n_cov; k->Rq = c->quat_n_cov;
    k->max_dist = c->marker_max_dist; k->switch_thres = c->marker_switch_thres; k->reset_gap = c->reset_gap;
    // vision view (raw T_SC): R_R_L = R_I_L R_I_R^T ; P_L_R = P_L_I - R_R_L P_R_I (vision.cpp:476-481)
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int x = 0; x < 3; ++x) s += c->tsc_left[i * 4 + x] * c->tsc_right[j * 4 + x];
            k->R_RL[i * 3 + j] = s;
        }
    for (int i = 0; i < 3; ++i) {
        const double t = k->R_RL[i * 3] * c->tsc_right[3] + k->R_RL[i * 3 + 1] * c->tsc_right[7] + k->R_RL[i * 3 + 2] * c->tsc_right[11];
        k->P_LR[i] = c->tsc_left[i * 4 + 3] - t;
    }
    // in-air stereo (vision.cpp:402-408): T_L_R = [R_IR R_IL^T | P_LI - (R_IR R_IL^T) P_RI] with the raw T_SC
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double s = 0.0;
            for (int x = 0; x < 3; ++x) s += c->tsc_right[i * 4 + x] * c->tsc_left[j * 4 + x];
            k->T_LR_air[i * 4 + j] = s;
        }
    }
    for (int i = 0; i < 3; ++i) {
        const double t = k->T_LR_air[i * 4] * c->tsc_right[3] + k->T_LR_air[i * 4 + 1] * c->tsc_right[7] + k->T_LR_air[i * 4 + 2] * c->tsc_right[11];
        k->T_LR_air[i * 4 + 3] = c->tsc_left[i * 4 + 3] - t;
    }
    for (int cam = 0; cam < 2; ++cam)
        for (int i = 0; i < 4; ++i) { k->cam_k[cam][i] = c->cam_k[cam][i]; k->cam_d[cam][i] = c->cam_d[cam][i]; }
    k->a0 = c->n_air / c->n_glass;
    k->a1 = c->n_glass / c->n_water;
    k->air_lt_glass = c->n_air < c->n_glass;      // vision.cpp:511
    k->glass_gt_water = c->n_glass > c->n_water;  // vision.cpp:531
    k->d_air = c->d_air; k->d_glass = c->d_glass;
    for (int i = 0; i < 3; ++i) k->normal[i] = c->normal[i];
    k->dect_thres = c->marker_dect_dist_thres;
    k->rod_s = sin(-3.1415926 / 4);
    k->rod_c = cos(-3.1415926 / 4);
    k->n_markers = c->n_markers;
    k->flags = c->flags;
    k->imu_g = c->imu_g;
    double Lil[16];
    host_quat_left(k->Q_IL, Lil);
    for (int m = 0; m < c->n_markers; ++m) {
        MarkerConst& mk = tab->mk[m];
        mk.id = c->marker_id[m];
        for (int i = 0; i < 3; ++i) mk.p[i] = c->marker_pos[m * 3 + i];
        if (matlab) rotmat_to_quat_eig(&c->marker_rot[m * 9], mk.q);
        else R2q(&c->marker_rot[m * 9], mk.q);  // main.cpp:201
        double Rq[16], A1[16];
        host_quat_right(mk.q, Rq);
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                double s = 0.0;
                for (int x = 0; x < 4; ++x) s += Rq[i * 4 + x] * Lil[x * 4 + j];
                A1[i * 4 + j] = s;
            }
        for (int i = 0; i < 4; ++i)  // * L2 = diag(1,-1,-1,-1)
            for (int j = 0; j < 4; ++j) mk.CM[i * 4 + j] = (j == 0) ? A1[i * 4 + j] : -A1[i * 4 + j];
    }
    return FBUS_OK;
}

// constants of the Gauss-Newton refinement (R3)
inline void make_gn_consts(const fbus_config* c, const DevConsts* k, GnConsts* g) {
    g->d0 = c->d_air; g->d1 = c->d_glass;
    g->k1 = c->n_air / c->n_glass; g->k2 = c->n_air / c->n_water;
    g->size = c->marker_size;
    g->tol = c->gn_tol;
    for (int i = 0; i < 3; ++i) g->P_LR[i] = k->P_LR[i];
    const double* m = k->R_RL;
    const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    const double id = 1.0 / det;
    g->R_RL_inv[0] = (m[4] * m[8] - m[5] * m[7]) * id; g->R_RL_inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; g->R_RL_inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
    g->R_RL_inv[3] = (m[5] * m[6] - m[3] * m[8]) * id; g->R_RL_inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; g->R_RL_inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
    g->R_RL_inv[6] = (m[3] * m[7] - m[4] * m[6]) * id; g->R_RL_inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; g->R_RL_inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
}

inline void config_default(fbus_config* c) {
    memset(c, 0, sizeof *c);
    // C++/config/camerainfo1.yml (== matlab/config/camerainfo.yml): the calibration of the bundled logs
    const double tl[16] = {-0.999862, 0.015685, -0.00548, 0.059967, -0.015639, -0.999843, -0.00827, 0.000127837,
                           -0.005609, -0.008183, 0.999951, -0.002, 0, 0, 0, 1};
    const double tr[16] = {-0.999826, 0.00929485, -0.0161445, -0.0601272, -0.00937869, -0.999942, 0.00514829, 0.000124714,
                           -0.0160959, 0.00529897, 0.999857, -0.002, 0, 0, 0, 1};
    memcpy(c->tsc_left, tl, sizeof tl);
    memcpy(c->tsc_right, tr, sizeof tr);
    // C++/config/paramconfig.yml:44-57
    c->accel_n_cov = 0.001; c->gyro_n_cov = 0.0001; c->accel_b_cov = 0.001; c->gyro_b_cov = 0.0001;
    c->pos_n_cov = 0.001; c->quat_n_cov = 0.001;
    c->marker_max_dist = 2.0; c->marker_switch_thres = 0.5;
    // filter.hpp:29-34
    const double p0[6] = {0.0001, 0.01, 0.0001, 1e-2, 1e-2, 100.0};
    memcpy(c->p0_diag, p0, sizeof p0);
    c->reset_gap = 0.1;  // filter.cpp:462
    // paramconfig.yml:30-42
    c->n_air = 1.00; c->n_water = 1.32; c->n_glass = 1.49; c->d_air = 0.002; c->d_glass = 0.02;
    c->normal[0] = 0; c->normal[1] = 0; c->normal[2] = 1;
    c->marker_dect_dist_thres = 2.0;  // paramconfig.yml:27
    // camerainfo1.yml K / D of both cameras
    const double kk[2][4] = {{246.134, 246.265, 325.504, 178.694}, {245.124, 244.704, 341.197, 179.214}};
    const double dd[2][4] = {{0.584804, 0.158016, -0.5657, 0.272636}, {0.590953, 0.140311, -0.490475, 0.206821}};
    memcpy(c->cam_k, kk, sizeof kk);
    memcpy(c->cam_d, dd, sizeof dd);
    c->marker_size = 0.28;            // vision.hpp:114 (paramconfig.yml:23 says 0.48 but is unused)
    // C++/config/markersetup.yml
    static const int ids[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 17, 18};
    static const double pos[12][3] = {{0, 0, 0}, {0, 0.61, 0.285}, {0, 0.61, 1.185}, {0, 0.61, 2.085}, {0, 0.61, 2.985},
                                      {0, 0.265, 4.12}, {0, -0.635, 4.12}, {0, -1.535, 4.12}, {0, -2.435, 4.12},
                                      {0, -2.7, 0}, {0, -1.8, 0}, {0, -0.9, 0}};
    static const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    static const double Rx90[9] = {1, 0, 0, 0, 0, -1, 0, 1, 0};
    static const double Rx180[9] = {1, 0, 0, 0, -1, 0, 0, 0, -1};
    c->n_markers = 12;
    for (int m = 0; m < 12; ++m) {
        c->marker_id[m] = ids[m];
        for (int i = 0; i < 3; ++i) c->marker_pos[m * 3 + i] = pos[m][i];
        const double* R = (m == 0 || m >= 9) ? I3 : (m <= 4 ? Rx90 : Rx180);
        memcpy(&c->marker_rot[m * 9], R, sizeof I3);
    }
    c->flags = 0;
    c->imu_g = 9.802;  // camerainfo1.yml "g" (IMUInfo.g, common.hpp:148)
}

// the constants of matlab/FBUS_EKF.m:28-41,83-112 on top of the defaults: P0, measurement noise, FBUS_FLAG_MATLAB
inline void config_matlab(fbus_config* c) {
    config_default(c);
    const double p0[6] = {0.0001, 0.1, 0.0001, 0.001, 0.001, 100.0};  // FBUS_EKF.m:86-98
    memcpy(c->p0_diag, p0, sizeof p0);
    c->pos_n_cov = 0.01; c->quat_n_cov = 0.01;                        // FBUS_EKF.m:32-33
    // FBUS_EKF.m:36-39 and 101-105: systemNoise = diag(1e-3 I, 1e-4 I, 1e-3 I, 1e-4 I) on the v, theta, b_a, b_g rows -- the
    // values of paramconfig.yml, i.e. the defaults
    c->flags = FBUS_FLAG_MATLAB;
}

}  // namespace fbus
