// fbus_oracle.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).  See fbus_oracle.h.
//
// Dense scalar restatement of the reference hot path.  It deliberately keeps the reference's
// evaluation structure (dense 18x18 products with structural zeros, HP computed twice, LDLT solve,
// (I-KH)P, general rotation/quaternion helpers) so that it can serve as the CPU baseline
// ("kind":"port") and as the parity checker.  Citations are relative to /root/reference.
//
// Build: g++ -O2 -ffp-contract=off -std=c++17 -fPIC -shared -pthread (see oracle/Makefile).

#include "fbus_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include <pthread.h>
#include <sched.h>

namespace {

// common.hpp:14 -- the reference overrides M_PI for all of its code (parity trap A.3-1)
constexpr double REF_M_PI = 3.1415926;

// ------------------------------------------------------------------------------------------
// tiny dense helpers (row-major)
// ------------------------------------------------------------------------------------------
template <int R, int K, int C>
inline void matmul(const double* A, const double* B, double* out) {  // out[RxC] = A[RxK] * B[KxC]
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += A[i * K + k] * B[k * C + j];
            out[i * C + j] = s;
        }
}
template <int R, int K, int C>
inline void matmul_bt(const double* A, const double* B, double* out) {  // out[RxC] = A[RxK] * B[CxK]^T
    for (int i = 0; i < R; ++i)
        for (int j = 0; j < C; ++j) {
            double s = 0.0;
            for (int k = 0; k < K; ++k) s += A[i * K + k] * B[j * K + k];
            out[i * C + j] = s;
        }
}
inline void mat3_vec(const double* M, const double* v, double* o) {
    for (int i = 0; i < 3; ++i) o[i] = M[i * 3] * v[0] + M[i * 3 + 1] * v[1] + M[i * 3 + 2] * v[2];
}
inline void mat3_t(const double* M, double* T) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) T[i * 3 + j] = M[j * 3 + i];
}
inline double norm3(const double* v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

// matrix_math.hpp:9-23
inline double Signum(double x) { return x < 0 ? -1.0 : 1.0; }
inline double Absolute(double x) { return x < 0 ? -x : x; }
// matrix_math.hpp:26-36
inline void skew(const double* v, double* m) {
    m[0] = 0; m[1] = -v[2]; m[2] = v[1];
    m[3] = v[2]; m[4] = 0; m[5] = -v[0];
    m[6] = -v[1]; m[7] = v[0]; m[8] = 0;
}
// matrix_math.hpp:38-62 (q = w,x,y,z)
inline void quat_left(const double* q, double* m) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double t[16] = {w, -x, -y, -z, x, w, -z, y, y, z, w, -x, z, -y, x, w};
    std::memcpy(m, t, sizeof t);
}
// matrix_math.hpp:64-88
inline void quat_right(const double* q, double* m) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double t[16] = {w, -x, -y, -z, x, w, z, -y, y, -z, w, x, z, y, -x, w};
    std::memcpy(m, t, sizeof t);
}
// Eigen quaternion product, q = (w,x,y,z)
inline void qmul(const double* a, const double* b, double* o) {
    const double w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
    const double x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
    const double y = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
    const double z = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
    o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
inline void qconj(const double* a, double* o) { o[0] = a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = -a[3]; }
inline void qnormalize(double* q) {
    const double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}
// Eigen Quaternion::toRotationMatrix; applied literally to non-unit q as well
inline void q2R(const double* q, double* R) {
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w;
    const double txx = tx * x, txy = ty * x, txz = tz * x;
    const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
    R[0] = 1 - (tyy + tzz); R[1] = txy - twz;       R[2] = txz + twy;
    R[3] = txy + twz;       R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
    R[6] = txz - twy;       R[7] = tyz + twx;       R[8] = 1 - (txx + tyy);
}
// Eigen Quaterniond(Matrix3d); no normalisation
inline void R2q(const double* m, double* q) {
    double t = m[0] + m[4] + m[8];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q[0] = 0.5 * t;
        t = 0.5 / t;
        q[1] = (m[7] - m[5]) * t;
        q[2] = (m[2] - m[6]) * t;
        q[3] = (m[3] - m[1]) * t;
    } else {
        int i = 0;
        if (m[4] > m[0]) i = 1;
        if (m[8] > m[i * 3 + i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(m[i * 3 + i] - m[j * 3 + j] - m[k * 3 + k] + 1.0);
        q[1 + i] = 0.5 * t;
        t = 0.5 / t;
        q[0] = (m[k * 3 + j] - m[j * 3 + k]) * t;
        q[1 + j] = (m[j * 3 + i] + m[i * 3 + j]) * t;
        q[1 + k] = (m[k * 3 + i] + m[i * 3 + k]) * t;
    }
}
// Eigen AngleAxisd(angle, axis).matrix()
inline void angleaxis_matrix(double angle, const double* u, double* R) {
    const double s = std::sin(angle), c = std::cos(angle);
    const double sa[3] = {s * u[0], s * u[1], s * u[2]};
    const double ca[3] = {(1 - c) * u[0], (1 - c) * u[1], (1 - c) * u[2]};
    double tmp;
    tmp = ca[0] * u[1]; R[1] = tmp - sa[2]; R[3] = tmp + sa[2];
    tmp = ca[0] * u[2]; R[2] = tmp + sa[1]; R[6] = tmp - sa[1];
    tmp = ca[1] * u[2]; R[5] = tmp - sa[0]; R[7] = tmp + sa[0];
    R[0] = ca[0] * u[0] + c; R[4] = ca[1] * u[1] + c; R[8] = ca[2] * u[2] + c;
}
inline double det3(const double* m) {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

// Eigen LDLT (robust Cholesky with diagonal pivoting), restated: S (n x n SPD) solves S X = B
// (B n x m, overwritten with X).  Algorithm as published in Eigen/src/Cholesky/LDLT.h.
template <int N>
void ldlt_solve(const double* S, double* B, int m) {
    double A[N * N];
    std::memcpy(A, S, sizeof A);
    int tr[N];
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double big = std::fabs(A[k * N + k]);
        for (int i = k + 1; i < N; ++i)
            if (std::fabs(A[i * N + i]) > big) { big = std::fabs(A[i * N + i]); piv = i; }
        tr[k] = piv;
        if (piv != k) {  // symmetric swap working on the lower triangle only
            const int s = N - piv - 1;
            for (int j = 0; j < k; ++j) std::swap(A[k * N + j], A[piv * N + j]);
            for (int i = 0; i < s; ++i) std::swap(A[(piv + 1 + i) * N + k], A[(piv + 1 + i) * N + piv]);
            std::swap(A[k * N + k], A[piv * N + piv]);
            for (int i = k + 1; i < piv; ++i) std::swap(A[i * N + k], A[piv * N + i]);
        }
        const int rs = N - k - 1;
        if (k > 0) {
            double temp[N];
            for (int j = 0; j < k; ++j) temp[j] = A[j * N + j] * A[k * N + j];
            double s = 0.0;
            for (int j = 0; j < k; ++j) s += A[k * N + j] * temp[j];
            A[k * N + k] -= s;
            for (int i = 0; i < rs; ++i) {
                double s2 = 0.0;
                for (int j = 0; j < k; ++j) s2 += A[(k + 1 + i) * N + j] * temp[j];
                A[(k + 1 + i) * N + k] -= s2;
            }
        }
        const double piv_val = A[k * N + k];
        if (rs > 0 && std::fabs(piv_val) > 0.0)
            for (int i = 0; i < rs; ++i) A[(k + 1 + i) * N + k] /= piv_val;
    }
    // solve: X = P^T L^-T D^-1 L^-1 P B
    for (int k = 0; k < N; ++k)
        if (tr[k] != k)
            for (int c = 0; c < m; ++c) std::swap(B[k * m + c], B[tr[k] * m + c]);
    for (int i = 0; i < N; ++i)  // unit lower forward
        for (int j = 0; j < i; ++j)
            for (int c = 0; c < m; ++c) B[i * m + c] -= A[i * N + j] * B[j * m + c];
    const double tol = 1.0 / 1.7976931348623157e308;
    for (int i = 0; i < N; ++i)
        for (int c = 0; c < m; ++c) {
            if (std::fabs(A[i * N + i]) > tol) B[i * m + c] /= A[i * N + i];
            else B[i * m + c] = 0.0;
        }
    for (int i = N - 1; i >= 0; --i)  // unit upper backward (L^T)
        for (int j = i + 1; j < N; ++j)
            for (int c = 0; c < m; ++c) B[i * m + c] -= A[j * N + i] * B[j * m + c];
    for (int k = N - 1; k >= 0; --k)
        if (tr[k] != k)
            for (int c = 0; c < m; ++c) std::swap(B[k * m + c], B[tr[k] * m + c]);
}

// Smallest-eigenvalue unit eigenvector of a symmetric 3x3 (stands in for EigenSolver<Matrix3d>,
// vision.cpp:679-696: only that eigenvector is used and its sign is re-fixed afterwards, so any
// accurate symmetric method is equivalent -- SURVEY A.1-5).  Cyclic Jacobi, 12 sweeps max.
void smallest_eigvec_sym3(const double* Min, double* z) {
    double A[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    std::memcpy(A, Min, sizeof A);
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = std::fabs(A[1]) + std::fabs(A[2]) + std::fabs(A[5]);
        if (off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[p * 3 + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * 3 + q] - A[p * 3 + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k * 3 + p], akq = A[k * 3 + q];
                    A[k * 3 + p] = c * akp - s * akq;
                    A[k * 3 + q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p * 3 + k], aqk = A[q * 3 + k];
                    A[p * 3 + k] = c * apk - s * aqk;
                    A[q * 3 + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = c * vkp - s * vkq;
                    V[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
    // same column-selection rule as vision.cpp:680-694 applied to the Jacobi eigenvalues
    int col;
    if (A[0] < A[4]) col = (A[0] < A[8]) ? 0 : 2;
    else col = (A[4] < A[8]) ? 1 : 2;
    double v[3] = {V[col], V[3 + col], V[6 + col]};
    const double n = norm3(v);
    z[0] = v[0] / n; z[1] = v[1] / n; z[2] = v[2] / n;
}

// ------------------------------------------------------------------------------------------
// per-run constants derived from fbus_config
// ------------------------------------------------------------------------------------------
struct Consts {
    // FILTER view: T = diag(-1,-1,1,1) * TSC_left (filter.hpp:67-69)
    double R_IL[9], Q_IL[4], P_IL[3];
    // marker map
    int n_markers;
    int marker_id[FBUS_MAX_MARKERS];
    double marker_p[FBUS_MAX_MARKERS][3], marker_q[FBUS_MAX_MARKERS][4];
    double Q[12];   // diag of noiseCovariance (filter.hpp:108-115)
    double Rn[7];   // diag of observeNoiseCovariance (filter.hpp:118-120)
    double P0[6];
    double max_dist, switch_thres, reset_gap;
    // VISION view: raw T_SC (vision.hpp:83-84)
    double R_RL[9], P_LR[3];
    double n_air, n_glass, n_water, d_air, d_glass, normal[3], dect_thres;
    int flags;
};

void make_consts(const fbus_config* c, Consts* k) {
    const double flip[3] = {-1.0, -1.0, 1.0};
    double P_LI[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) k->R_IL[i * 3 + j] = flip[i] * c->tsc_left[i * 4 + j];
        P_LI[i] = flip[i] * c->tsc_left[i * 4 + 3];
    }
    R2q(k->R_IL, k->Q_IL);  // filter.cpp:370 -- NOT normalised
    double RT[9];
    mat3_t(k->R_IL, RT);
    double nRT[9];
    for (int i = 0; i < 9; ++i) nRT[i] = -RT[i];
    mat3_vec(nRT, P_LI, k->P_IL);  // filter.cpp:372: P_I_L = -R_I_L^T * P_L_I
    k->n_markers = c->n_markers;
    for (int m = 0; m < c->n_markers; ++m) {
        k->marker_id[m] = c->marker_id[m];
        for (int i = 0; i < 3; ++i) k->marker_p[m][i] = c->marker_pos[m * 3 + i];
        R2q(&c->marker_rot[m * 9], k->marker_q[m]);  // main.cpp:201
    }
    for (int i = 0; i < 3; ++i) {
        k->Q[i] = c->accel_n_cov; k->Q[3 + i] = c->gyro_n_cov;
        k->Q[6 + i] = c->accel_b_cov; k->Q[9 + i] = c->gyro_b_cov;
        k->Rn[i] = c->pos_n_cov;
    }
    for (int i = 3; i < 7; ++i) k->Rn[i] = c->quat_n_cov;
    for (int i = 0; i < 6; ++i) k->P0[i] = c->p0_diag[i];
    k->max_dist = c->marker_max_dist; k->switch_thres = c->marker_switch_thres; k->reset_gap = c->reset_gap;
    // vision.cpp:476-481 with the RAW T_SC of both cameras
    double R_IL_raw[9], R_IR_raw[9], P_LI_raw[3], P_RI_raw[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { R_IL_raw[i * 3 + j] = c->tsc_left[i * 4 + j]; R_IR_raw[i * 3 + j] = c->tsc_right[i * 4 + j]; }
        P_LI_raw[i] = c->tsc_left[i * 4 + 3]; P_RI_raw[i] = c->tsc_right[i * 4 + 3];
    }
    matmul_bt<3, 3, 3>(R_IL_raw, R_IR_raw, k->R_RL);  // R_R_L = R_I_L * R_I_R^T
    double t[3];
    mat3_vec(k->R_RL, P_RI_raw, t);
    for (int i = 0; i < 3; ++i) k->P_LR[i] = P_LI_raw[i] - t[i];  // P_L_R = P_L_I - R_R_L * P_R_I
    k->n_air = c->n_air; k->n_glass = c->n_glass; k->n_water = c->n_water;
    k->d_air = c->d_air; k->d_glass = c->d_glass;
    for (int i = 0; i < 3; ++i) k->normal[i] = c->normal[i];
    k->dect_thres = c->marker_dect_dist_thres;
    k->flags = c->flags;
}

int find_marker(const Consts& k, int id) {
    for (int m = 0; m < k.n_markers; ++m)
        if (k.marker_id[m] == id) return m;
    return -1;
}

// ------------------------------------------------------------------------------------------
// one filter (NominalState + ErrorState, common.hpp:205-247)
// ------------------------------------------------------------------------------------------
struct Filter {
    double t, q[4], R[9], p[3], v[3], ba[3], bg[3], g[3], pv[3], qv[4];
    double P[324];
    int prev_marker_id, initialised, status;
};

struct Det {  // DetectionResult, filter.hpp:39-56
    int id;
    double p[3], q[4];
};

void filter_ctor(const Consts& k, Filter* f) {  // filter.hpp:63-137, common.hpp:222-224
    std::memset(f, 0, sizeof *f);
    f->q[0] = 1.0;
    f->qv[0] = 1.0;
    for (int b = 0; b < 6; ++b)
        for (int i = 0; i < 3; ++i) f->P[(b * 3 + i) * 18 + b * 3 + i] = k.P0[b];
    f->prev_marker_id = 0;
}

// filter.cpp:588-616
void UpdateCovariance(const Consts& k, Filter* f, double dt, const double* accel, const double* gyro) {
    double w[3], a[3];
    for (int i = 0; i < 3; ++i) { w[i] = gyro[i] - f->bg[i]; a[i] = accel[i] - f->ba[i]; }
    double F[324];
    for (int i = 0; i < 324; ++i) F[i] = 0.0;
    for (int i = 0; i < 18; ++i) F[i * 18 + i] = 1.0;
    double Sa[9], Sw[9], nR[9], T[9];
    skew(a, Sa); skew(w, Sw);
    for (int i = 0; i < 9; ++i) nR[i] = -f->R[i];
    matmul<3, 3, 3>(nR, Sa, T);  // (-R * skew(a)) * dt
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            const double I = (i == j) ? 1.0 : 0.0;
            F[i * 18 + 3 + j] = I * dt;
            F[(3 + i) * 18 + 6 + j] = T[i * 3 + j] * dt;
            F[(3 + i) * 18 + 9 + j] = nR[i * 3 + j] * dt;
            F[(3 + i) * 18 + 15 + j] = I * dt;
            F[(6 + i) * 18 + 6 + j] = I - Sw[i * 3 + j] * dt;
            F[(6 + i) * 18 + 12 + j] = -I * dt;
        }
    double FP[324], FPFt[324];
    matmul<18, 18, 18>(F, f->P, FP);
    matmul_bt<18, 18, 18>(FP, F, FPFt);
    // Gamma Q Gamma^T with Gamma[3:15,0:12] = I (filter.hpp:124-125): dense product as in the reference
    double G[18 * 12], Qm[144], GQ[18 * 12], GQGt[324];
    for (int i = 0; i < 18 * 12; ++i) G[i] = 0.0;
    for (int i = 0; i < 12; ++i) G[(3 + i) * 12 + i] = 1.0;
    for (int i = 0; i < 144; ++i) Qm[i] = 0.0;
    for (int i = 0; i < 12; ++i) Qm[i * 12 + i] = k.Q[i];
    matmul<18, 12, 12>(G, Qm, GQ);
    matmul_bt<18, 12, 18>(GQ, G, GQGt);
    for (int i = 0; i < 324; ++i) FPFt[i] += GQGt[i];
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) f->P[i * 18 + j] = (FPFt[i * 18 + j] + FPFt[j * 18 + i]) / 2.0;
}

// filter.cpp:533-582
void UpdateNominalState(Filter* f, double dt, const double* accel, const double* gyro) {
    double w[3];
    for (int i = 0; i < 3; ++i) w[i] = gyro[i] - f->bg[i];
    const double wn = norm3(w);
    double qh[4], R0[9];
    q2R(f->q, R0);
    if (wn > 10e-5) {
        const double ax[3] = {w[0] / wn, w[1] / wn, w[2] / wn};
        const double ah = wn * dt / 2;
        const double dqh[4] = {std::cos(ah / 2), std::sin(ah / 2) * ax[0], std::sin(ah / 2) * ax[1], std::sin(ah / 2) * ax[2]};
        qmul(f->q, dqh, qh);
        qnormalize(qh);
        const double af = wn * dt;
        const double dq[4] = {std::cos(af / 2), std::sin(af / 2) * ax[0], std::sin(af / 2) * ax[1], std::sin(af / 2) * ax[2]};
        double qn[4];
        qmul(f->q, dq, qn);
        qnormalize(qn);
        std::memcpy(f->q, qn, sizeof qn);
    } else {
        const double dqh[4] = {1, 0.5 * dt * w[0] / 2, 0.5 * dt * w[1] / 2, 0.5 * dt * w[2] / 2};
        qmul(f->q, dqh, qh);
        qnormalize(qh);
        const double dq[4] = {1, 0.5 * dt * w[0], 0.5 * dt * w[1], 0.5 * dt * w[2]};
        double qn[4];
        qmul(f->q, dq, qn);
        qnormalize(qn);
        std::memcpy(f->q, qn, sizeof qn);
    }
    double Rh[9];
    q2R(qh, Rh);
    q2R(f->q, f->R);
    double a[3], k1[3], k2[3], k4[3];
    for (int i = 0; i < 3; ++i) a[i] = accel[i] - f->ba[i];
    mat3_vec(R0, a, k1); mat3_vec(Rh, a, k2); mat3_vec(f->R, a, k4);
    for (int i = 0; i < 3; ++i) { k1[i] += f->g[i]; k2[i] += f->g[i]; k4[i] += f->g[i]; }
    for (int i = 0; i < 3; ++i) {
        const double v0 = f->v[i];
        const double k3 = k2[i];
        f->v[i] = v0 + dt / 6 * (k1[i] + 2 * k2[i] + 2 * k3 + k4[i]);
        const double kp1 = v0, kp2 = v0 + k1[i] * dt / 2, kp3 = v0 + k2[i] * dt / 2, kp4 = v0 + k3 * dt / 2;
        f->p[i] = f->p[i] + dt / 6 * (kp1 + 2 * kp2 + 2 * kp3 + kp4);
    }
}

// filter.cpp:483-531 over an explicit candidate range
size_t BatchImuProcessing(const Consts& k, Filter* f, const double* t, const double* data, size_t B, size_t b,
                          size_t first, size_t count, double t_end) {
    const double start = f->t;
    size_t imuCnt = 0;  // samples erased afterwards (filter.cpp:492-503,520)
    for (size_t i = first; i < first + count; ++i) {
        if (t[i] < start) { imuCnt++; continue; }
        if (t[i] > t_end) break;
        imuCnt++;
        double accel[3], gyro[3];
        for (int c = 0; c < 3; ++c) { accel[c] = data[(i * 6 + c) * B + b]; gyro[c] = data[(i * 6 + 3 + c) * B + b]; }
        const double dt = t[i] - f->t;
        UpdateCovariance(k, f, dt, accel, gyro);
        UpdateNominalState(f, dt, accel, gyro);
        f->t = t[i];
    }
    return imuCnt;
}

// nearest-marker scan shared by filter.cpp:329-341, 418-430, 639-658
int nearest(const Det* d, int n, double* min_dist_out) {
    int idx = 0;
    double md = 10;
    for (int c = 0; c < n; ++c) {
        const double dist = norm3(d[c].p);
        if (dist < md) { md = dist; idx = c; }
    }
    *min_dist_out = md;
    return idx;
}

// filter.cpp:291-399
bool InitializePose(const Consts& k, Filter* f, const Det* d, int n, double t_det, size_t n_imu_before) {
    if (!(n_imu_before > 0 && n > 0)) { f->status |= FBUS_ST_INIT_FAILED; return false; }
    double md;
    const int idx = nearest(d, n, &md);
    if (md > k.max_dist) { f->status |= FBUS_ST_INIT_FAILED; return false; }
    const int m = find_marker(k, d[idx].id);
    if (m < 0) { f->status |= FBUS_ST_INIT_FAILED; return false; }
    double cq[4], t1[4], qig[4];
    qconj(d[idx].q, cq);
    qmul(k.marker_q[m], cq, t1);
    qmul(t1, k.Q_IL, qig);  // Q_I_G = Q_M_G * conj(Q_M_L) * Q_I_L (not normalised)
    f->t = t_det;
    std::memcpy(f->q, qig, sizeof qig);
    q2R(qig, f->R);
    double RP[3], RIT[9], RR[9], RRP[3];
    mat3_vec(f->R, k.P_IL, RP);
    mat3_t(k.R_IL, RIT);
    matmul<3, 3, 3>(f->R, RIT, RR);
    mat3_vec(RR, d[idx].p, RRP);
    for (int i = 0; i < 3; ++i) f->p[i] = k.marker_p[m][i] - RP[i] - RRP[i];
    f->g[0] = 9.8; f->g[1] = 0; f->g[2] = 0;  // filter.cpp:387
    return true;
}

// filter.cpp:405-477
void ResetSystemState(const Consts& k, Filter* f, const Det* d, int n, double t_det) {
    if (n <= 0) return;
    double md;
    const int idx = nearest(d, n, &md);
    if (md > k.max_dist) { f->status |= FBUS_ST_RESET_SKIPPED; return; }
    const int m = find_marker(k, d[idx].id);
    if (m < 0) { f->status |= FBUS_ST_RESET_SKIPPED; return; }
    double cq[4], t1[4];
    qconj(d[idx].q, cq);
    qmul(k.marker_q[m], cq, t1);
    qmul(t1, k.Q_IL, f->qv);
    double RIG[9], nRIG[9], RIT[9], M[9], v1[3], v2[3];
    q2R(f->qv, RIG);
    for (int i = 0; i < 9; ++i) nRIG[i] = -RIG[i];
    mat3_t(k.R_IL, RIT);
    matmul<3, 3, 3>(nRIG, RIT, M);
    mat3_vec(M, d[idx].p, v1);
    mat3_vec(RIG, k.P_IL, v2);
    for (int i = 0; i < 3; ++i) f->pv[i] = v1[i] + k.marker_p[m][i] - v2[i];
    if (t_det - f->t > k.reset_gap && f->initialised) {
        f->t = t_det;
        std::memcpy(f->q, f->qv, sizeof f->qv);
        for (int i = 0; i < 3; ++i) { f->p[i] = f->pv[i]; f->v[i] = 0; f->ba[i] = 0; f->bg[i] = 0; }
        f->status |= FBUS_ST_RESET_DONE;
    }
}

// filter.cpp:622-739
void ObservationUpdate(const Consts& k, Filter* f, const Det* d, int n) {
    if (n <= 0) return;
    int idx = 0;
    double md = 10, prev_dist = 0;
    int prev_idx = 0;
    for (int c = 0; c < n; ++c) {
        const double dist = norm3(d[c].p);
        if (dist < md) { md = dist; idx = c; }
        if (d[c].id == f->prev_marker_id) { prev_dist = dist; prev_idx = c; }
    }
    if (Absolute(prev_dist - md) < k.switch_thres && prev_dist != 0) idx = prev_idx;
    const int m = find_marker(k, d[idx].id);
    if (m < 0) { f->status |= FBUS_ST_UPDATE_SKIPPED; return; }
    f->prev_marker_id = d[idx].id;
    const double* yP = d[idx].p;
    const double* yQ = d[idx].q;
    const double* PGM = k.marker_p[m];
    // hP_L_M = R_I_L * R^T * (P_G_M - P_G_I - R * P_I_L)
    double RT[9], RILRT[9], RP[3], dv[3], hP[3];
    mat3_t(f->R, RT);
    matmul<3, 3, 3>(k.R_IL, RT, RILRT);
    mat3_vec(f->R, k.P_IL, RP);
    for (int i = 0; i < 3; ++i) dv[i] = PGM[i] - f->p[i] - RP[i];
    mat3_vec(RILRT, dv, hP);
    // hQ_L_M = Q_I_L * conj(q) * Q_M_G
    double cq[4], t1[4], hQ[4];
    qconj(f->q, cq);
    qmul(k.Q_IL, cq, t1);
    qmul(t1, k.marker_q[m], hQ);
    // H
    double H[7 * 18];
    for (int i = 0; i < 7 * 18; ++i) H[i] = 0.0;
    double nRIL[9], H00[9];
    for (int i = 0; i < 9; ++i) nRIL[i] = -k.R_IL[i];
    matmul<3, 3, 3>(nRIL, RT, H00);
    double dp[3], rtdp[3], S3[9], H06[9];
    for (int i = 0; i < 3; ++i) dp[i] = PGM[i] - f->p[i];
    mat3_vec(RT, dp, rtdp);
    skew(rtdp, S3);
    matmul<3, 3, 3>(k.R_IL, S3, H06);
    double Rq[16], Lil[16], L2[16], Lq[16], L1[12], A1[16], A2[16], A3[16], Hq[12];
    quat_right(k.marker_q[m], Rq);
    quat_left(k.Q_IL, Lil);
    quat_left(f->q, Lq);
    for (int i = 0; i < 16; ++i) L2[i] = 0.0;
    L2[0] = 1; L2[5] = -1; L2[10] = -1; L2[15] = -1;
    for (int i = 0; i < 12; ++i) L1[i] = 0.0;
    L1[3] = 0.5; L1[7] = 0.5; L1[11] = 0.5;
    matmul<4, 4, 4>(Rq, Lil, A1);
    matmul<4, 4, 4>(A1, L2, A2);
    matmul<4, 4, 4>(A2, Lq, A3);
    matmul<4, 4, 3>(A3, L1, Hq);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { H[i * 18 + j] = H00[i * 3 + j]; H[i * 18 + 6 + j] = H06[i * 3 + j]; }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 3; ++j) H[(3 + i) * 18 + 6 + j] = Hq[i * 3 + j];
    // sign disambiguation, filter.cpp:698-706 (strict >)
    double k1 = 0, k2 = 0;
    for (int i = 0; i < 4; ++i) { k1 += (yQ[i] - hQ[i]) * (yQ[i] - hQ[i]); k2 += (yQ[i] + hQ[i]) * (yQ[i] + hQ[i]); }
    if (k1 > k2) {
        for (int i = 0; i < 4; ++i) hQ[i] = -hQ[i];
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 3; ++j) H[(3 + i) * 18 + 6 + j] = -H[(3 + i) * 18 + 6 + j];
    }
    // S = H P H^T + R ; K^T = S.ldlt().solve(H P)
    double HP[7 * 18], S[49], KT[7 * 18];
    matmul<7, 18, 18>(H, f->P, HP);
    matmul_bt<7, 18, 7>(HP, H, S);
    for (int i = 0; i < 7; ++i) S[i * 7 + i] += k.Rn[i];
    matmul<7, 18, 18>(H, f->P, KT);  // H*P evaluated a second time, filter.cpp:711
    ldlt_solve<7>(S, KT, 18);
    double r[7];
    for (int i = 0; i < 3; ++i) r[i] = yP[i] - hP[i];
    for (int i = 0; i < 4; ++i) r[3 + i] = yQ[i] - hQ[i];
    double dx[18];
    for (int i = 0; i < 18; ++i) {
        double s = 0.0;
        for (int j = 0; j < 7; ++j) s += KT[j * 18 + i] * r[j];
        dx[i] = s;
    }
    for (int i = 0; i < 3; ++i) { f->p[i] += dx[i]; f->v[i] += dx[3 + i]; }
    {   // VectorToQuaterniond, matrix_math.hpp:90-99 (NaN at exactly zero, reproduced)
        const double* th = dx + 6;
        const double vn = norm3(th);
        const double dq[4] = {std::cos(vn / 2), th[0] / vn * std::sin(vn / 2), th[1] / vn * std::sin(vn / 2), th[2] / vn * std::sin(vn / 2)};
        double qn[4];
        qmul(f->q, dq, qn);
        qnormalize(qn);
        std::memcpy(f->q, qn, sizeof qn);  // rotmatI2G deliberately NOT refreshed (A.3-2)
    }
    for (int i = 0; i < 3; ++i) { f->ba[i] += dx[9 + i]; f->bg[i] += dx[12 + i]; f->g[i] += dx[15 + i]; }
    // P = (I - K H) P ; symmetrise
    double IKH[324], Pn[324];
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) {
            double s = 0.0;
            for (int c = 0; c < 7; ++c) s += KT[c * 18 + i] * H[c * 18 + j];
            IKH[i * 18 + j] = ((i == j) ? 1.0 : 0.0) - s;
        }
    matmul<18, 18, 18>(IKH, f->P, Pn);
    if (k.flags & FBUS_FLAG_JOSEPH) {
        // opt-in Joseph form (north_star; not what the reference executes): (I-KH) P (I-KH)^T + K R K^T
        double J1[324];
        matmul_bt<18, 18, 18>(Pn, IKH, J1);
        for (int i = 0; i < 18; ++i)
            for (int j = 0; j < 18; ++j) {
                double s = 0.0;
                for (int c = 0; c < 7; ++c) s += KT[c * 18 + i] * k.Rn[c] * KT[c * 18 + j];
                Pn[i * 18 + j] = J1[i * 18 + j] + s;
            }
    }
    for (int i = 0; i < 18; ++i)
        for (int j = 0; j < 18; ++j) f->P[i * 18 + j] = (Pn[i * 18 + j] + Pn[j * 18 + i]) / 2.0;
}

int gather_dets(const fbus_det_frames* det, size_t frame, size_t b, Det* out) {
    const size_t B = det->batch, m = det->max_markers;
    int n = 0;
    for (size_t s = 0; s < m; ++s) {
        const int id = det->id[(frame * m + s) * B + b];
        if (id < 0) continue;
        out[n].id = id;
        const double* base = det->pose + ((frame * m + s) * 7) * B + b;
        for (int c = 0; c < 3; ++c) out[n].p[c] = base[c * B];
        for (int c = 0; c < 4; ++c) out[n].q[c] = base[(3 + c) * B];
        ++n;
    }
    return n;
}

void check_finite(Filter* f) {
    bool ok = std::isfinite(f->t);
    for (int i = 0; i < 4; ++i) ok = ok && std::isfinite(f->q[i]);
    for (int i = 0; i < 3; ++i) ok = ok && std::isfinite(f->p[i]) && std::isfinite(f->v[i]);
    if (!ok) f->status |= FBUS_ST_NONFINITE;
}

// ------------------------------------------------------------------------------------------
// R1: vision.cpp:488-608 for one marker; corners = 16 float32 (Lx0,Ly0,..,Lx3,Ly3,Rx0,..,Ry3)
// ------------------------------------------------------------------------------------------
bool RefractionTriangulation(const Consts& k, const float* c16, double* P3 /*[4][3]*/) {
    bool out_of_range = false;
    for (int i = 0; i < 4; ++i) {
        const double lp[3] = {(double)c16[2 * i], (double)c16[2 * i + 1], 1.0};
        const double rp[3] = {(double)c16[8 + 2 * i], (double)c16[8 + 2 * i + 1], 1.0};
        const double ln = norm3(lp), rn = norm3(rp);
        double r0L[3], r0R[3];
        for (int j = 0; j < 3; ++j) { r0L[j] = lp[j] / ln; r0R[j] = rp[j] / rn; }
        const double* nv = k.normal;
        const double a0 = k.n_air / k.n_glass;
        const double v0L = r0L[0] * nv[0] + r0L[1] * nv[1] + r0L[2] * nv[2];
        const double v0R = r0R[0] * nv[0] + r0R[1] * nv[1] + r0R[2] * nv[2];
        double b0L, b0R;
        if (k.n_air < k.n_glass) {
            b0L = std::sqrt(1 - a0 * a0 * (1 - v0L * v0L)) - a0 * v0L;
            b0R = std::sqrt(1 - a0 * a0 * (1 - v0R * v0R)) - a0 * v0R;
        } else {
            b0L = a0 * v0L - std::sqrt(1 - a0 * a0 * (1 - v0L * v0L));
            b0R = a0 * v0R - std::sqrt(1 - a0 * a0 * (1 - v0R * v0R));
        }
        double r1L[3], r1R[3];
        for (int j = 0; j < 3; ++j) { r1L[j] = a0 * r0L[j] + b0L * nv[j]; r1R[j] = a0 * r0R[j] + b0R * nv[j]; }
        const double a1 = k.n_glass / k.n_water;
        const double v1L = r1L[0] * nv[0] + r1L[1] * nv[1] + r1L[2] * nv[2];
        const double v1R = r1R[0] * nv[0] + r1R[1] * nv[1] + r1R[2] * nv[2];
        double b1L, b1R;
        if (k.n_glass > k.n_water) {
            b1L = std::sqrt(1 - a1 * a1 * (1 - v1L * v1L)) - a1 * v1L;
            b1R = std::sqrt(1 - a1 * a1 * (1 - v1R * v1R)) - a1 * v1R;
        } else {
            b1L = a1 * v1L - std::sqrt(1 - a1 * a1 * (1 - v1L * v1L));
            b1R = a1 * v1R - std::sqrt(1 - a1 * a1 * (1 - v1R * v1R));
        }
        double r2L[3], r2R[3];
        for (int j = 0; j < 3; ++j) { r2L[j] = a1 * r1L[j] + b1L * nv[j]; r2R[j] = a1 * r1R[j] + b1R * nv[j]; }
        const double d0 = k.d_air, d1 = k.d_glass;
        double P1L[3], P1R[3];
        for (int j = 0; j < 3; ++j) {
            const double P0L = (d0 / v0L) * r0L[j], P0R = (d0 / v0R) * r0R[j];
            P1L[j] = P0L + (d1 / v1L) * r1L[j];
            P1R[j] = P0R + (d1 / v1R) * r1R[j];
        }
        double r2RL[3], t[3], P1RL[3];
        mat3_vec(k.R_RL, r2R, r2RL);
        mat3_vec(k.R_RL, P1R, t);
        for (int j = 0; j < 3; ++j) P1RL[j] = k.P_LR[j] + t[j];
        const double cr[3] = {r2L[1] * r2RL[2] - r2L[2] * r2RL[1], r2L[2] * r2RL[0] - r2L[0] * r2RL[2],
                              r2L[0] * r2RL[1] - r2L[1] * r2RL[0]};
        const double dd[3] = {P1RL[0] - P1L[0], P1RL[1] - P1L[1], P1RL[2] - P1L[2]};
        double X1[9], X2[9], X3[9];
        for (int j = 0; j < 3; ++j) {
            X1[j * 3] = cr[j]; X1[j * 3 + 1] = dd[j];  X1[j * 3 + 2] = r2RL[j];
            X2[j * 3] = cr[j]; X2[j * 3 + 1] = r2L[j]; X2[j * 3 + 2] = dd[j];
            X3[j * 3] = cr[j]; X3[j * 3 + 1] = r2L[j]; X3[j * 3 + 2] = r2RL[j];
        }
        const double t1 = det3(X1) / det3(X3);
        const double t2 = -det3(X2) / det3(X3);
        double P[3];
        for (int j = 0; j < 3; ++j) P[j] = 0.5 * (P1L[j] + t1 * r2L[j] + P1RL[j] + t2 * r2RL[j]);
        P3[i * 3 + 0] = -P[0]; P3[i * 3 + 1] = -P[1]; P3[i * 3 + 2] = P[2];  // R_I_C = diag(-1,-1,1)
        if (norm3(P) > k.dect_thres) { out_of_range = true; break; }
    }
    return !out_of_range;
}

// Right singular vector of the smallest singular value of a 6x4 matrix by one-sided (Hestenes) Jacobi SVD; stands in for
// Eigen::JacobiSVD(A).matrixV().col(3) (vision.cpp:432-437; singular values sorted decreasingly, the sign cancels below).
void smallest_right_singular_vec_6x4(const double* Ain, double* v) {
    double A[24], V[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    std::memcpy(A, Ain, sizeof A);
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < 3; ++p)
            for (int q = p + 1; q < 4; ++q) {
                double a = 0, b = 0, c = 0;
                for (int r = 0; r < 6; ++r) { a += A[r * 4 + p] * A[r * 4 + p]; b += A[r * 4 + q] * A[r * 4 + q]; c += A[r * 4 + p] * A[r * 4 + q]; }
                off = std::fmax(off, std::fabs(c) / std::sqrt(a * b + 1e-300));
                if (c == 0.0) continue;
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / std::sqrt(1.0 + t * t), sn = cs * t;
                for (int r = 0; r < 6; ++r) {
                    const double x = A[r * 4 + p], y = A[r * 4 + q];
                    A[r * 4 + p] = cs * x - sn * y;
                    A[r * 4 + q] = sn * x + cs * y;
                }
                for (int r = 0; r < 4; ++r) {
                    const double x = V[r * 4 + p], y = V[r * 4 + q];
                    V[r * 4 + p] = cs * x - sn * y;
                    V[r * 4 + q] = sn * x + cs * y;
                }
            }
        if (off < 1e-17) break;
    }
    int best = 0;
    double bn = 1e300;
    for (int j = 0; j < 4; ++j) {
        double nj = 0;
        for (int r = 0; r < 6; ++r) nj += A[r * 4 + j] * A[r * 4 + j];
        if (nj < bn) { bn = nj; best = j; }
    }
    for (int r = 0; r < 4; ++r) v[r] = V[r * 4 + best];
}

// N3: vision.cpp:395-466 for one marker (land mode)
bool NormalTriangulation(const fbus_config* cfg, const Consts& k, const float* c16, double* P3) {
    double T0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}, TLR[12];
    double R_IL[9], R_IR[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) { R_IL[i * 3 + j] = cfg->tsc_left[i * 4 + j]; R_IR[i * 3 + j] = cfg->tsc_right[i * 4 + j]; }
    double Rm[9];
    matmul_bt<3, 3, 3>(R_IR, R_IL, Rm);  // T_I_R * T_I_L^T
    const double P_LI[3] = {cfg->tsc_left[3], cfg->tsc_left[7], cfg->tsc_left[11]};
    const double P_RI[3] = {cfg->tsc_right[3], cfg->tsc_right[7], cfg->tsc_right[11]};
    double t[3];
    mat3_vec(Rm, P_RI, t);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) TLR[i * 4 + j] = Rm[i * 3 + j]; TLR[i * 4 + 3] = P_LI[i] - t[i]; }
    bool out = false;
    for (int i = 0; i < 4; ++i) {
        const double lp[3] = {(double)c16[2 * i], (double)c16[2 * i + 1], 1.0};
        const double rp[3] = {(double)c16[8 + 2 * i], (double)c16[8 + 2 * i + 1], 1.0};
        double Sl[9], Sr[9], A[24];
        skew(lp, Sl); skew(rp, Sr);
        matmul<3, 3, 4>(Sl, T0, A);
        matmul<3, 3, 4>(Sr, TLR, A + 12);
        double P[4];
        smallest_right_singular_vec_6x4(A, P);
        if (P[3] == 0) continue;
        double Pn[3] = {-(P[0] / P[3]), -(P[1] / P[3]), P[2] / P[3]};
        const double sg = Signum(Pn[2]);
        for (int j = 0; j < 3; ++j) P3[i * 3 + j] = sg * Pn[j];
        if (norm3(Pn) > k.dect_thres) { out = true; break; }
    }
    return !out;
}

// R2: vision.cpp:635-759 for one marker; C = 4 corners x 3
void ComputeMarkerPose(const double* C, double* p, double* q, double* Rml) {
    const double* c0 = C; const double* c1 = C + 3; const double* c2 = C + 6; const double* c3 = C + 9;
    double v[6][3];
    for (int j = 0; j < 3; ++j) {
        v[0][j] = c1[j] - c0[j]; v[1][j] = c2[j] - c0[j]; v[2][j] = c3[j] - c0[j];
        v[3][j] = c2[j] - c1[j]; v[4][j] = c3[j] - c1[j]; v[5][j] = c3[j] - c2[j];
    }
    double M[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = v[0][i] * v[0][j];
            for (int e = 1; e < 6; ++e) s = s + v[e][i] * v[e][j];
            M[i * 3 + j] = s;
        }
    double Z[3];
    smallest_eigvec_sym3(M, Z);
    if (Z[2] > 0.1) {
        for (int j = 0; j < 3; ++j) Z[j] = -1 * Z[j];
    } else if (Z[2] < -0.1) {
    } else {
        const double s = -Signum(c0[0]) * Signum(Z[0]);
        for (int j = 0; j < 3; ++j) Z[j] = s * Z[j];
    }
    double sum[3];
    for (int j = 0; j < 3; ++j) sum[j] = c0[j] + c1[j] + c2[j] + c3[j];
    const double D = 0.25 * (Z[0] * sum[0] + Z[1] * sum[1] + Z[2] * sum[2]);
    double Pp[4][3];
    for (int i = 0; i < 4; ++i) {
        const double* c = C + 3 * i;
        const double t = (Z[0] * c[0] + Z[1] * c[1] + Z[2] * c[2]) - D;
        for (int j = 0; j < 3; ++j) Pp[i][j] = c[j] - t * Z[j];
    }
    double V12[3], V14[3], m[3];
    for (int j = 0; j < 3; ++j) { V12[j] = Pp[1][j] - Pp[0][j]; V14[j] = Pp[3][j] - Pp[0][j]; }
    const double n12 = norm3(V12), n14 = norm3(V14);
    for (int j = 0; j < 3; ++j) m[j] = V12[j] / n12 + V14[j] / n14;
    double Rm[9], Rmm[3], X[3], Y[3];
    angleaxis_matrix(-REF_M_PI / 4, Z, Rm);
    mat3_vec(Rm, m, Rmm);
    const double mn = norm3(m);
    for (int j = 0; j < 3; ++j) X[j] = Rmm[j] / mn;
    Y[0] = Z[1] * X[2] - Z[2] * X[1]; Y[1] = Z[2] * X[0] - Z[0] * X[2]; Y[2] = Z[0] * X[1] - Z[1] * X[0];
    double R[9] = {X[0], Y[0], Z[0], X[1], Y[1], Z[1], X[2], Y[2], Z[2]};
    R2q(R, q);
    for (int j = 0; j < 3; ++j) p[j] = Pp[0][j];
    if (Rml) std::memcpy(Rml, R, sizeof R);
}

template <class Fn>
void parallel_for(size_t n, int n_threads, Fn fn) {
    if (n_threads <= 1 || n < 2) { fn(0, n); return; }
    // one thread per core the process may run on, pinned 1:1
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
        if (!cpus.empty()) {
            cpu_set_t one;
            CPU_ZERO(&one);
            CPU_SET(cpus[i % cpus.size()], &one);
            pthread_setaffinity_np(th.back().native_handle(), sizeof one, &one);
        }
    }
    for (auto& x : th) x.join();
}

}  // namespace

struct orc_handle {
    Consts k;
    std::vector<Filter> f;
    // continuation rule of fbus_step_windows (include/fbus_ekf.h): first unconsumed IMU sample per filter after the last call,
    // and the identity of that call
    std::vector<size_t> cursor;
    const void* last_imu_data = nullptr;
    size_t last_n_samples = 0, last_w1 = 0;
};

extern "C" {

orc_handle* orc_create(const fbus_config* cfg, size_t batch) {
    orc_handle* h = new orc_handle;
    make_consts(cfg, &h->k);
    h->f.resize(batch);
    for (auto& f : h->f) filter_ctor(h->k, &f);
    return h;
}
void orc_destroy(orc_handle* h) { delete h; }

// filter.cpp:256-285
int orc_init_gravity_gyrobias(orc_handle* h, const fbus_imu_stream* imu, size_t first, size_t count) {
    const size_t B = imu->batch;
    if (B != h->f.size() || first + count > imu->n_samples) return FBUS_E_BADARG;
    if (count == 0) return FBUS_OK;
    for (size_t b = 0; b < B; ++b) {
        double am[3] = {0, 0, 0}, gm[3] = {0, 0, 0};
        for (size_t i = first; i < first + count; ++i)
            for (int c = 0; c < 3; ++c) { am[c] = am[c] + imu->data[(i * 6 + c) * B + b]; gm[c] = gm[c] + imu->data[(i * 6 + 3 + c) * B + b]; }
        Filter& f = h->f[b];
        const double n = (double)(int)count;
        double a[3];
        for (int c = 0; c < 3; ++c) { f.bg[c] = gm[c] / n; a[c] = am[c] / n; }
        f.g[0] = 0; f.g[1] = 0; f.g[2] = -norm3(a);
    }
    return FBUS_OK;
}

int orc_init_position_quaternion(orc_handle* h, const fbus_det_frames* det, size_t frame, size_t n_imu_before) {
    if (det->batch != h->f.size() || frame >= det->n_frames) return FBUS_E_BADARG;
    std::vector<Det> d(det->max_markers);
    for (size_t b = 0; b < h->f.size(); ++b) {
        Filter& f = h->f[b];
        const int n = gather_dets(det, frame, b, d.data());
        if (InitializePose(h->k, &f, d.data(), n, det->t[frame], n_imu_before)) f.initialised = 1;
    }
    return FBUS_OK;
}

int orc_propagate(orc_handle* h, const fbus_imu_stream* imu, size_t first, size_t count, double t_end) {
    if (imu->batch != h->f.size() || first + count > imu->n_samples) return FBUS_E_BADARG;
    for (size_t b = 0; b < h->f.size(); ++b) {
        BatchImuProcessing(h->k, &h->f[b], imu->t, imu->data, imu->batch, b, first, count, t_end);
        check_finite(&h->f[b]);
    }
    return FBUS_OK;
}

int orc_reset_state(orc_handle* h, const fbus_det_frames* det, size_t frame) {
    if (det->batch != h->f.size() || frame >= det->n_frames) return FBUS_E_BADARG;
    std::vector<Det> d(det->max_markers);
    for (size_t b = 0; b < h->f.size(); ++b) {
        const int n = gather_dets(det, frame, b, d.data());
        ResetSystemState(h->k, &h->f[b], d.data(), n, det->t[frame]);
    }
    return FBUS_OK;
}

int orc_update(orc_handle* h, const fbus_det_frames* det, size_t frame) {
    if (det->batch != h->f.size() || frame >= det->n_frames) return FBUS_E_BADARG;
    std::vector<Det> d(det->max_markers);
    for (size_t b = 0; b < h->f.size(); ++b) {
        const int n = gather_dets(det, frame, b, d.data());
        if (n == 0) h->f[b].status |= FBUS_ST_NO_DETECTION;
        ObservationUpdate(h->k, &h->f[b], d.data(), n);
        check_finite(&h->f[b]);
    }
    return FBUS_OK;
}

// body of FilterThreadFunction, filter.cpp:207-235, per frame and per filter
int orc_step_windows(orc_handle* h, const fbus_imu_stream* imu, const fbus_det_frames* det,
                     const uint32_t* win_off, size_t w0, size_t w1, double* trace, int n_threads) {
    const size_t B = h->f.size();
    if (imu->batch != B || det->batch != B || w1 > det->n_frames || w0 > w1) return FBUS_E_BADARG;
    const bool resume = w0 > 0 && w0 == h->last_w1 && imu->data == h->last_imu_data && imu->n_samples == h->last_n_samples;
    h->cursor.resize(B, 0);
    h->last_imu_data = imu->data;
    h->last_n_samples = imu->n_samples;
    h->last_w1 = w1;
    parallel_for(B, n_threads, [&](size_t lo, size_t hi) {
        std::vector<Det> d(det->max_markers);
        for (size_t b = lo; b < hi; ++b) {
            Filter& f = h->f[b];
            // `cursor` = first IMU sample still in the reference's imuMeasuementBuffer_: samples are
            // erased only when a frame consumes them (filter.cpp:390,520); frames that do nothing
            // leave them buffered for the next frame.
            size_t cursor = resume ? h->cursor[b] : win_off[w0];
            for (size_t w = w0; w < w1; ++w) {
                const int n = gather_dets(det, w, b, d.data());
                const size_t first = cursor, count = win_off[w + 1] - cursor;
                const double t_det = det->t[w];
                if (n == 0) {
                    // the reference's filter thread is only woken by non-empty detection lists
                    // (vision.cpp:136-140): nothing happens for this filter in this frame
                    f.status |= FBUS_ST_NO_DETECTION;
                } else if (!f.initialised) {
                    size_t cnt = 0;  // imuCnt: leading buffered samples not later than the frame (filter.cpp:299-305)
                    for (size_t i = first; i < first + count; ++i) { if (imu->t[i] > t_det) break; ++cnt; }
                    // only those are erased (filter.cpp:390); later samples stay buffered for the next frame
                    if (InitializePose(h->k, &f, d.data(), n, t_det, cnt)) { f.initialised = 1; cursor = first + cnt; }
                } else {
                    ResetSystemState(h->k, &f, d.data(), n, t_det);
                    cursor = first + BatchImuProcessing(h->k, &f, imu->t, imu->data, B, b, first, count, t_det);
                    ObservationUpdate(h->k, &f, d.data(), n);
                }
                if (trace) {
                    double* row = trace + ((w - w0) * 17) * B + b;
                    row[0] = f.t;
                    for (int c = 0; c < 3; ++c) row[(1 + c) * B] = f.p[c];
                    for (int c = 0; c < 4; ++c) row[(4 + c) * B] = f.q[c];
                    for (int c = 0; c < 3; ++c) { row[(8 + c) * B] = f.v[c]; row[(11 + c) * B] = f.ba[c]; row[(14 + c) * B] = f.bg[c]; }
                }
            }
            h->cursor[b] = cursor;
            check_finite(&f);
        }
    });
    return FBUS_OK;
}

int orc_refract_solve(const fbus_config* cfg, const float* corners, size_t n, double* pose, double* corners3d,
                      int32_t* valid, int n_threads) {
    Consts k;
    make_consts(cfg, &k);
    parallel_for(n, n_threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            float c16[16];
            for (int c = 0; c < 16; ++c) c16[c] = corners[(size_t)c * n + i];
            double P3[12];
            for (int c = 0; c < 12; ++c) P3[c] = 0.0;
            const bool ok = RefractionTriangulation(k, c16, P3);
            double p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
            if (ok) ComputeMarkerPose(P3, p, q, nullptr);
            for (int c = 0; c < 3; ++c) pose[(size_t)c * n + i] = p[c];
            for (int c = 0; c < 4; ++c) pose[(size_t)(3 + c) * n + i] = q[c];
            if (corners3d) for (int c = 0; c < 12; ++c) corners3d[(size_t)c * n + i] = P3[c];
            if (valid) valid[i] = ok ? 1 : 0;
        }
    });
    return FBUS_OK;
}

int orc_inair_solve(const fbus_config* cfg, const float* corners, size_t n, double* pose, double* corners3d, int32_t* valid) {
    Consts k;
    make_consts(cfg, &k);
    for (size_t i = 0; i < n; ++i) {
        float c16[16];
        for (int c = 0; c < 16; ++c) c16[c] = corners[(size_t)c * n + i];
        double P3[12];
        for (int c = 0; c < 12; ++c) P3[c] = 0.0;
        const bool ok = NormalTriangulation(cfg, k, c16, P3);
        double p[3] = {0, 0, 0}, q[4] = {1, 0, 0, 0};
        if (ok) ComputeMarkerPose(P3, p, q, nullptr);
        for (int c = 0; c < 3; ++c) pose[(size_t)c * n + i] = p[c];
        for (int c = 0; c < 4; ++c) pose[(size_t)(3 + c) * n + i] = q[c];
        if (corners3d) for (int c = 0; c < 12; ++c) corners3d[(size_t)c * n + i] = P3[c];
        if (valid) valid[i] = ok ? 1 : 0;
    }
    return FBUS_OK;
}

// N4: cv::fisheye::undistortPoints(distorted, undistorted, K, D), OpenCV 3.4.3 (vision.cpp:203,253,318,369)
int orc_undistort_fisheye(const fbus_config* cfg, const float* pixels, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i)
        for (int e = 0; e < 8; ++e) {
            const int cam = e >> 2;
            const double* K = cfg->cam_k[cam];
            const double* D = cfg->cam_d[cam];
            const double u = pixels[(size_t)(2 * e) * n + i], v = pixels[(size_t)(2 * e + 1) * n + i];
            const double pwx = (u - K[2]) / K[0], pwy = (v - K[3]) / K[1];
            double scale = 1.0;
            double theta_d = std::sqrt(pwx * pwx + pwy * pwy);
            theta_d = std::min(std::max(-3.14159265358979323846 / 2., theta_d), 3.14159265358979323846 / 2.);
            if (theta_d > 1e-8) {
                double theta = theta_d;
                const double EPS = 1e-8;
                for (int j = 0; j < 10; j++) {
                    const double theta2 = theta * theta, theta4 = theta2 * theta2, theta6 = theta4 * theta2, theta8 = theta6 * theta2;
                    const double k0_theta2 = D[0] * theta2, k1_theta4 = D[1] * theta4, k2_theta6 = D[2] * theta6, k3_theta8 = D[3] * theta8;
                    const double theta_fix = (theta * (1 + k0_theta2 + k1_theta4 + k2_theta6 + k3_theta8) - theta_d) /
                                             (1 + 3 * k0_theta2 + 5 * k1_theta4 + 7 * k2_theta6 + 9 * k3_theta8);
                    theta = theta - theta_fix;
                    if (std::fabs(theta_fix) < EPS) break;
                }
                scale = std::tan(theta) / theta_d;
            }
            out[(size_t)(2 * e) * n + i] = (float)(pwx * scale);
            out[(size_t)(2 * e + 1) * n + i] = (float)(pwy * scale);
        }
    return FBUS_OK;
}

int orc_marker_pose(const fbus_config* cfg, const double* corners3d, size_t n, double* pose) {
    (void)cfg;
    for (size_t i = 0; i < n; ++i) {
        double C[12], p[3], q[4];
        for (int c = 0; c < 12; ++c) C[c] = corners3d[(size_t)c * n + i];
        ComputeMarkerPose(C, p, q, nullptr);
        for (int c = 0; c < 3; ++c) pose[(size_t)c * n + i] = p[c];
        for (int c = 0; c < 4; ++c) pose[(size_t)(3 + c) * n + i] = q[c];
    }
    return FBUS_OK;
}

#define ORC_COPY(field, n, dir)                                                     \
    if (s->field)                                                                   \
        for (int c = 0; c < (n); ++c) {                                             \
            if (dir) s->field[(size_t)c * B + b] = f.field[c];                      \
            else f.field[c] = s->field[(size_t)c * B + b];                          \
        }

static void copy_state(orc_handle* h, fbus_state_soa* s, int get) {
    const size_t B = h->f.size();
    for (size_t b = 0; b < B; ++b) {
        Filter& f = h->f[b];
        if (s->t) { if (get) s->t[b] = f.t; else f.t = s->t[b]; }
        ORC_COPY(q, 4, get) ORC_COPY(R, 9, get) ORC_COPY(p, 3, get) ORC_COPY(v, 3, get)
        ORC_COPY(ba, 3, get) ORC_COPY(bg, 3, get) ORC_COPY(g, 3, get) ORC_COPY(pv, 3, get)
        ORC_COPY(qv, 4, get) ORC_COPY(P, 324, get)
        if (s->prev_marker_id) { if (get) s->prev_marker_id[b] = f.prev_marker_id; else f.prev_marker_id = s->prev_marker_id[b]; }
        if (s->initialised) { if (get) s->initialised[b] = f.initialised; else f.initialised = s->initialised[b]; }
        if (s->status) { if (get) s->status[b] = f.status; else f.status = s->status[b]; }
    }
}

int orc_get_state(orc_handle* h, fbus_state_soa* out) {
    if (out->batch != h->f.size()) return FBUS_E_BADARG;
    copy_state(h, out, 1);
    return FBUS_OK;
}
int orc_set_state(orc_handle* h, const fbus_state_soa* in) {
    if (in->batch != h->f.size()) return FBUS_E_BADARG;
    copy_state(h, const_cast<fbus_state_soa*>(in), 0);
    return FBUS_OK;
}

// statistics as defined in include/fbus_ekf.h (new; no reference counterpart)
int orc_stats(orc_handle* h, const double* truth_p, const double* truth_q, double* out) {
    const size_t B = h->f.size();
    for (int i = 0; i < FBUS_NSTATS; ++i) out[i] = 0.0;
    for (size_t b = 0; b < B; ++b) {
        const Filter& f = h->f[b];
        double e[6], qt[4], cq[4], dq[4];
        for (int c = 0; c < 3; ++c) e[c] = f.p[c] - truth_p[(size_t)c * B + b];
        for (int c = 0; c < 4; ++c) qt[c] = truth_q[(size_t)c * B + b];
        qconj(f.q, cq);
        qmul(cq, qt, dq);
        const double sg = dq[0] < 0 ? -1.0 : 1.0;
        for (int c = 0; c < 3; ++c) e[3 + c] = 2.0 * sg * dq[1 + c];
        // 6x6 pose block of P: rows/cols {0,1,2,6,7,8}
        const int ix[6] = {0, 1, 2, 6, 7, 8};
        double A[36], L[36];
        for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) A[i * 6 + j] = f.P[ix[i] * 18 + ix[j]];
        bool ok = true;
        for (int i = 0; i < 36; ++i) L[i] = 0.0;
        for (int j = 0; j < 6 && ok; ++j) {
            double s = A[j * 6 + j];
            for (int c = 0; c < j; ++c) s -= L[j * 6 + c] * L[j * 6 + c];
            if (!(s > 0.0)) { ok = false; break; }
            L[j * 6 + j] = std::sqrt(s);
            for (int i = j + 1; i < 6; ++i) {
                double s2 = A[i * 6 + j];
                for (int c = 0; c < j; ++c) s2 -= L[i * 6 + c] * L[j * 6 + c];
                L[i * 6 + j] = s2 / L[j * 6 + j];
            }
        }
        double y[6], nees = 0.0;
        if (ok) {
            for (int i = 0; i < 6; ++i) {
                double s = e[i];
                for (int c = 0; c < i; ++c) s -= L[i * 6 + c] * y[c];
                y[i] = s / L[i * 6 + i];
                nees += y[i] * y[i];
            }
        }
        const double ep2 = e[0] * e[0] + e[1] * e[1] + e[2] * e[2];
        const double et2 = e[3] * e[3] + e[4] * e[4] + e[5] * e[5];
        if (ok && std::isfinite(ep2) && std::isfinite(et2) && std::isfinite(nees)) {
            out[0] += ep2; out[1] += et2; out[2] += nees; out[3] += 1.0;
            if (std::sqrt(ep2) > out[5]) out[5] = std::sqrt(ep2);
        } else {
            out[4] += 1.0;
        }
    }
    return FBUS_OK;
}

}  // extern "C"
