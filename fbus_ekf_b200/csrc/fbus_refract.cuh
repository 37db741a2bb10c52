// fbus_refract.cuh -- refractive flat-port marker-pose solve, per-marker device inlines.
//   R1  VISION::RefractionTriangulation  (C++/src/vision.cpp:488-608)
//   R2  VISION::ComputeMarkerPose        (C++/src/vision.cpp:635-759)
// One thread per marker; everything lives in registers.  Also host-compilable (tests only).
#pragma once

#include "fbus_math.cuh"

namespace fbus {

// common.hpp:14 -- the reference overrides M_PI; R2's -pi/4 rotation must use this value (A.3-1)
constexpr double REF_M_PI = 3.1415926;

// Snell refraction of unit ray r at a plane with normal nv (vision.cpp:505-543):
//   r' = alpha*r + beta*nv, beta = sqrt(1-alpha^2(1-v^2)) - alpha*v  (or its negation, see `rule`)
FBUS_HD void refract_ray(const double* r, const double* nv, double alpha, int rule, double* out, double* v_out) {
    const double v = r[0] * nv[0] + r[1] * nv[1] + r[2] * nv[2];
    const double root = sqrt(1 - alpha * alpha * (1 - v * v));
    const double beta = rule ? (root - alpha * v) : (alpha * v - root);
    FBUS_UNROLL
    for (int j = 0; j < 3; ++j) out[j] = alpha * r[j] + beta * nv[j];
    *v_out = v;
}

FBUS_HD double det3_cols(const double* a, const double* b, const double* c) {  // det of [a b c] (columns)
    // cofactor expansion along the first row of the matrix whose columns are a,b,c
    return a[0] * (b[1] * c[2] - c[1] * b[2]) - b[0] * (a[1] * c[2] - c[1] * a[2]) + c[0] * (a[1] * b[2] - b[1] * a[2]);
}

// R1 for one stereo corner pair: (xl,yl),(xr,yr) float32-valued normalised coordinates -> P (3), flipped
// by R_I_C = diag(-1,-1,1); returns |P| (the unflipped norm, vision.cpp:601)
FBUS_HD double triangulate_corner(const DevConsts& k, double xl, double yl, double xr, double yr, double* Pout) {
    double r0L[3] = {xl, yl, 1.0}, r0R[3] = {xr, yr, 1.0};
    {
        const double il = 1.0 / norm3(r0L), ir = 1.0 / norm3(r0R);
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) { r0L[j] *= il; r0R[j] *= ir; }
    }
    double r1L[3], r1R[3], r2L[3], r2R[3], v0L, v0R, v1L, v1R;
    refract_ray(r0L, k.normal, k.a0, k.air_lt_glass, r1L, &v0L);
    refract_ray(r0R, k.normal, k.a0, k.air_lt_glass, r1R, &v0R);
    refract_ray(r1L, k.normal, k.a1, k.glass_gt_water, r2L, &v1L);
    refract_ray(r1R, k.normal, k.a1, k.glass_gt_water, r2R, &v1R);
    double P1L[3], P1R[3];
    {
        const double s0L = k.d_air / v0L, s0R = k.d_air / v0R, s1L = k.d_glass / v1L, s1R = k.d_glass / v1R;
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) {
            P1L[j] = s0L * r0L[j] + s1L * r1L[j];
            P1R[j] = s0R * r0R[j] + s1R * r1R[j];
        }
    }
    double rR[3], pR[3];
    mat3_vec(k.R_RL, r2R, rR);
    mat3_vec(k.R_RL, P1R, pR);
    FBUS_UNROLL
    for (int j = 0; j < 3; ++j) pR[j] += k.P_LR[j];
    const double c[3] = {r2L[1] * rR[2] - r2L[2] * rR[1], r2L[2] * rR[0] - r2L[0] * rR[2], r2L[0] * rR[1] - r2L[1] * rR[0]};
    const double d[3] = {pR[0] - P1L[0], pR[1] - P1L[1], pR[2] - P1L[2]};
    const double inv3 = 1.0 / det3_cols(c, r2L, rR);
    const double t1 = det3_cols(c, d, rR) * inv3;
    const double t2 = -det3_cols(c, r2L, d) * inv3;
    double P[3];
    FBUS_UNROLL
    for (int j = 0; j < 3; ++j) P[j] = 0.5 * (P1L[j] + t1 * r2L[j] + pR[j] + t2 * rR[j]);
    Pout[0] = -P[0]; Pout[1] = -P[1]; Pout[2] = P[2];
    return norm3(P);
}

// unit eigenvector of the smallest eigenvalue of a symmetric 3x3 (stands in for EigenSolver<Matrix3d>,
// vision.cpp:679-696; only this eigenvector is used and its sign is re-fixed by the caller).
// Closed-form eigenvalue (trigonometric) -> null vector of (M - lambda I) by the largest row cross product
// -> one Rayleigh-quotient polish.  The smallest eigenvalue of the corner scatter matrix is well separated
// (the two large ones are nearly equal for a square marker), so this is well conditioned.
FBUS_HD void null_vec(const double* M, double lam, double* v) {
    const double r0[3] = {M[0] - lam, M[1], M[2]}, r1[3] = {M[3], M[4] - lam, M[5]}, r2[3] = {M[6], M[7], M[8] - lam};
    const double c0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
    const double c1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
    const double c2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
    const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
    const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
    const double n2 = c2[0] * c2[0] + c2[1] * c2[1] + c2[2] * c2[2];
    double b0 = c0[0], b1 = c0[1], b2 = c0[2], nb = n0;
    if (n1 > nb) { b0 = c1[0]; b1 = c1[1]; b2 = c1[2]; nb = n1; }
    if (n2 > nb) { b0 = c2[0]; b1 = c2[1]; b2 = c2[2]; nb = n2; }
    const double inv = 1.0 / sqrt(nb);
    v[0] = b0 * inv; v[1] = b1 * inv; v[2] = b2 * inv;
}
FBUS_HD void smallest_eigvec_sym3(const double* M, double* z) {
    const double q = (M[0] + M[4] + M[8]) / 3.0;
    const double p1 = M[1] * M[1] + M[2] * M[2] + M[5] * M[5];
    const double a = M[0] - q, b = M[4] - q, c = M[8] - q;
    const double p2 = a * a + b * b + c * c + 2.0 * p1;
    const double p = sqrt(p2 / 6.0);
    double lam = q;
    if (p > 0.0) {
        const double ip = 1.0 / p;
        const double B0 = a * ip, B4 = b * ip, B8 = c * ip, B1 = M[1] * ip, B2 = M[2] * ip, B5 = M[5] * ip;
        double r = 0.5 * (B0 * (B4 * B8 - B5 * B5) - B1 * (B1 * B8 - B5 * B2) + B2 * (B1 * B5 - B4 * B2));
        r = r < -1.0 ? -1.0 : (r > 1.0 ? 1.0 : r);
        const double phi = acos(r) / 3.0;
        lam = q + 2.0 * p * cos(phi + 2.0943951023931953);  // smallest eigenvalue
    }
    double v[3];
    null_vec(M, lam, v);
    // Rayleigh-quotient polish
    double Mv[3];
    mat3_vec(M, v, Mv);
    lam = v[0] * Mv[0] + v[1] * Mv[1] + v[2] * Mv[2];
    null_vec(M, lam, z);
}

FBUS_HD double signum_ref(double x) { return x < 0 ? -1.0 : 1.0; }  // matrix_math.hpp:9-15

// R2 for one marker: C[12] = 4 corners x 3 -> p (corner 0 projected), q = Quaterniond([X Y Z])
// rod_s, rod_c = sin/cos of -REF_M_PI/4 (computed once on the host)
FBUS_HD void marker_pose(const double* C, double rod_s, double rod_c, double* p, double* q) {
    const double* c0 = C; const double* c1 = C + 3; const double* c2 = C + 6; const double* c3 = C + 9;
    double M[9];
    {
        double v[6][3];
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) {
            v[0][j] = c1[j] - c0[j]; v[1][j] = c2[j] - c0[j]; v[2][j] = c3[j] - c0[j];
            v[3][j] = c2[j] - c1[j]; v[4][j] = c3[j] - c1[j]; v[5][j] = c3[j] - c2[j];
        }
        FBUS_UNROLL
        for (int i = 0; i < 3; ++i)
            FBUS_UNROLL
            for (int j = i; j < 3; ++j) {
                double s = 0.0;
                FBUS_UNROLL
                for (int e = 0; e < 6; ++e) s += v[e][i] * v[e][j];
                M[i * 3 + j] = s;
                M[j * 3 + i] = s;
            }
    }
    double Z[3];
    smallest_eigvec_sym3(M, Z);
    {   // vision.cpp:697-709
        double s = 1.0;
        if (Z[2] > 0.1) s = -1.0;
        else if (Z[2] < -0.1) s = 1.0;
        else s = -signum_ref(c0[0]) * signum_ref(Z[0]);
        Z[0] *= s; Z[1] *= s; Z[2] *= s;
    }
    const double D = 0.25 * (Z[0] * (c0[0] + c1[0] + c2[0] + c3[0]) + Z[1] * (c0[1] + c1[1] + c2[1] + c3[1]) +
                             Z[2] * (c0[2] + c1[2] + c2[2] + c3[2]));
    double P1[3], P2[3], P4[3];
    {
        const double t1 = (Z[0] * c0[0] + Z[1] * c0[1] + Z[2] * c0[2]) - D;
        const double t2 = (Z[0] * c1[0] + Z[1] * c1[1] + Z[2] * c1[2]) - D;
        const double t4 = (Z[0] * c3[0] + Z[1] * c3[1] + Z[2] * c3[2]) - D;
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) {
            P1[j] = c0[j] - t1 * Z[j];
            P2[j] = c1[j] - t2 * Z[j];
            P4[j] = c3[j] - t4 * Z[j];
        }
    }
    double m[3];
    {
        double V12[3], V14[3];
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) { V12[j] = P2[j] - P1[j]; V14[j] = P4[j] - P1[j]; }
        const double i12 = 1.0 / norm3(V12), i14 = 1.0 / norm3(V14);
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) m[j] = V12[j] * i12 + V14[j] * i14;
    }
    // Rm = AngleAxisd(-REF_M_PI/4, Z).matrix() (Rodrigues) ; X = Rm*m/|m|
    double X[3];
    {
        const double sa[3] = {rod_s * Z[0], rod_s * Z[1], rod_s * Z[2]};
        const double ca[3] = {(1 - rod_c) * Z[0], (1 - rod_c) * Z[1], (1 - rod_c) * Z[2]};
        double Rm[9];
        double tmp;
        tmp = ca[0] * Z[1]; Rm[1] = tmp - sa[2]; Rm[3] = tmp + sa[2];
        tmp = ca[0] * Z[2]; Rm[2] = tmp + sa[1]; Rm[6] = tmp - sa[1];
        tmp = ca[1] * Z[2]; Rm[5] = tmp - sa[0]; Rm[7] = tmp + sa[0];
        Rm[0] = ca[0] * Z[0] + rod_c; Rm[4] = ca[1] * Z[1] + rod_c; Rm[8] = ca[2] * Z[2] + rod_c;
        double Rmm[3];
        mat3_vec(Rm, m, Rmm);
        const double im = 1.0 / norm3(m);
        FBUS_UNROLL
        for (int j = 0; j < 3; ++j) X[j] = Rmm[j] * im;
    }
    const double Y[3] = {Z[1] * X[2] - Z[2] * X[1], Z[2] * X[0] - Z[0] * X[2], Z[0] * X[1] - Z[1] * X[0]};
    const double R[9] = {X[0], Y[0], Z[0], X[1], Y[1], Z[1], X[2], Y[2], Z[2]};
    R2q(R, q);
    p[0] = P1[0]; p[1] = P1[1]; p[2] = P1[2];
}

}  // namespace fbus
