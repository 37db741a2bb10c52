// fbus_kernel_lane.cuh -- K12 for SMALL batches: the fused window kernel with NINE LANES PER FILTER.
//
// Why: with one thread per filter a batch of a few thousand filters cannot fill 148 SMs -- 4 096 filters are 128 warps of
// covariance work, each a serial chain of ~660 dependent-ish FP64 instructions per IMU sample, and the throughput is the
// per-filter latency times the batch (0.47e6 filter-steps/s per filter whatever the CTA shape, profiles/probes/RESULTS.md).
// Here the covariance of ONE filter is spread over nine lanes of a warp (three filters per warp, 27 of 32 lanes busy):
//
//   lane l (0..8) of a filter owns the FULL columns l and l+9 of the symmetric 18x18 covariance in registers (36 doubles).
//
//   propagate  P <- F P F^T + Qbar   (FILTER::UpdateCovariance, filter.cpp:588-616) in two applications of the same
//     column operator  y = F x  (33 FMA: rows 0..8 change, F = I + the six small blocks of filter.cpp:598-604):
//       A.  M[:, j] = F P[:, j]            for both own columns                      (local)
//       T.  the top nine entries of both columns go to shared memory, lane l reads ROW l of M back   (18 st + 18 ld)
//       B.  P'[:, l] = F (row l of M)^T    -- by symmetry of P this is column l of F P F^T; its rows 9..17 are
//           copies of M[l][9..17], i.e. bit-identical to what the lanes owning columns 9..17 hold, and column l+9 of P'
//           is M[:, l+9] as it stands.
//     Only the top-left 9x9 is evaluated in two association orders (P'[i][j] by lane j, P'[j][i] by lane i); it is
//     symmetrised once per frame (the reference symmetrises after every step, filter.cpp:614-615; the difference is at the
//     rounding level and bounded by the window length).
//   update     (FILTER::ObservationUpdate, filter.cpp:622-739): the NOMINAL warp (one lane per filter, 32 filters) runs
//     the state-only prologue of the structured update (update_prologue in fbus_math.cuh: predicted measurement, Hs, S,
//     Cholesky, C = Hs^T S^-1 Hs = Lc Lc^T, y) from the 6x6 block P6 the lanes publish; the lanes then form their two
//     columns of Z = Lc^T G locally (G = rows {0,1,2,6,7,8} of P = six entries of each own column), exchange the Z
//     columns through shared memory and sweep  P -= Z^T Z  (exactly symmetric by construction) and  dx = Z^T y.
//
// The nominal warp is the one of the two-warp kernel (fbus_kernel_split.cuh, nominal_role<.., LANE = true>): detection scan,
// F6b / F5, F2, the streams, one IMU sample ahead of the covariance lanes through the same two-deep coefficient ring.
// CTA = 32 filters = 11 covariance warps + 1 nominal warp = 384 threads, every barrier CTA-wide.
//
// This is the layout north_star sketches ("one warp or warp-group per filter with P held in registers and shared memory");
// the thread-per-filter kernels stay for every batch that fills the GPU, where they execute 1.6x fewer FP64
// warp-instructions per filter-step.
#pragma once

#include "fbus_kernel_split.cuh"

namespace fbus {

constexpr int LANE_TS = 19;               // row stride of the per-filter transpose scratch (18 + 1: bank-conflict-free row reads)
constexpr int LANE_T = 9 * LANE_TS;       // doubles of transpose scratch per filter
constexpr size_t LANE_SMEM = (size_t)(LX_TOTAL * 32 + LANE_T * 32) * sizeof(double);

// y = F x on one column (rows 0..8 change):  filter.cpp:598-604 with A = -R[a]x dt, B = -R dt and W = F[theta,theta] - I in
// full (-[w]x dt for the C++ semantics, expm(-[w]x dt) - I for the MATLAB ones: the nominal warp decides, the lanes just apply it)
__device__ __forceinline__ void lane_apply_F(double* x, const double* A, const double* Bm, const double* W, double dt) {
    double y[9];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double s = x[i];
        s += dt * x[3 + i];
        y[i] = s;
        double t = x[3 + i];
        t += dt * x[15 + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            t += A[i * 3 + c] * x[6 + c];
            t += Bm[i * 3 + c] * x[9 + c];
        }
        y[3 + i] = t;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        double t = x[6 + i];
        t -= dt * x[12 + i];
#pragma unroll
        for (int c = 0; c < 3; ++c) t += W[i * 3 + c] * x[6 + c];
        y[6 + i] = t;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) x[i] = y[i];
}

// COVARIANCE lanes: warp cw (0..10) holds filters 3cw .. 3cw+2 of the CTA, lane = 9*g + l
template <bool JOSEPH>
__device__ __forceinline__ void lane_cov_role(const WinParams& prm, const DevConsts& k, double* smem, SplitShared& sh, int32_t (*sflag)[32],
                                              int cw, int lane) {
    constexpr int NT = LANE_NT;
    const size_t B = prm.B;
    const int g = lane / 9, l = lane - 9 * g;
    const int f0 = cw * 3 + g;
    const bool act = (g < 3) && (f0 < 32);
    const int f = act ? f0 : 31;  // idle lanes shadow filter 31 for their (unused) reads and never write
    const size_t b0 = (size_t)blockIdx.x * 32 + f;
    const bool live = act && b0 < B;
    const size_t b = (b0 < B) ? b0 : B - 1;
    double* const X = smem + f;                                   // exchange area, entry stride 32
    double* const T = smem + (size_t)LX_TOTAL * 32 + (size_t)f * LANE_T;  // transpose scratch of this filter
    // own columns l and l + 9 (full 18 entries each) from the packed upper triangle
    double c0[18], c1[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) {
        const int a0 = i < l ? i : l, b0i = i < l ? l : i;            // (min, max) of (i, l)
        const int a1 = i < l + 9 ? i : l + 9, b1i = i < l + 9 ? l + 9 : i;
        c0[i] = prm.P[(size_t)(a0 * NX - (a0 * (a0 - 1)) / 2 + (b0i - a0)) * B + b];
        c1[i] = prm.P[(size_t)(a1 * NX - (a1 * (a1 - 1)) / 2 + (b1i - a1)) * B + b];
    }
    // process noise on the own diagonal entries (GammaQGamma^T is diagonal, filter.hpp:108-125): column l gets it in row l,
    // column l+9 in row l+9
    const double q0 = (l >= 3 && l < 6) ? k.Qd[0] : (l >= 6 ? k.Qd[1] : 0.0);
    const double q1 = (l < 3) ? k.Qd[2] : (l < 6 ? k.Qd[3] : 0.0);
    for (uint32_t w = prm.w0; w < prm.w1; ++w) {
        cta_bar<NT>();  // (a) the nominal warp has posted the IMU range
        const int fp = (int)((w - prm.w0) & 1u);
        const uint32_t lo = sh.lo_hi[fp][0][0], hi = sh.lo_hi[fp][1][0];
        bool touched = false;
        for (uint32_t i = lo; i < hi; ++i) {
            cta_bar<NT>();  // record (i) is complete; the nominal warp moves on to sample i+1
            const int slot = (int)((i - lo) & 1u);
            const bool valid = act && sflag[slot][f] != 0;
            const double* rec = X + (size_t)slot * LX_REC * 32;
            double A[9], Bm[9], Wm[9];
#pragma unroll
            for (int e = 0; e < 9; ++e) { A[e] = rec[(size_t)e * 32]; Bm[e] = rec[(size_t)(9 + e) * 32]; Wm[e] = rec[(size_t)(18 + e) * 32]; }
            const double dt = rec[(size_t)27 * 32];
            if (valid) {  // A: M = F P on both columns
                lane_apply_F(c0, A, Bm, Wm, dt);
                lane_apply_F(c1, A, Bm, Wm, dt);
#pragma unroll
                for (int r = 0; r < 9; ++r) {  // T: publish the rows of M that change
                    T[r * LANE_TS + l] = c0[r];
                    T[r * LANE_TS + 9 + l] = c1[r];
                }
            }
            __syncwarp();
            if (valid) {  // B: column l of F P F^T = F (row l of M)^T
#pragma unroll
                for (int c = 0; c < 18; ++c) c0[c] = T[l * LANE_TS + c];
                lane_apply_F(c0, A, Bm, Wm, dt);
#pragma unroll
                for (int r = 3; r < 9; ++r) c0[r] += (r == l) ? q0 : 0.0;
#pragma unroll
                for (int r = 9; r < 15; ++r) c1[r] += (r == l + 9) ? q1 : 0.0;
                touched = true;
            }
            __syncwarp();  // the row reads are done before the next sample's publish
        }
        if (__any_sync(0xffffffffu, touched)) {  // symmetrise the top-left 9x9 (the only part evaluated in two orders)
            if (touched) {
#pragma unroll
                for (int r = 0; r < 9; ++r) T[r * LANE_TS + l] = c0[r];
            }
            __syncwarp();
            if (touched) {
#pragma unroll
                for (int r = 0; r < 9; ++r) c0[r] = 0.5 * (c0[r] + T[l * LANE_TS + r]);
            }
            __syncwarp();
        }
        cta_bar<NT>();  // (r) update requests posted
        if (sh.any_upd[0]) {
            const bool req = act && sflag[2][f] != 0;
            // the 6x6 block of the p / theta rows and columns for the prologue: lanes l in {0,1,2,6,7,8} own its columns
            if (req && (l < 3 || l >= 6)) {
                const int ci = l < 3 ? l : l - 3;
#pragma unroll
                for (int m = 0; m < 6; ++m) X[(size_t)(LX_P6 + m * 6 + ci) * 32] = c0[m < 3 ? m : m + 3];
            }
            cta_bar<NT>();  // (p) P6 published
            cta_bar<NT>();  // (c) the nominal warp has posted Lc and y
            // own columns of the factor Z with (I-KH)P = P - Z^T Z, dx = Z^T y:
            //   default  Z = X G   (7 x 18), X = L^-1 Hs, y = z = L^-1 r      (S = L L^T)
            //   Joseph   Z = Lc^T G (6 x 18), Lc Lc^T = C_J, y = Lc^-1 u       (update_prologue)
            constexpr int NZ = JOSEPH ? 6 : 7;
            double z0[NZ], z1[NZ];
            if (req) {
                if constexpr (JOSEPH) {
                    double Cm[21];
#pragma unroll
                    for (int c = 0; c < 21; ++c) Cm[c] = X[(size_t)(LX_CM + c) * 32];
#pragma unroll
                    for (int kz = 0; kz < 6; ++kz) {
                        double s0 = Cm[kz * (kz + 1) / 2 + kz] * c0[kz < 3 ? kz : kz + 3];
                        double s1 = Cm[kz * (kz + 1) / 2 + kz] * c1[kz < 3 ? kz : kz + 3];
#pragma unroll
                        for (int m = kz + 1; m < 6; ++m) {
                            s0 += Cm[m * (m + 1) / 2 + kz] * c0[m < 3 ? m : m + 3];
                            s1 += Cm[m * (m + 1) / 2 + kz] * c1[m < 3 ? m : m + 3];
                        }
                        z0[kz] = s0;
                        z1[kz] = s1;
                    }
                } else {
#pragma unroll
                    for (int kz = 0; kz < 7; ++kz) {
                        double s0 = 0.0, s1 = 0.0;
#pragma unroll
                        for (int m = 0; m < 6; ++m) {
                            const double x = X[(size_t)(LX_SCR + kz * 6 + m) * 32];
                            s0 += x * c0[m < 3 ? m : m + 3];
                            s1 += x * c1[m < 3 ? m : m + 3];
                        }
                        z0[kz] = s0;
                        z1[kz] = s1;
                    }
                }
                double d0 = 0.0, d1 = 0.0;  // dx = Z^T y for the own columns
#pragma unroll
                for (int kz = 0; kz < NZ; ++kz) {
                    T[kz * LANE_TS + l] = z0[kz];
                    T[kz * LANE_TS + 9 + l] = z1[kz];
                    const double yk = X[(size_t)(LX_Y + kz) * 32];
                    d0 += yk * z0[kz];
                    d1 += yk * z1[kz];
                }
                X[(size_t)(LX_DX + l) * 32] = d0;
                X[(size_t)(LX_DX + 9 + l) * 32] = d1;
            }
            __syncwarp();  // the filter's Z columns are in its scratch
            if (req) {
#pragma unroll
                for (int r = 0; r < 18; ++r) {  // P -= Z^T Z on both columns (exactly symmetric: the same products in the same order on both sides)
                    double v0 = c0[r], v1 = c1[r];
#pragma unroll
                    for (int kz = 0; kz < NZ; ++kz) {
                        const double zr = T[kz * LANE_TS + r];
                        v0 -= zr * z0[kz];
                        v1 -= zr * z1[kz];
                    }
                    c0[r] = v0;
                    c1[r] = v1;
                }
            }
            __syncwarp();
            cta_bar<NT>();  // (d) dx posted
        }
    }
    if (live) {  // upper triangle back: column j is stored by its owner for rows i <= j
#pragma unroll
        for (int i = 0; i < 18; ++i) {
            if (i <= l) prm.P[(size_t)(i * NX - (i * (i - 1)) / 2 + (l - i)) * B + b] = c0[i];
            if (i <= l + 9) prm.P[(size_t)(i * NX - (i * (i - 1)) / 2 + (l + 9 - i)) * B + b] = c1[i];
        }
    }
}

template <bool JOSEPH, bool IMU32, bool MATLAB = false>
__global__ void __launch_bounds__(LANE_NT, 1) ekf_window_lane_kernel(const __grid_constant__ WinParams prm, const __grid_constant__ DevConsts k) {
    extern __shared__ double smem[];
    __shared__ SplitShared sh;
    __shared__ int32_t sflag[3][32];
    const int wi = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (wi < 11) {
        lane_cov_role<JOSEPH>(prm, k, smem, sh, sflag, wi, lane);
    } else {
        const size_t b0 = (size_t)blockIdx.x * 32 + lane;
        const bool live = b0 < prm.B;
        nominal_role<32, false, IMU32, true, JOSEPH, MATLAB>(prm, k, smem, sh, sflag, lane, live ? b0 : prm.B - 1, live);
    }
}

}  // namespace fbus
