// math_probe.cu -- accuracy of the call-free device math helpers (fbus_math.cuh) against the CUDA library functions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I fbus_ekf_b200/csrc -o math_probe profiles/probes/math_probe.cu
#include <cstdio>
#include <cmath>
#include <cstdint>
#include "fbus_math.cuh"
__device__ double ulps(double a, double b) { return b == 0.0 ? fabs(a) / 4.9e-324 : fabs(a - b) / (fabs(b) * 2.220446049250313e-16); }
__global__ void k(double* out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double m[4] = {0, 0, 0, 0};
    for (int j = i; j < n; j += gridDim.x * blockDim.x) {
        const double u = (j + 0.5) / n;
        const double ang = (u - 0.5) * 200.0;                 // [-100, 100] rad
        const double small = (u - 0.5) * 1e-2;                // the range the filter actually uses
        const double pos = exp((u - 0.5) * 60.0);             // 1e-13 .. 1e13
        double s, c, s0, c0;
        fbus::sincos_d(ang, &s, &c); sincos(ang, &s0, &c0);
        m[0] = fmax(m[0], fmax(fabs(s - s0), fabs(c - c0)) / 2.220446049250313e-16);
        fbus::sincos_d(small, &s, &c); sincos(small, &s0, &c0);
        m[1] = fmax(m[1], fmax(ulps(s, s0), ulps(c, c0)));
        m[2] = fmax(m[2], ulps(fbus::rsqrt_d(pos), rsqrt(pos)));
        m[3] = fmax(m[3], ulps(fbus::rcp_d(pos), 1.0 / pos));
    }
    for (int q = 0; q < 4; ++q) out[i * 4 + q] = m[q];
}
int main() {
    const int T = 148 * 256;
    double* d; cudaMalloc(&d, T * 4 * 8);
    k<<<148, 256>>>(d, 1 << 24);
    double* h = new double[T * 4];
    cudaMemcpy(h, d, T * 4 * 8, cudaMemcpyDeviceToHost);
    double m[4] = {0, 0, 0, 0};
    for (int i = 0; i < T; ++i) for (int q = 0; q < 4; ++q) m[q] = fmax(m[q], h[i * 4 + q]);
    printf("sincos_d on [-100,100] rad: max abs error %.2f eps ; on [-5e-3,5e-3]: %.2f ulp ; rsqrt_d: %.2f ulp ; rcp_d: %.2f ulp (vs CUDA sincos / rsqrt / division)\n", m[0], m[1], m[2], m[3]);
    return 0;
}
