// dfma_operands.cu -- DFMA issue rate vs operand pattern on sm_100a: does a stream of FMAs whose three 64-bit source operands
// are all different registers sustain the same rate as one that reuses an operand (operand reuse cache / register banks)?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_operands dfma_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256, 1) k(double* out, long long* cyc, int iters, double seed) {
    double x[16], y[16], acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = seed + threadIdx.x * 1e-9 + i * 1e-3; y[i] = 1.0 - x[i] * 1e-3; acc[i] = i; }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {  // shared multiplier: acc[i] += x[0] * y[i]   (one operand reused by consecutive FMAs)
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fma(x[r], y[i], acc[i]);
        } else if (MODE == 1) {  // all different: acc[i] += x[i] * y[(i + r) % 16]
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fma(x[i], y[(i + r * 5 + 1) & 15], acc[i]);
        } else {  // 3x3 block product pattern as in the filter: C[i][j] += A[i][k] * B[k][j]
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int i = 0; i < 3; ++i)
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int kk = 0; kk < 3; ++kk) acc[i * 3 + j] = fma(x[i * 3 + kk], y[kk * 3 + j], acc[i * 3 + j]);
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int nfma, double* out, long long* cyc) {
    const int iters = 4096;
    k<MODE><<<148, 256>>>(out, cyc, iters, 0.5);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    const double per_smsp = 2.0 * iters * nfma;  // two warps per scheduler
    printf("%-58s %.2f clk per DFMA per scheduler (2.00 = pipe peak) -> %.0f %% of peak\n", name, avg / per_smsp, 200.0 * per_smsp / avg);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&cyc, 148 * 8);
    run<0>("acc[i] += x[r] * y[i] (multiplier shared by 16 FMAs)", 64, out, cyc);
    run<1>("acc[i] += x[i] * y[j] (three different registers each)", 64, out, cyc);
    run<2>("3x3 block product C += A B", 108, out, cyc);
    return 0;
}
