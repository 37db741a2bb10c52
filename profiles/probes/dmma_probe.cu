// dmma_probe.cu -- does the FP64 tensor-core path (mma.sync.aligned.m8n8k4 f64, "DMMA") beat the FP64 CUDA cores (DFMA) on this
// part, in peak and on the filter's hot contraction?  north_star keeps tensor cores "only if ncu shows they beat CUDA-core FP64
// on these 15x15/6x15 contractions".
//
//   1. peak: independent DMMA chains vs independent DFMA chains, all SMs, 8 warps per SM            -> TFLOP/s each
//   2. the propagate's product M = F P (18x18, F = I + six small blocks, filter.cpp:598-604) for a batch of filters
//        a. structured DFMA: one lane per column, 33 FMA per column (what the kernels execute: 594 FMA per product)
//        b. dense DMMA: one warp per filter, F and P padded 18 -> 24, 3x3 output tiles x 6 k-steps = 54 mma.m8n8k4
//      -> products per second each (both register-resident, no memory traffic: the pure arithmetic comparison)
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o dmma_probe dmma_probe.cu ; run on a B200.
// ncu counters for the same kernels: see profiles/probes/RESULTS.md.
#include <cuda_runtime.h>

#include <cstdio>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int ILP>
__global__ void __launch_bounds__(256) dmma_peak(double* out, int iters, double seed) {
    double c[ILP][2];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = seed * i; c[i][1] = seed + i; }
    const double a = seed * 1.0000001, b = 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void __launch_bounds__(256) dfma_peak(double* out, int iters, double seed) {
    double c[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) c[i] = seed * i;
    const double a = seed * 1.0000001, b = 0.9999999;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (a) structured: y = F x on one column per lane (33 DFMA), `reps` products back to back (each product = 18 column applications;
//     a warp of 32 lanes covers 32 columns per pass, so 18 passes of a warp = 32 products)
__global__ void __launch_bounds__(256) fp_structured(double* out, int reps, double seed) {
    double x[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) x[i] = seed + i * 1e-3 + threadIdx.x * 1e-6;
    double A[9], Bm[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { A[i] = 1e-3 * (i + 1) * seed; Bm[i] = -5e-3 + 1e-4 * i; }
    const double dt = 5e-3, u0 = 1e-4, u1 = -2e-4, u2 = 3e-4;
    for (int r = 0; r < reps; ++r) {
        double y[9];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double s = x[i];
            s += dt * x[3 + i];
            y[i] = s;
            double t = x[3 + i];
            t += dt * x[15 + i];
#pragma unroll
            for (int c = 0; c < 3; ++c) { t += A[i * 3 + c] * x[6 + c]; t += Bm[i * 3 + c] * x[9 + c]; }
            y[3 + i] = t;
        }
        double t0 = x[6], t1 = x[7], t2 = x[8];
        t0 -= dt * x[12]; t1 -= dt * x[13]; t2 -= dt * x[14];
        t0 += u2 * x[7]; t0 -= u1 * x[8];
        t1 += u0 * x[8]; t1 -= u2 * x[6];
        t2 += u1 * x[6]; t2 -= u0 * x[7];
        y[6] = t0; y[7] = t1; y[8] = t2;
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = y[i];
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 18; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// (b) dense: C(24x24) = A(24x24) B(24x24) with mma.m8n8k4, one warp per product, fragments in registers
__global__ void __launch_bounds__(256) fp_dmma(double* out, int reps, double seed) {
    double a[3][6], b[6][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 6; ++k) { a[i][k] = (i * 2 == k ? 1.0 : 1e-3 * seed) + threadIdx.x * 1e-9; b[k][i] = seed + 1e-3 * (k + i); }
    double acc = 0;
    for (int r = 0; r < reps; ++r) {
        double c[3][3][2];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { c[i][j][0] = 0; c[i][j][1] = 0; }
#pragma unroll
        for (int k = 0; k < 6; ++k)
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) dmma(c[i][j][0], c[i][j][1], a[i][k], b[k][j]);
        // the result becomes the next B operand (P <- F P), as in the filter: keeps the chain honest
#pragma unroll
        for (int j = 0; j < 3; ++j)
#pragma unroll
            for (int i = 0; i < 3; ++i) { b[2 * i][j] = c[i][j][0]; b[2 * i + 1][j] = c[i][j][1]; acc += c[i][j][0]; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <class F>
static double time_ms(F launch) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount, blocks = sms, threads = 256;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads * 4);
    const int iters = 20000;
    {
        const double ms = time_ms([&] { dfma_peak<8><<<blocks, threads>>>(out, iters, 1.0); });
        const double flop = 2.0 * 8 * (double)iters * blocks * threads;
        std::printf("DFMA peak      : %8.2f TFLOP/s  (%d SMs x 8 warps, 8 independent chains per thread)\n", flop / ms / 1e9, sms);
    }
    {
        const double ms = time_ms([&] { dmma_peak<8><<<blocks, threads>>>(out, iters, 1.0); });
        const double flop = 512.0 * 8 * (double)iters * blocks * (threads / 32);  // m8n8k4: 8*8*4 FMA per warp instruction
        std::printf("DMMA peak      : %8.2f TFLOP/s  (mma.sync.m8n8k4.f64, 8 independent accumulators per warp)\n", flop / ms / 1e9);
    }
    {
        const int reps = 18 * 2000;  // 2000 passes over 18 columns per lane
        const double ms = time_ms([&] { fp_structured<<<blocks, threads>>>(out, reps, 1.0); });
        const double products = (double)reps / 18.0 * blocks * threads;  // one product = 18 column applications
        std::printf("F*P structured : %8.3e products/s  (33 DFMA per column, 594 per product; %.2f TFLOP/s executed)\n", products / (ms * 1e-3),
                    products * 594 * 2 / ms / 1e9);
    }
    {
        const int reps = 2000;
        const double ms = time_ms([&] { fp_dmma<<<blocks, threads>>>(out, reps, 1.0); });
        const double products = (double)reps * blocks * (threads / 32);
        std::printf("F*P dense DMMA : %8.3e products/s  (18 -> 24 padded, 54 mma.m8n8k4 per product; %.2f TFLOP/s executed)\n", products / (ms * 1e-3),
                    products * 54 * 512 / ms / 1e9);
    }
    cudaFree(out);
    return 0;
}
