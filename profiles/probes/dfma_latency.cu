// dfma_latency.cu -- dependent-issue latency and per-warp throughput of DFMA on sm_100a, as a function of the number of
// independent chains per warp (ILP) and of warps per scheduler.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int ILP>
void run(int threads, double* out, long long* cyc) {
    const int iters = 2048;
    chain<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += (double)h[i]; avg /= 148;
    const double n = (double)iters * 8 * ILP;  // DFMA per warp
    printf("warps/SM %2d (per scheduler %.1f)  ILP %2d : %6.2f clk per DFMA per warp, %5.2f clk per dependent step, pipe %.0f%%\n", threads / 32, threads / 128.0, ILP,
           avg / n, avg / (iters * 8.0), 100.0 * n * (threads / 128.0 < 1 ? 1 : threads / 128.0) * 2.0 / avg);
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 148 * 8);
    for (int threads : {32, 128, 256}) {
        run<1>(threads, out, cyc); run<2>(threads, out, cyc); run<4>(threads, out, cyc); run<6>(threads, out, cyc); run<8>(threads, out, cyc); run<12>(threads, out, cyc); run<16>(threads, out, cyc);
    }
    return 0;
}
