// tmem_probe.cu -- microbenchmark: can tensor memory (TMEM) serve as per-thread scratch for a non-GEMM FP64 kernel?
// Measures, per SM, the sustained bytes/clk of
//   (a) tcgen05.ld 32x32b (.x2/.x8/.x16/.x32) from the warp's own lane partition,
//   (b) tcgen05.st 32x32b,
//   (c) ld.shared.f64 from a conflict-free [entry][thread] layout (what the window kernel does today),
//   (d) (a) and (c) issued by different warps at the same time (are the two paths additive?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_probe tmem_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t tmem_alloc_all(uint32_t* slot) {  // whole CTA; warp 0 allocates 512 columns
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    return *slot;
}
__device__ __forceinline__ void tmem_free_all(uint32_t addr) {
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr));
}
__device__ __forceinline__ void ld2(uint32_t a, double& v) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
    v = __hiloint2double((int)r1, (int)r0);
}
__device__ __forceinline__ void st2(uint32_t a, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"((uint32_t)__double2loint(v)), "r"((uint32_t)__double2hiint(v)));
}
__device__ __forceinline__ void ld16(uint32_t a, double* v) {  // 8 doubles
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __hiloint2double((int)r[2 * i + 1], (int)r[2 * i]);
}
__device__ __forceinline__ void st16(uint32_t a, const double* v) {  // 8 doubles
    uint32_t r[16];
#pragma unroll
    for (int i = 0; i < 8; ++i) { r[2 * i] = (uint32_t)__double2loint(v[i]); r[2 * i + 1] = (uint32_t)__double2hiint(v[i]); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
                   "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]));
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: TMEM ld.x2 ; 1: TMEM ld.x16 ; 2: TMEM st.x2 ; 3: LDS.64 ; 4: warps 0-3 TMEM ld.x16, warps 4-7 LDS.64 ; 5: round trip check
__global__ void __launch_bounds__(256) probe(int mode, int iters, double* out, long long* cyc) {
    extern __shared__ double sm[];
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t base = tmem_alloc_all(&slot);
    const uint32_t mine = base + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's lane partition
    // initialise: column pair c holds the double (lane*1000 + c) ; warps 4-7 use columns 256..511
    const uint32_t col0 = (warp >= 4) ? 256u : 0u;
    for (int c = 0; c < 128; ++c) st2(mine + col0 + 2 * c, (double)(threadIdx.x * 1000 + c));
    wait_st();
    for (int e = 0; e < 64; ++e) sm[e * 256 + threadIdx.x] = (double)(threadIdx.x * 1000 + e);
    __syncthreads();
    double acc = 0.0;
    long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            double v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) ld2(mine + col0 + 2 * ((it * 16 + c) & 127), v[c]);
            wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) acc += v[c];
        }
    } else if (mode == 1) {
        for (int it = 0; it < iters; ++it) {
            double v[16];
            ld16(mine + col0 + 16 * ((2 * it) & 15), v);
            ld16(mine + col0 + 16 * ((2 * it + 1) & 15), v + 8);
            wait_ld();
#pragma unroll
            for (int c = 0; c < 16; ++c) acc += v[c];
        }
    } else if (mode == 2) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int c = 0; c < 16; ++c) st2(mine + col0 + 2 * ((it * 16 + c) & 127), acc + c);
            acc += 1.0;
        }
        wait_st();
    } else if (mode == 3) {
        for (int it = 0; it < iters; ++it) {
            double v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = sm[((it * 16 + c) & 63) * 256 + threadIdx.x];
#pragma unroll
            for (int c = 0; c < 16; ++c) acc += v[c];
        }
    } else if (mode == 4) {
        if (warp < 4) {
            for (int it = 0; it < iters; ++it) {
                double v[16];
                ld16(mine + 16 * ((2 * it) & 15), v);
                ld16(mine + 16 * ((2 * it + 1) & 15), v + 8);
                wait_ld();
#pragma unroll
                for (int c = 0; c < 16; ++c) acc += v[c];
            }
        } else {
            for (int it = 0; it < iters; ++it) {
                double v[16];
#pragma unroll
                for (int c = 0; c < 16; ++c) v[c] = sm[((it * 16 + c) & 63) * 256 + threadIdx.x];
#pragma unroll
                for (int c = 0; c < 16; ++c) acc += v[c];
            }
        }
    } else if (mode == 6) {  // st.x16
        for (int it = 0; it < iters; ++it) {
            double v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = acc + c;
            st16(mine + col0 + 16 * ((2 * it) & 15), v);
            st16(mine + col0 + 16 * ((2 * it + 1) & 15), v);
            acc += 1.0;
        }
        wait_st();
    } else if (mode == 7) {  // block pattern at unaligned columns: 9 doubles at column 18*b = ld.x16 + ld.x2, 9 FMAs, st back
        for (int it = 0; it < iters; ++it) {
            const uint32_t a = mine + col0 + 18 * (uint32_t)(it % 13);
            double v[9];
            ld16(a, v);
            ld2(a + 16, v[8]);
            wait_ld();
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = v[c] * 1.0000001 + 1e-9;
            st16(a, v);
            st2(a + 16, v[8]);
            acc += v[0];
        }
        wait_st();
    } else if (mode == 8) {  // same pattern through shared memory
        for (int it = 0; it < iters; ++it) {
            const int b = (it % 7) * 9;
            double v[9];
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = sm[(b + c) * 256 + threadIdx.x];
#pragma unroll
            for (int c = 0; c < 9; ++c) v[c] = v[c] * 1.0000001 + 1e-9;
#pragma unroll
            for (int c = 0; c < 9; ++c) sm[(b + c) * 256 + threadIdx.x] = v[c];
            acc += v[0];
        }
    } else if (mode == 9) {  // unaligned-column correctness: write 9 doubles at column 18*5+2, read back
        double v[9], w[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) v[c] = threadIdx.x * 7.0 + c;
        st16(mine + col0 + 92, v); st2(mine + col0 + 108, v[8]);
        wait_st();
        ld16(mine + col0 + 92, w); ld2(mine + col0 + 108, w[8]);
        wait_ld();
        for (int c = 0; c < 9; ++c) acc += fabs(w[c] - v[c]);
    } else {  // round trip: every thread must read back exactly what it wrote
        double v;
        ld2(mine + col0 + 2 * 77, v);
        wait_ld();
        acc = v - (double)(threadIdx.x * 1000 + 77);
    }
    long long t1 = clock64();
    out[blockIdx.x * 256 + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    tmem_free_all(base);
}

int main() {
    const int nblk = 148, iters = 4096;
    double* out; long long* cyc;
    CK(cudaMalloc(&out, nblk * 256 * sizeof(double)));
    CK(cudaMalloc(&cyc, nblk * sizeof(long long)));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 256 * 8));
    const char* names[] = {"TMEM ld 32x32b.x2 (8 warps)", "TMEM ld 32x32b.x16 (8 warps)", "TMEM st 32x32b.x2 (8 warps)", "LDS.64 (8 warps)",
                           "TMEM ld.x16 (warps 0-3) + LDS.64 (warps 4-7)", "round trip", "TMEM st 32x32b.x16 (8 warps)",
                           "TMEM block r/w (ld.x16+x2, 9 FMA, st.x16+x2)", "smem block r/w (9 LDS, 9 FMA, 9 STS)", "unaligned column round trip"};
    for (int mode = 0; mode < 10; ++mode) {
        probe<<<nblk, 256, 64 * 256 * 8>>>(mode, iters, out, cyc);
        CK(cudaDeviceSynchronize());
        long long h[148]; double ho[256];
        CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ho, out, sizeof(ho), cudaMemcpyDeviceToHost));
        double avg = 0; for (int i = 0; i < nblk; ++i) avg += (double)h[i]; avg /= nblk;
        if (mode == 7 || mode == 8) { printf("%-50s %10.0f clk  %7.1f B/clk/SM r+w (%.1f clk per 9-double block per warp)\n", names[mode], avg, (double)iters * 9 * 8 * 256 * 2 / avg, avg / iters); continue; }
        if (mode == 5 || mode == 9) { double mx = 0; for (int i = 0; i < 256; ++i) mx = fmax(mx, fabs(ho[i])); printf("%-50s max |readback - written| = %g\n", names[mode], mx); continue; }
        const double bytes = (double)iters * 16 * 8 * 256;  // per CTA (= per SM)
        printf("%-50s %10.0f clk  %7.1f B/clk/SM  (%.2f clk per warp-level 8-byte access)\n", names[mode], avg, bytes / avg, avg / (iters * 16.0));
    }
    return 0;
}
