// fbus_tmem.cuh -- the packed covariance of a filter in TENSOR MEMORY (sm_100a), one TMEM lane per filter.
//
// The window kernel is bound by on-chip bandwidth as much as by the FP64 pipe: every propagate step reads ~220 and
// writes ~90 doubles of P per filter.  Shared memory moves 128 B/clk/SM.  Tensor memory (256 KB per SM, 128 lanes x 512
// 32-bit columns) is reachable from ordinary warps with tcgen05.ld / tcgen05.st; measured on this pool's B200s with
// profiles/probes/tmem_probe.cu: 219 B/clk/SM for 8-double loads, 301 B/clk/SM for 8-double stores, and 329 B/clk/SM
// for the read-modify-write of 3x3 blocks that the covariance algebra consists of -- 2.6x shared memory.  With the
// 32x32b shape, thread t of warp w owns lane 32*(w%4)+t: a private 2 KB scratch per thread, exactly the access pattern
// of "one filter per thread".  No tensor-core instruction is involved; TMEM is used as a second register file.
//
// Layout per lane: the 21 upper blocks (bi <= bj) of the 6x6 block matrix, each a full row-major 3x3 (diagonal blocks
// stored with both triangles), block (bi,bj) at column 18*(bj(bj+1)/2 + bi): 378 of the 512 columns.
#pragma once

#include "fbus_math.cuh"

#ifndef FBUS_TMEM_WIDE_ST
#define FBUS_TMEM_WIDE_ST 1  // measured: 8.64e9 (wide) vs 8.57e9 (nine 2-column stores) filter-steps/s
#endif

namespace fbus {

__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_ld1(uint32_t a, double& v) {  // one double = two columns
    asm volatile("{\n\t.reg .b32 t<2>;\n\ttcgen05.ld.sync.aligned.32x32b.x2.b32 {t0, t1}, [%1];\n\tmov.b64 %0, {t0, t1};\n\t}" : "=d"(v) : "r"(a) : "memory");
}
__device__ __forceinline__ void tm_st1(uint32_t a, double v) {
    asm volatile("{\n\t.reg .b32 t<2>;\n\tmov.b64 {t0, t1}, %1;\n\ttcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {t0, t1};\n\t}" ::"r"(a), "d"(v) : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t a, double* v) {  // eight doubles = sixteen columns
    asm volatile(
        "{\n\t.reg .b32 t<16>;\n\t"
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15}, [%8];\n\t"
        "mov.b64 %0, {t0, t1};\n\tmov.b64 %1, {t2, t3};\n\tmov.b64 %2, {t4, t5};\n\tmov.b64 %3, {t6, t7};\n\t"
        "mov.b64 %4, {t8, t9};\n\tmov.b64 %5, {t10, t11};\n\tmov.b64 %6, {t12, t13};\n\tmov.b64 %7, {t14, t15};\n\t}"
        : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]), "=d"(v[4]), "=d"(v[5]), "=d"(v[6]), "=d"(v[7])
        : "r"(a)
        : "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t a, const double* v) {
    asm volatile(
        "{\n\t.reg .b32 t<16>;\n\t"
        "mov.b64 {t0, t1}, %1;\n\tmov.b64 {t2, t3}, %2;\n\tmov.b64 {t4, t5}, %3;\n\tmov.b64 {t6, t7}, %4;\n\t"
        "mov.b64 {t8, t9}, %5;\n\tmov.b64 {t10, t11}, %6;\n\tmov.b64 {t12, t13}, %7;\n\tmov.b64 {t14, t15}, %8;\n\t"
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {t0,t1,t2,t3,t4,t5,t6,t7,t8,t9,t10,t11,t12,t13,t14,t15};\n\t}"
        ::"r"(a), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7])
        : "memory");
}

// whole CTA: warp 0 allocates all 512 columns (one CTA per SM: the 255-register threads fill the register file)
__device__ __forceinline__ uint32_t tm_alloc_cta(uint32_t* slot) {
    if ((threadIdx.x >> 5) == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return *slot;
}
__device__ __forceinline__ void tm_free_cta(uint32_t addr) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if ((threadIdx.x >> 5) == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(addr) : "memory");
}

__device__ constexpr uint32_t tm_blk_col(int bi, int bj) { return 18u * (uint32_t)(bj * (bj + 1) / 2 + bi); }  // bi <= bj

// Same interface as CovX (fbus_math.cuh).  Every call must be made by all 32 lanes of the warp (tcgen05 .sync.aligned);
// lanes that have nothing to do run the same code with neutral operands.
template <bool TLR = false>
struct CovTM {
    static constexpr bool kTLR = TLR;
    static constexpr bool kBlocked = true;  // block access is the cheap unit
    uint32_t base;         // TMEM address: lane partition of this warp, column 0 of the covariance
    double* TL = nullptr;  // TLR: the top-left 9x9 lives in the caller's registers
    __device__ __forceinline__ void fence_st() const { tm_wait_st(); }
    __device__ __forceinline__ double ld(int i, int j) const {
        if (TLR && i < 9 && j < 9) return TL[tlidx(i, j)];
        int bi = i / 3, bj = j / 3, r = i % 3, c = j % 3;
        if (bi > bj) { const int tb = bi; bi = bj; bj = tb; const int tr = r; r = c; c = tr; }
        double v;
        tm_ld1(base + tm_blk_col(bi, bj) + 2u * (uint32_t)(r * 3 + c), v);
        tm_wait_ld();
        return v;
    }
    __device__ __forceinline__ void st(int i, int j, double v) const {
        if (TLR && i < 9 && j < 9) { TL[tlidx(i, j)] = v; return; }
        int bi = i / 3, bj = j / 3, r = i % 3, c = j % 3;
        if (bi > bj) { const int tb = bi; bi = bj; bj = tb; const int tr = r; r = c; c = tr; }
        tm_st1(base + tm_blk_col(bi, bj) + 2u * (uint32_t)(r * 3 + c), v);
        if (bi == bj && r != c) tm_st1(base + tm_blk_col(bi, bj) + 2u * (uint32_t)(c * 3 + r), v);  // both triangles
    }
    // raw block as stored (bi <= bj), no wait
    __device__ __forceinline__ void ldraw(int bi, int bj, double* X) const {
        const uint32_t a = base + tm_blk_col(bi, bj);
        tm_ld8(a, X);
        tm_ld1(a + 16u, X[8]);
    }
    __device__ __forceinline__ void straw(int bi, int bj, const double* X) const {
        const uint32_t a = base + tm_blk_col(bi, bj);
#if FBUS_TMEM_WIDE_ST
        tm_st8(a, X);  // one 16-column store: needs the eight values in consecutive registers (ptxas adds moves)
        tm_st1(a + 16u, X[8]);
#else
        FBUS_UNROLL
        for (int e = 0; e < 9; ++e) tm_st1(a + 2u * (uint32_t)e, X[e]);  // nine 2-column stores straight from where the values are
#endif
    }
    // X[r*3+c] = P[3bi+r][3bj+c], any order of bi, bj
    __device__ __forceinline__ void ldblk(int bi, int bj, double* X) const {
        if (TLR && bi < 3 && bj < 3) {
            FBUS_UNROLL
            for (int r = 0; r < 3; ++r)
                FBUS_UNROLL
                for (int c = 0; c < 3; ++c) X[r * 3 + c] = TL[tlidx(3 * bi + r, 3 * bj + c)];
            return;
        }
        if (bi <= bj) {
            ldraw(bi, bj, X);
            tm_wait_ld();
        } else {
            double T[9];
            ldraw(bj, bi, T);
            tm_wait_ld();
            FBUS_UNROLL
            for (int r = 0; r < 3; ++r)
                FBUS_UNROLL
                for (int c = 0; c < 3; ++c) X[r * 3 + c] = T[c * 3 + r];
        }
    }
    // asynchronous form: issue the load of the stored block (min, max), wait_ld() once for a group, then fix() transposes
    // in registers when the caller asked for the block below the diagonal
    __device__ __forceinline__ void ldblk_nw(int bi, int bj, double* X) const {
        if (TLR && bi < 3 && bj < 3) { ldblk(bi, bj, X); return; }
        if (bi <= bj) ldraw(bi, bj, X);
        else ldraw(bj, bi, X);
    }
    __device__ __forceinline__ void wait_ld() const { tm_wait_ld(); }
    // cross blocks of rows 1, 2 (see CovX): one copy here
    __device__ __forceinline__ void ldtr_nw(int bi, int k, double* X, bool) const { ldblk_nw(bi, k, X); }
    __device__ __forceinline__ void sttr(int bi, int k, const double* X) const { stblk(bi, k, X); }
    __device__ __forceinline__ void fix(int bi, int bj, double* X) const {
        if (TLR && bi < 3 && bj < 3) return;
        if (bi > bj) {
            double t;
            t = X[1]; X[1] = X[3]; X[3] = t;
            t = X[2]; X[2] = X[6]; X[6] = t;
            t = X[5]; X[5] = X[7]; X[7] = t;
        }
    }
    __device__ __forceinline__ void lddiag(int b, double* X) const { ldblk(b, b, X); }
    __device__ __forceinline__ void ldany(int bi, int bj, double* X) const { ldblk(bi, bj, X); }
    __device__ __forceinline__ void stblk(int bi, int bj, const double* X) const {  // bi < bj
        if (TLR && bj < 3) {
            FBUS_UNROLL
            for (int r = 0; r < 3; ++r)
                FBUS_UNROLL
                for (int c = 0; c < 3; ++c) TL[tlidx(3 * bi + r, 3 * bj + c)] = X[r * 3 + c];
            return;
        }
        straw(bi, bj, X);
    }
    __device__ __forceinline__ void stdiag(int b, const double* X) const {  // upper 6 of X
        if (TLR && b < 3) {
            FBUS_UNROLL
            for (int r = 0; r < 3; ++r)
                FBUS_UNROLL
                for (int c = r; c < 3; ++c) TL[tlidx(3 * b + r, 3 * b + c)] = X[r * 3 + c];
            return;
        }
        const double F[9] = {X[0], X[1], X[2], X[1], X[4], X[5], X[2], X[5], X[8]};
        straw(b, b, F);
    }
};

}  // namespace fbus
