// sfk_device.cuh -- device helpers shared by the kernel files.
#pragma once

#include <cuda_runtime.h>

namespace sfk {

// sqrt for x >= 0: MUFU.RSQ64H seed (~2^-22), one Newton step on 1/sqrt
// (~2^-44), one residual correction of the root (<= 1 ulp), branch free.
// Callers test the sign first.
__device__ __forceinline__ double sqrt_pos(double x) {
    x += 0x1p-996;   // 0 -> 2^-498 (far below every rMin); exact no-op for x > 1e-284; a power of two is an immediate operand
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    const double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    const double s = x * y;
    const double r = fma(-s, s, x);
    return fma(r, 0.5 * y, s);
}

// The same root in one instruction less on the FP64 pipe: coupled (Goldschmidt) iteration on
// g ~ sqrt(x) and h ~ 1/(2 sqrt(x)), then the residual correction.  The seed reads only the high
// word (MUFU.RSQ64H), clamped to the smallest normal number so that x == 0 returns 0 instead of
// 0 * Inf; 2 x 10^7 random arguments: equal to sqrt_pos and to the correctly rounded root.
__device__ __forceinline__ double sqrt_pos0(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(__hiloint2double(max(__double2hiint(x), 0x00100000), 0)));
    double g = x * y, h = 0.5 * y;
    const double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    return fma(fma(-g, g, x), h, g);
}

// ---- TMA: 1-D bulk copies global -> shared, completion on an mbarrier -------------------
__device__ __forceinline__ uint32_t smem_addr_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_addr_u32(bar)), "r"(parity), "r"(0x989680u)   // suspend-time hint: sleep, do not spin
            : "memory");
    } while (!done);
}
// src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_addr_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr_u32(bar))
        : "memory");
}

// Go's math.Pow for small non-negative integer exponents (src/math/pow.go):
// repeated squaring on frexp mantissas.
__device__ inline double go_pow_int(double x, int yi) {
    if (yi == 0 || x == 1) return 1;
    if (yi == 1) return x;
    if (isnan(x)) return x;
    if (x == 0) return 0;   // y > 0
    if (isinf(x)) return pow(x, static_cast<double>(yi));
    double a1 = 1.0;
    int ae = 0, xe;
    double x1 = frexp(x, &xe);
    for (int i = yi; i != 0; i >>= 1) {
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    return ldexp(a1, ae);
}

}  // namespace sfk
