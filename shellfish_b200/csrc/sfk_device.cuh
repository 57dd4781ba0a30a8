// sfk_device.cuh -- device helpers shared by the kernel files.
#pragma once

#include <cuda_runtime.h>

namespace sfk {

// sqrt for x >= 0: MUFU.RSQ64H seed (~2^-22), one Newton step on 1/sqrt
// (~2^-44), one residual correction of the root (<= 1 ulp), branch free.
// Callers test the sign first.
__device__ __forceinline__ double sqrt_pos(double x) {
    x += 1e-300;   // 0 -> 1e-150 (still far below every rMin); exact no-op for x > 1e-284
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double hx = 0.5 * x;
    const double e = fma(-hx * y, y, 0.5);
    y = fma(y, e, y);
    const double s = x * y;
    const double r = fma(-s, s, x);
    return fma(r, 0.5 * y, s);
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
