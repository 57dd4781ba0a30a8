// sfk_sgmath.cuh -- Savitzky-Golay smoothing by sliding polynomial moments
// (default path of los_kernel_fast; SFK_FLAG_LOS_TAPLOOP selects the literal tap loop).
//
// The taps of an order-4 Savitzky-Golay kernel (math/interpolate/filter.go:218-318) are,
// exactly, polynomials of degree 4 in the tap offset: k[j] = sum_k c_k u_j^k with
// u_j = (j - center)/center.  Hence
//     (K * y)[i] = sum_j k[j] y[i + j - center] = sum_k c_k M_k(i),
//     M_k(i) = sum_j u_j^k y[i + j - center]                      (five moments),
// and the moments of output i + 1 follow from those of output i by a binomial shift
// plus the sample that leaves and the one that enters the window:
//     M_k(i+1) = sum_{l<=k} C(k,l) (-1/center)^(k-l) M_l(i) - (-1 - 1/center)^k y_out + y_in.
// One direct evaluation (window x 5 FMAs) followed by PER - 1 shifts (~25 FMAs each)
// replaces PER x window x 2 FMAs.  The shift is a unipotent recursion, so rounding
// errors grow polynomially with the number of steps; it is re-anchored every PER
// outputs (PER = 8: errors stay at the 1e-15 level, measured in tests/test_sgmoments.py).
//
// The same source is compiled by nvcc for the kernel and by g++ for tests/hostcheck.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define SFK_SG_HD __host__ __device__
#else
#define SFK_SG_HD
#endif
#if defined(__CUDA_ARCH__)
#define SFK_SG_UNROLL _Pragma("unroll")
#else
#define SFK_SG_UNROLL
#endif

namespace sfk {

struct SgPoly {
    double c0[5], c1[5];   // value / derivative taps as polynomials of u
    double T[5][5];        // binomial shift by one sample, T[k][l] = C(k,l) (-1/center)^(k-l), l <= k
    double lo[5];          // (-1 - 1/center)^k: weight of the sample leaving the window
};

// M += x * (1, u, u^2, u^3, u^4); up = {u, u^2, u^3, u^4} of this tap
SFK_SG_HD inline void sg_accumulate(double M[5], double x, const double *up) {
    M[0] += x;
    M[1] = fma(x, up[0], M[1]);
    M[2] = fma(x, up[1], M[2]);
    M[3] = fma(x, up[2], M[3]);
    M[4] = fma(x, up[3], M[4]);
}

// The same for the two taps at offsets +u and -u at once (yp, ym their samples): even moments
// take the sum, odd ones the difference -- 7 operations instead of 10 (los_kernel_fast's anchor pass);
// checked in tests/test_sgmoments.py.
SFK_SG_HD inline void sg_accumulate_pair(double M[5], double yp, double ym, const double *up) {
    const double s = yp + ym, d = yp - ym;
    M[0] += s;
    M[1] = fma(d, up[0], M[1]);
    M[2] = fma(s, up[1], M[2]);
    M[3] = fma(d, up[2], M[3]);
    M[4] = fma(s, up[3], M[4]);
}

SFK_SG_HD inline void sg_outputs(const double M[5], const SgPoly &p, double &s0, double &s1) {
    double a = 0.0, b = 0.0;
SFK_SG_UNROLL
    for (int k = 0; k < 5; k++) { a = fma(p.c0[k], M[k], a); b = fma(p.c1[k], M[k], b); }
    s0 = a; s1 = b;
}

// moments of output i -> output i + 1 (descending k: every row reads not yet updated lower moments)
SFK_SG_HD inline void sg_advance(double M[5], const SgPoly &p, double yOut, double yIn) {
SFK_SG_UNROLL
    for (int k = 4; k >= 0; k--) {
        double m = M[k];
SFK_SG_UNROLL
        for (int l = 0; l < 5; l++)
            if (l < k) m = fma(p.T[k][l], M[l], m);
        M[k] = fma(-p.lo[k], yOut, m) + yIn;
    }
}

}  // namespace sfk
