// sfk_shellmath.cuh -- per-node arithmetic of the shell property integrals of
// `shellfish stats` (los/analyze/shell.go:58-275): the Penna-Dines surface
// R(phi, theta), cosNorm, and the quantities Volume / SurfaceArea / Axes /
// RadialRange sum over solid-angle samples.
//
// The same source is compiled twice: by nvcc as device code for
// shell_props_kernel (sfk_shellprops.cu) and by g++ into tests/hostcheck (test
// infrastructure), so the CPU suite can check this arithmetic against the literal
// Monte-Carlo restatement without a GPU.  Nothing in libsfk.so calls the host instantiation.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define SFK_HD __host__ __device__
#else
#define SFK_HD
#endif

namespace sfk {

constexpr int kShellMaxOrder = 4;
// Per-halo sums of one pass over the quadrature nodes.
enum {
    SS_R3 = 0,   // mean r^3                         -> Volume, shell.go:59-68
    SS_AREA,     // mean r^2 / cosNorm               -> SurfaceArea, :229-237; Axes' norm, :133
    SS_XX, SS_YY, SS_ZZ, SS_XY, SS_YZ, SS_ZX,   // mean area * x x ...   Axes, :123-128
    SS_X, SS_Y, SS_Z,                           // mean area * x ...     Axes, :129-131
    SS_RMIN, SS_RMAX,                           // RadialRange, :268-283
    SS_COUNT
};

// Go's math.Pow for small non-negative integer exponents (src/math/pow.go):
// repeated squaring on frexp mantissas.
SFK_HD inline double go_pow_small(double x, int yi) {
    if (yi == 0 || x == 1) return 1;
    if (yi == 1) return x;
    if (x != x) return x;
    if (x == 0) return 0;   // y > 0
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

// PennaFunc(cs, order, order, 2)(phi, theta), los/analyze/penna.go:62-82.  The
// reference calls math.Pow four times per term; the powers are the same numbers
// whatever term asks for them, so each distinct one is formed once.  The product
// keeps the reference's association: ((((c * sinIJ) * cosK) * sinJ) * cosI).
SFK_HD inline double penna_radius(const double *cs, int order, double phi, double theta) {
    double sinPhi, cosPhi, sinTh, cosTh;
    sincos(phi, &sinPhi, &cosPhi);
    sincos(theta, &sinTh, &cosTh);
    double cosI[kShellMaxOrder], sinJ[kShellMaxOrder], sinIJ[2 * kShellMaxOrder - 1];
    for (int i = 0; i < order; i++) { cosI[i] = go_pow_small(cosPhi, i); sinJ[i] = go_pow_small(sinPhi, i); }
    for (int e = 0; e < 2 * order - 1; e++) sinIJ[e] = go_pow_small(sinTh, e);
    int idx = 0;
    double sum = 0.0;
    for (int k = 0; k < 2; k++) {
        const double cosK = go_pow_small(cosTh, k);
        for (int j = 0; j < order; j++)
            for (int i = 0; i < order; i++) {
                sum += cs[idx] * sinIJ[i + j] * cosK * sinJ[j] * cosI[i];
                idx++;
            }
    }
    return sum;
}

// cartesian, shell.go:27-32
SFK_HD inline void shell_cartesian(double phi, double theta, double r, double &x, double &y, double &z) {
    double sinP, cosP, sinT, cosT;
    sincos(phi, &sinP, &cosP);
    sincos(theta, &sinT, &cosT);
    x = r * sinT * cosP; y = r * sinT * sinP; z = r * cosT;
}

// cosNorm, shell.go:201-226: cosine between r-hat and the surface normal from
// the two diagonals of a +-1e-3 rad patch.
SFK_HD inline double shell_cos_norm(const double *cs, int order, double phi, double theta) {
    const double dp = 1e-3, dt = 1e-3;
    double x00, y00, z00, x01, y01, z01, x10, y10, z10, x11, y11, z11;
    shell_cartesian(phi - dp, theta - dt, penna_radius(cs, order, phi - dp, theta - dt), x00, y00, z00);
    shell_cartesian(phi - dp, theta + dt, penna_radius(cs, order, phi - dp, theta + dt), x01, y01, z01);
    shell_cartesian(phi + dp, theta - dt, penna_radius(cs, order, phi + dp, theta - dt), x10, y10, z10);
    shell_cartesian(phi + dp, theta + dt, penna_radius(cs, order, phi + dp, theta + dt), x11, y11, z11);
    const double dxa = x00 - x11, dya = y00 - y11, dza = z00 - z11;
    const double dxb = x01 - x10, dyb = y01 - y10, dzb = z01 - z10;
    double xn = dya * dzb - dza * dyb;
    double yn = dza * dxb - dxa * dzb;
    double zn = dxa * dyb - dya * dxb;
    const double norm = sqrt(xn * xn + yn * yn + zn * zn);
    xn = xn / norm; yn = yn / norm; zn = zn / norm;
    double xl, yl, zl;
    shell_cartesian(phi, theta, 1, xl, yl, zl);
    return xl * xn + yl * yn + zl * zn;
}

// One solid-angle sample (phi, theta) with weight w (the weights of a pass sum to
// one, so the accumulated values are the means the reference divides out).
SFK_HD inline void shell_node(const double *cs, int order, double phi, double theta, double w, double *acc) {
    const double r = penna_radius(cs, order, phi, theta);
    acc[SS_R3] += w * (r * r * r);
    const double area = r * r / shell_cos_norm(cs, order, phi, theta);
    double x, y, z;
    shell_cartesian(phi, theta, r, x, y, z);
    const double wa = w * area;
    acc[SS_AREA] += wa;
    acc[SS_XX] += wa * x * x; acc[SS_YY] += wa * y * y; acc[SS_ZZ] += wa * z * z;
    acc[SS_XY] += wa * x * y; acc[SS_YZ] += wa * y * z; acc[SS_ZX] += wa * z * x;
    acc[SS_X] += wa * x; acc[SS_Y] += wa * y; acc[SS_Z] += wa * z;
    if (r < acc[SS_RMIN]) acc[SS_RMIN] = r;
    if (r > acc[SS_RMAX]) acc[SS_RMAX] = r;
}

// Node q of the product rule: Gauss-Legendre in cos(theta) (nTheta nodes glX with
// weights glW summing to 2) times nPhi equally spaced azimuths.  randomAngle
// (shell.go:21-24) draws phi = 2 pi u, theta = acos(2 v - 1) with u, v uniform:
// the rule integrates the same measure.
SFK_HD inline void shell_rule_node(int q, int nPhi, const double *glX, const double *glW, double &phi,
                                   double &theta, double &w) {
    const int it = q / nPhi, ip = q - it * nPhi;
    phi = 6.283185307179586 * ((ip + 0.5) / nPhi);
    theta = acos(glX[it]);
    w = glW[it] * 0.5 / nPhi;
}

// Bin of one sample of Shell.AngularFractionProfile (shell.go:311-318): Go's int()
// truncates toward zero, so (lr - lrMin)/dlr in (-1, 0) lands in bin 0; NaN converts to
// the most negative int64 on amd64 and is skipped.  Returns -1 for "continue".
SFK_HD inline int shell_fraction_bin(double r, double lrMin, double dlr, int bins) {
    const double lr = log(r);
    const double t = (lr - lrMin) / dlr;
    if (!(t > -1.0)) return -1;               // negative index or NaN
    if (t >= static_cast<double>(bins)) return -1;
    return static_cast<int>(t);
}

// Direction q of the angular-fraction profile's sample: a histogram is not a smooth
// functional, so instead of the Gauss-Legendre rule it uses equal-area strata -- nTheta
// midpoints in cos(theta) x nPhi midpoints in phi, every direction with the same weight,
// so the histogram holds integer counts exactly like the reference's.
SFK_HD inline void shell_stratum(int q, int nTheta, int nPhi, double &phi, double &theta) {
    const int it = q / nPhi, ip = q - it * nPhi;
    phi = 6.283185307179586 * ((ip + 0.5) / nPhi);
    theta = acos(-1.0 + (it + 0.5) * (2.0 / nTheta));
}

}  // namespace sfk
