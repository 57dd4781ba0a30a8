// sfk_tables.cpp -- O(rings + spokes + window) constant tables built once per
// context on the host: ring normals, ring rotation matrices, LoS unit vectors,
// Savitzky-Golay taps, mean matter density.  These are the "negligible" rows
// a2-a4 of the hot-path table; all per-particle, per-LoS and per-halo numeric
// work runs in the CUDA kernels.  Compiled with -ffp-contract=off so float32 /
// float64 expressions round operation by operation like Go on amd64.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>

#include "sfk_internal.h"
#include "sfk_sgmath.cuh"

namespace sfk {

// Go evaluates untyped constant expressions exactly before rounding.
static const double kTwoPi = 6.28318530717958647692528676655900576839433879875021;
static const double kHalfPi = 1.57079632679489661923132169163975144209858469968755;
static const double k8PiG = 1.67700477795333747720267596902695825200521524500611e-9;

namespace {

// math/rand/xorshift.go:13-33 (xorshift128, 32-bit seed).
class Xorshift {
public:
    explicit Xorshift(uint64_t seed) : w_(static_cast<uint32_t>(seed)) {}
    double next() {
        double res;
        do {
            uint32_t t = x_ ^ (x_ << 11);
            x_ = y_; y_ = z_; z_ = w_;
            w_ = w_ ^ (w_ >> 19) ^ (t ^ (t >> 8));
            res = static_cast<double>(0xFFFFFFFFu - w_) / static_cast<double>(0xFFFFFFFFu);
        } while (res == 1.0);
        return res;
    }
    // math/rand/rand.go:91-96
    double uniform(double lo, double hi) { return (next() * (hi - lo)) + lo; }
private:
    uint32_t x_ = 123456789u, y_ = 362436069u, z_ = 521288629u, w_;
};

using Mat3 = std::array<float, 9>;

// math/mat/mat32.go:53-74: accumulate into a zeroed output, float32 throughout.
Mat3 mul(const Mat3 &a, const Mat3 &b) {
    Mat3 out{};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                float prod = a[3 * i + k] * b[3 * k + j];
                out[3 * i + j] = out[3 * i + j] + prod;
            }
    return out;
}

// los/geom/rotation.go:76-81
void spherical_angles(const float v[3], float &phi, float &theta) {
    phi = static_cast<float>(std::atan2(static_cast<double>(v[1]), static_cast<double>(v[0])));
    float n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    double r = std::sqrt(static_cast<double>(n2));
    theta = static_cast<float>(std::acos(static_cast<double>(v[2]) / r));
}

// los/geom/rotation.go:26-38: z-x-z Euler matrix, entries rounded to float32.
Mat3 euler(float phi, float theta, float psi) {
    auto cs = [](float a, float &c, float &s) {
        c = static_cast<float>(std::cos(static_cast<double>(a)));
        s = static_cast<float>(std::sin(static_cast<double>(a)));
    };
    float c1, s1, c2, s2, c3, s3;
    cs(phi, c1, s1); cs(theta, c2, s2); cs(psi, c3, s3);
    Mat3 r1{c1, -s1, 0, s1, c1, 0, 0, 0, 1};
    Mat3 r2{1, 0, 0, 0, c2, -s2, 0, s2, c2};
    Mat3 r3{c3, -s3, 0, s3, c3, 0, 0, 0, 1};
    return mul(r3, mul(r2, r1));
}

// los/geom/rotation.go:49-57: M * from = to.
Mat3 euler_between(const float from[3], const float to[3]) {
    float p1, t1, p2, t2;
    spherical_angles(from, p1, t1);
    spherical_angles(to, p2, t2);
    float half = static_cast<float>(kHalfPi);
    return euler(half - p1, t1 - t2, p2 - half);
}

}  // namespace

void build_tables(const sfk_params &p, Tables &t) {
    const int rings = static_cast<int>(p.rings), n = static_cast<int>(p.spokes);
    t.norms.assign(3 * rings, 0.f);
    // normVecs, cmd/shell.go:566-592 (every halo re-seeds: same normals)
    if (rings == 3) {
        const float axes[9] = {0, 0, 1, 0, 1, 0, 1, 0, 0};
        std::memcpy(t.norms.data(), axes, sizeof axes);
    } else {
        Xorshift gen(p.seed);
        for (int i = 0; i < rings; i++) {
            for (;;) {
                double x = gen.uniform(-1, +1), y = gen.uniform(-1, +1), z = gen.uniform(-1, +1);
                double r = std::sqrt(x * x + y * y + z * z);
                if (r < 1) {
                    t.norms[3 * i] = static_cast<float>(x / r);
                    t.norms[3 * i + 1] = static_cast<float>(y / r);
                    t.norms[3 * i + 2] = static_cast<float>(z / r);
                    break;
                }
            }
        }
    }
    // Halo.Init, los/halo.go:65-100
    const float zAxis[3] = {0, 0, 1};
    t.rots.resize(9 * rings);
    t.irots.resize(9 * rings);
    for (int i = 0; i < rings; i++) {
        Mat3 rot = euler_between(&t.norms[3 * i], zAxis);
        Mat3 irot = euler_between(zAxis, &t.norms[3 * i]);
        std::memcpy(&t.rots[9 * i], rot.data(), sizeof(float) * 9);
        std::memcpy(&t.irots[9 * i], irot.data(), sizeof(float) * 9);
    }
    t.ringCos.resize(n); t.ringSin.resize(n); t.ringPhi.resize(n);
    for (int i = 0; i < n; i++) {
        t.ringPhi[i] = static_cast<double>(i) / static_cast<double>(n) * kTwoPi;
        t.ringSin[i] = std::sin(t.ringPhi[i]);
        t.ringCos[i] = std::cos(t.ringPhi[i]);
    }
    t.dPhi = 1 / static_cast<double>(n) * kTwoPi;
    const int bins = static_cast<int>(p.radial_bins);
    t.dr0 = std::log(p.r_max_mult / p.r_min_mult) / static_cast<double>(bins);
    t.binInvG.resize(bins + 1);
    for (int k = 0; k <= bins; k++) t.binInvG[k] = 1.0 / std::exp(static_cast<double>(k) * t.dr0);
}

// Go's math.Pow: repeated squaring on frexp mantissas (rounds more than once,
// unlike libm's correctly rounded pow).
double go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (std::isnan(x) || std::isnan(y)) return NAN;
    if (x == 0 || std::isinf(x) || std::isinf(y)) return std::pow(x, y);
    if (y == 0.5) return std::sqrt(x);
    if (y == -0.5) return 1 / std::sqrt(x);
    double yi, yf = std::modf(std::fabs(y), &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) return std::pow(x, y);
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = std::exp(yf * std::log(x));
    }
    int xe;
    double x1 = std::frexp(x, &xe);
    for (int64_t i = static_cast<int64_t>(yi); i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    return std::ldexp(a1, ae);
}

// cosmo/param.go:12-41, cosmo/const.go:3-8
double rho_average(double H0, double omegaM, double omegaL, double z) {
    const double mpc = 3.08560e+22, msun = 1.98900e+30;
    double h0mks = (H0 * 1000) / mpc;
    double h100 = H0 / 100;
    double h0mksh = h0mks / h100;
    double hub = std::sqrt(omegaM * go_pow(1.0 + 0.0, 3.0) + omegaL) * h0mksh;
    double rhoCritMks = 3.0 * hub * hub / k8PiG;
    double rhoCrit = rhoCritMks * go_pow(mpc, 3) / msun;
    return rhoCrit * omegaM * go_pow(1 + z, 3.0);
}

namespace {

// math/mat/mat.go:247-348: Crout LU with implicit row scaling, then the
// reference's forward/back substitution (pivot applied as ys[pivot[i]] = b[i]).
bool lu_solve(std::vector<double> &m, int n, std::vector<double> &b) {
    std::vector<double> vv(n);
    std::vector<int> pivot(n);
    for (int i = 0; i < n; i++) {
        double big = 0.0;
        for (int j = 0; j < n; j++) big = std::fabs(m[i * n + j]) > big ? std::fabs(m[i * n + j]) : big;
        if (big == 0) return false;
        vv[i] = 1 / big;
    }
    int imax = 0;
    for (int k = 0; k < n; k++) {
        double big = 0.0;
        for (int i = k; i < n; i++) {
            double tmp = vv[i] * std::fabs(m[i * n + k]);
            if (tmp > big) { big = tmp; imax = i; }
        }
        if (k != imax) {
            for (int j = 0; j < n; j++) std::swap(m[imax * n + j], m[k * n + j]);
            vv[imax] = vv[k];
        }
        pivot[k] = imax;
        if (m[k * n + k] == 0) m[k * n + k] = 1e-20;
        for (int i = k + 1; i < n; i++) {
            m[i * n + k] /= m[k * n + k];
            double f = m[i * n + k];
            for (int j = k + 1; j < n; j++) {
                double prod = f * m[k * n + j];
                m[i * n + j] -= prod;
            }
        }
    }
    std::vector<double> rhs(b), y(b);
    for (int i = 0; i < n; i++) y[pivot[i]] = rhs[i];
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int j = 0; j < i; j++) { double prod = m[i * n + j] * y[j]; sum += prod; }
        y[i] = y[i] - sum;
    }
    for (int i = n - 1; i >= 0; i--) {
        double sum = 0.0;
        for (int j = i + 1; j < n; j++) { double prod = m[i * n + j] * b[j]; sum += prod; }
        b[i] = (y[i] - sum) / m[i * n + i];
    }
    return true;
}

}  // namespace

// math/interpolate/filter.go:218-318
bool savgol_taps(int order, int width, int ld, double dx, std::vector<double> &taps) {
    if (width % 2 != 1 || width <= order || ld > order) return false;
    const int half = width / 2, w = order + 1;
    std::vector<double> a(w * w, 0.0), b(w, 0.0);
    for (int ipj = 0; ipj <= 2 * order; ipj++) {
        double sum = ipj == 0 ? 1.0 : 0.0;
        for (int k = 1; k <= half; k++) sum += go_pow(static_cast<double>(k), ipj);
        for (int k = 1; k <= half; k++) sum += go_pow(static_cast<double>(-k), ipj);
        int mm = std::min(ipj, 2 * order - ipj);
        for (int imj = -mm; imj <= mm; imj += 2) a[((ipj - imj) / 2) * w + (ipj + imj) / 2] = sum;
    }
    b[ld] = 1;
    if (!lu_solve(a, w, b)) return false;
    taps.assign(width, 0.0);
    for (int i = -half; i <= half; i++) {
        double sum = b[0], fac = 1.0;
        for (int mm = 1; mm <= order; mm++) {
            fac *= static_cast<double>(i);
            double prod = b[mm] * fac;
            sum += prod;
        }
        taps[i + half] = sum;
    }
    if (ld > 0) {
        int fact = 1;
        for (int i = 2; i <= ld; i++) fact *= i;
        double scale = static_cast<double>(fact) / go_pow(dx, static_cast<double>(ld));
        for (double &c : taps) c *= scale;
    }
    return true;
}

// Polynomial form of the Savitzky-Golay taps for the sliding-moment path (sfk_sgmath.cuh):
// least-squares fit of degree 4 in u_j = (j - center)/center through all taps (they are such
// polynomials up to rounding), normal equations solved in long double.
bool savgol_poly(const std::vector<double> &taps0, const std::vector<double> &taps1, SgPoly &out) {
    const int window = static_cast<int>(taps0.size());
    if (window < 7 || taps1.size() != taps0.size() || window % 2 != 1) return false;
    const int center = window / 2;
    long double A[5][5] = {}, b0[5] = {}, b1[5] = {};
    for (int j = 0; j < window; j++) {
        const long double u = static_cast<long double>(j - center) / center;
        long double pw[9];
        pw[0] = 1;
        for (int k = 1; k < 9; k++) pw[k] = pw[k - 1] * u;
        for (int k = 0; k < 5; k++) {
            for (int l = 0; l < 5; l++) A[k][l] += pw[k + l];
            b0[k] += pw[k] * taps0[j];
            b1[k] += pw[k] * taps1[j];
        }
    }
    for (int c = 0; c < 5; c++) {   // Gaussian elimination with partial pivoting, two right-hand sides
        int piv = c;
        for (int r = c + 1; r < 5; r++) if (fabsl(A[r][c]) > fabsl(A[piv][c])) piv = r;
        if (A[piv][c] == 0) return false;
        if (piv != c) {
            for (int l = 0; l < 5; l++) std::swap(A[c][l], A[piv][l]);
            std::swap(b0[c], b0[piv]); std::swap(b1[c], b1[piv]);
        }
        for (int r = c + 1; r < 5; r++) {
            const long double f = A[r][c] / A[c][c];
            for (int l = c; l < 5; l++) A[r][l] -= f * A[c][l];
            b0[r] -= f * b0[c]; b1[r] -= f * b1[c];
        }
    }
    for (int r = 4; r >= 0; r--) {
        long double s0 = b0[r], s1 = b1[r];
        for (int l = r + 1; l < 5; l++) { s0 -= A[r][l] * out.c0[l]; s1 -= A[r][l] * out.c1[l]; }
        out.c0[r] = static_cast<double>(s0 / A[r][r]);
        out.c1[r] = static_cast<double>(s1 / A[r][r]);
    }
    const long double d = -1.0L / center;
    static const int binom[5][5] = {{1, 0, 0, 0, 0}, {1, 1, 0, 0, 0}, {1, 2, 1, 0, 0}, {1, 3, 3, 1, 0}, {1, 4, 6, 4, 1}};
    for (int k = 0; k < 5; k++) {
        for (int l = 0; l < 5; l++) out.T[k][l] = l <= k ? static_cast<double>(binom[k][l] * powl(d, k - l)) : 0.0;
        out.lo[k] = static_cast<double>(powl(-1.0L + d, k));
    }
    return true;
}

// Table of fast_log (sfk_kernels.cu): for c_i = 1 + (i + 1/2)/128 the pair (1/c_i rounded to
// double, -ln of that rounded value), so that ln m = tab[2i+1] + log1p(m * tab[2i] - 1) is an
// identity and the rounding of 1/c_i costs nothing.
void build_log_table(std::vector<double> &tab) {
    tab.resize(256);
    for (int i = 0; i < 128; i++) {
        const long double c = 1.0L + (static_cast<long double>(i) + 0.5L) / 128.0L;
        const double invc = static_cast<double>(1.0L / c);
        tab[2 * i] = invc;
        tab[2 * i + 1] = static_cast<double>(-logl(static_cast<long double>(invc)));
    }
}

}  // namespace sfk
