/* orc_math.c -- oracle (TEST INFRASTRUCTURE): leaf math of the shell path.
 * Restates math/rand/xorshift.go, los/geom/rotation.go, math/mat/mat32.go,
 * math/mat/mat.go (LU), math/interpolate/{filter,spline}.go, cosmo/.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -ffast-math). */
#include "shellfish_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* Go folds untyped constant expressions exactly before rounding to float64,
 * so 2*math.Pi, math.Pi/2, 4*math.Pi/3 and 8.0*math.Pi*GMks are the correctly
 * rounded values of the exact products -- NOT what C gets from M_PI. */
#define GO_2PI 6.28318530717958647692528676655900576839433879875021
#define GO_PI_2 1.57079632679489661923132169163975144209858469968755
#define GO_8PI_G 1.67700477795333747720267596902695825200521524500611e-9

/* ------------------------------------------------------------------ rng */

typedef struct { uint32_t w, x, y, z; } xorshift_gen;

/* math/rand/xorshift.go:17-22 -- only the low 32 bits of the seed are used */
static void xs_init(xorshift_gen *g, uint64_t seed) {
    g->x = 123456789u;
    g->y = 362436069u;
    g->z = 521288629u;
    g->w = (uint32_t)seed;
}

/* math/rand/xorshift.go:24-33 */
static double xs_next(xorshift_gen *g) {
    for (;;) {
        uint32_t t = g->x ^ (g->x << 11);
        g->x = g->y; g->y = g->z; g->z = g->w;
        g->w = g->w ^ (g->w >> 19) ^ (t ^ (t >> 8));
        double res = (double)(0xFFFFFFFFu - g->w) / (double)0xFFFFFFFFu;
        if (res != 1.0) return res;
    }
}

/* math/rand/rand.go:91-96 */
static double xs_uniform(xorshift_gen *g, double low, double high) {
    if (low == 0.0 && high == 1.0) return xs_next(g);
    return (xs_next(g) * (high - low)) + low;
}

/* cmd/shell.go:566-592 */
void orc_norm_vecs(uint64_t seed, int rings, float *out) {
    xorshift_gen g;
    xs_init(&g, seed);
    if (rings == 3) {
        static const float axes[9] = {0, 0, 1, 0, 1, 0, 1, 0, 0};
        memcpy(out, axes, sizeof axes);
        return;
    }
    for (int i = 0; i < rings; i++) {
        for (;;) {
            double x = xs_uniform(&g, -1, +1);
            double y = xs_uniform(&g, -1, +1);
            double z = xs_uniform(&g, -1, +1);
            double r = sqrt(x * x + y * y + z * z);
            if (r < 1) {
                out[3 * i + 0] = (float)(x / r);
                out[3 * i + 1] = (float)(y / r);
                out[3 * i + 2] = (float)(z / r);
                break;
            }
        }
    }
}

/* ------------------------------------------------------------ rotations */

/* los/geom/rotation.go:76-81 */
void orc_spherical_angles(float x, float y, float z, float *phi, float *theta) {
    *phi = (float)atan2((double)y, (double)x);
    float s = x * x + y * y + z * z; /* formed in float32 */
    double r = sqrt((double)s);
    *theta = (float)acos((double)z / r);
}

/* math/mat/mat32.go:53-74 for 3x3 */
static void mult33(const float *a, const float *b, float *out) {
    for (int i = 0; i < 9; i++) out[i] = 0;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) {
                float p = a[3 * i + k] * b[3 * k + j];
                out[3 * i + j] = out[3 * i + j] + p;
            }
}

/* los/geom/rotation.go:26-38 */
void orc_euler_matrix(float phi, float theta, float psi, float out[9]) {
    float c1 = (float)cos((double)phi), s1 = (float)sin((double)phi);
    float c2 = (float)cos((double)theta), s2 = (float)sin((double)theta);
    float c3 = (float)cos((double)psi), s3 = (float)sin((double)psi);
    float R1[9] = {c1, -s1, 0, s1, c1, 0, 0, 0, 1};
    float R2[9] = {1, 0, 0, 0, c2, -s2, 0, s2, c2};
    float R3[9] = {c3, -s3, 0, s3, c3, 0, 0, 0, 1};
    float m21[9];
    mult33(R2, R1, m21);
    mult33(R3, m21, out);
}

/* los/geom/rotation.go:49-57 */
void orc_euler_matrix_between(const float v1[3], const float v2[3], float out[9]) {
    float phi1, theta1, phi2, theta2;
    orc_spherical_angles(v1[0], v1[1], v1[2], &phi1, &theta1);
    orc_spherical_angles(v2[0], v2[1], v2[2], &phi2, &theta2);
    float pi2 = (float)GO_PI_2;
    float a = pi2 - phi1, b = theta1 - theta2, c = phi2 - pi2;
    orc_euler_matrix(a, b, c, out);
}

/* los/geom/rotation.go:60-65 */
void orc_rotate_vec(float v[3], const float m[9]) {
    float a0 = m[0] * v[0], a1 = m[1] * v[1], a2 = m[2] * v[2];
    float b0 = m[3] * v[0], b1 = m[4] * v[1], b2 = m[5] * v[2];
    float c0 = m[6] * v[0], c1 = m[7] * v[1], c2 = m[8] * v[2];
    float v0 = (a0 + a1) + a2;
    float v1 = (b0 + b1) + b2;
    float v2 = (c0 + c1) + c2;
    v[0] = v0; v[1] = v1; v[2] = v2;
}

/* ------------------------------------------------------------------ pow */

/* Go's math.Pow (src/math/pow.go, Go stdlib -- not in the reference tree):
 * integer part of y by repeated squaring on Frexp mantissas, which rounds
 * more than once and so differs from a correctly rounded libm pow() in the
 * last ulp for y >= 3.  Restated for the call sites on the shell path
 * (filter.go:279-283, penna.go:38-47, cosmo/param.go:13,33,40). */
double orc_go_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 || isinf(x) || isinf(y)) return pow(x, y); /* IEEE special cases */
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);
    double yi, yf = modf(fabs(y), &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9.223372036854775808e18) return pow(x, y);
    double a1 = 1.0;
    int ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf--; yi++; }
        a1 = exp(yf * log(x));
    }
    int xe;
    double x1 = frexp(x, &xe);
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) {
            ae += xe;
            break;
        }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    return ldexp(a1, ae);
}

/* ---------------------------------------------------------------- cosmo */

/* cosmo/param.go:12-41 with cosmo/const.go:3-8 */
double orc_rho_average(double H0, double omegaM, double omegaL, double z) {
    const double MpcMks = 3.08560e+22, MSunMks = 1.98900e+30;
    (void)z;
    /* RhoCritical(H0, omegaM, omegaL, 0) */
    double H0Mks = (H0 * 1000) / MpcMks;
    double H100 = H0 / 100;
    double H0MksH = H0Mks / H100;
    double frac = sqrt(omegaM * orc_go_pow(1.0 + 0.0, 3.0) + omegaL);
    double H = frac * H0MksH;
    double rhoCritMks = 3.0 * H * H / GO_8PI_G;
    double rhoCrit = rhoCritMks * orc_go_pow(MpcMks, 3) / MSunMks;
    return rhoCrit * omegaM * orc_go_pow(1 + z, 3.0);
}

/* ------------------------------------------------------------------- LU */

/* math/mat/mat.go:247-295 factorizeInPlace; pivot is what forwardSubst uses */
static int lu_factorize(double *m, int n, int *pivot) {
    double *vv = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
        double big = 0.0;
        for (int j = 0; j < n; j++) {
            double tmp = fabs(m[i * n + j]);
            if (tmp > big) big = tmp;
        }
        if (big == 0) { free(vv); return -1; }
        vv[i] = 1 / big;
    }
    int imax = 0;
    for (int k = 0; k < n; k++) {
        double big = 0.0;
        for (int i = k; i < n; i++) {
            double tmp = vv[i] * fabs(m[i * n + k]);
            if (tmp > big) { big = tmp; imax = i; }
        }
        if (k != imax) {
            for (int j = 0; j < n; j++) {
                double t = m[imax * n + j];
                m[imax * n + j] = m[k * n + j];
                m[k * n + j] = t;
            }
            vv[imax] = vv[k];
        }
        pivot[k] = imax;
        if (m[k * n + k] == 0) m[k * n + k] = 1e-20;
        for (int i = k + 1; i < n; i++) {
            m[i * n + k] /= m[k * n + k];
            double tmp = m[i * n + k];
            for (int j = k + 1; j < n; j++) {
                double p = tmp * m[k * n + j];
                m[i * n + j] -= p;
            }
        }
    }
    free(vv);
    return 0;
}

/* math/mat/mat.go:300-348 SolveVector / forwardSubst / backSubst */
static void lu_solve_vec(const double *lu, const int *pivot, int n,
                         const double *bsIn, double *xs) {
    double *bs = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(bs, bsIn, sizeof(double) * (size_t)n);
    double *ys = xs;
    for (int i = 0; i < n; i++) ys[pivot[i]] = bs[i];
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int j = 0; j < i; j++) {
            double p = lu[i * n + j] * ys[j];
            sum += p;
        }
        ys[i] = (ys[i] - sum);
    }
    for (int i = n - 1; i >= 0; i--) {
        double sum = 0.0;
        for (int j = i + 1; j < n; j++) {
            double p = lu[i * n + j] * xs[j];
            sum += p;
        }
        xs[i] = (ys[i] - sum) / lu[i * n + i];
    }
    free(bs);
}

int orc_lu_solve(double *a, int n, double *b) {
    int *pivot = (int *)malloc(sizeof(int) * (size_t)n);
    if (lu_factorize(a, n, pivot) != 0) { free(pivot); return -1; }
    lu_solve_vec(a, pivot, n, b, b);
    free(pivot);
    return 0;
}

/* math/mat/mat.go InvertAt/SolveMatrix: column-by-column solve of identity */
int orc_lu_invert(double *a, int n, double *out) {
    int *pivot = (int *)malloc(sizeof(int) * (size_t)n);
    if (lu_factorize(a, n, pivot) != 0) { free(pivot); return -1; }
    double *col = (double *)malloc(sizeof(double) * (size_t)n);
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) col[i] = (i == j) ? 1.0 : 0.0;
        lu_solve_vec(a, pivot, n, col, col);
        for (int i = 0; i < n; i++) out[i * n + j] = col[i];
    }
    free(col);
    free(pivot);
    return 0;
}

/* -------------------------------------------------------------- Sav-Gol */

/* math/interpolate/filter.go:265-310 savgol + :218-262 constructors */
int orc_savgol_kernel(int order, int width, int ld, double dx, double *taps) {
    if (width % 2 != 1 || width <= order || ld > order) return -1;
    int m = order, n = width / 2, w = m + 1;
    double *a = (double *)calloc((size_t)(w * w), sizeof(double));
    double *b = (double *)calloc((size_t)w, sizeof(double));
    for (int ipj = 0; ipj <= m * 2; ipj++) {
        double ipj64 = (double)ipj;
        double sum = 0.0;
        if (ipj == 0) sum = 1.0;
        for (int k = 1; k <= n; k++) sum += orc_go_pow((double)k, ipj64);
        for (int k = 1; k <= n; k++) sum += orc_go_pow((double)(-k), ipj64);
        int mm = 2 * m - ipj;
        if (mm > ipj) mm = ipj;
        for (int imj = -mm; imj <= mm; imj += 2) {
            int i = (ipj + imj) / 2, j = (ipj - imj) / 2;
            a[j * w + i] = sum;
        }
    }
    b[ld] = 1;
    if (orc_lu_solve(a, w, b) != 0) { free(a); free(b); return -1; }
    for (int i = -n; i <= n; i++) {
        double sum = b[0], fac = 1.0, i64 = (double)i;
        for (int mm = 1; mm < m + 1; mm++) {
            fac *= i64;
            double p = b[mm] * fac;
            sum += p;
        }
        taps[i + n] = sum;
    }
    if (ld > 0) {
        /* NewSavGolDerivKernel, filter.go:244-262: cs[i] *= ld! / dx^ld */
        int fact = 1;
        for (int i = 2; i <= ld; i++) fact *= i;
        double scale = (double)fact / orc_go_pow(dx, (double)ld);
        for (int i = 0; i < width; i++) taps[i] *= scale;
    }
    free(a);
    free(b);
    return 0;
}

/* math/interpolate/filter.go:104-161, BoundaryCondition = Extension.
 * The three loops of the reference collapse to one clamped index because
 * n >= width is guaranteed by analyze.Smooth (smooth.go:51-53). */
void orc_convolve_extension(const double *taps, int width, const double *xs,
                            int n, double *out) {
    int center = width / 2;
    for (int i = 0; i < n; i++) {
        double sum = 0.0;
        for (int j = 0; j < width; j++) {
            int idx = i + j - center;
            double x = idx < 0 ? xs[0] : (idx >= n ? xs[n - 1] : xs[idx]);
            double p = x * taps[j];
            sum += p;
        }
        out[i] = sum;
    }
}

/* --------------------------------------------------------------- spline */

struct orc_spline {
    int n, incr, bad;
    double dx;
    double *xs, *ys, *y2s;
    double *a, *b, *c, *d;
};

/* math/interpolate/spline.go:266-294 */
static int tri_diag_at(const double *as, const double *bs, const double *cs,
                       const double *rs, double *out, int n) {
    double *tmp = (double *)calloc((size_t)n, sizeof(double));
    double beta = bs[0];
    if (beta == 0) { free(tmp); return -1; }
    out[0] = rs[0] / beta;
    for (int i = 1; i < n; i++) {
        tmp[i] = cs[i - 1] / beta;
        beta = bs[i] - as[i] * tmp[i];
        if (beta == 0) { free(tmp); return -1; }
        out[i] = (rs[i] - as[i] * out[i - 1]) / beta;
    }
    for (int i = n - 2; i >= 0; i--) out[i] -= tmp[i + 1] * out[i + 1];
    free(tmp);
    return 0;
}

/* math/interpolate/spline.go:28-77,203-237 */
orc_spline *orc_spline_new(const double *xs, const double *ys, int n) {
    if (n <= 1) return NULL;
    orc_spline *sp = (orc_spline *)calloc(1, sizeof *sp);
    sp->n = n;
    sp->xs = (double *)malloc(sizeof(double) * (size_t)n);
    sp->ys = (double *)malloc(sizeof(double) * (size_t)n);
    sp->y2s = (double *)calloc((size_t)n, sizeof(double));
    sp->a = (double *)calloc((size_t)n, sizeof(double));
    sp->b = (double *)calloc((size_t)n, sizeof(double));
    sp->c = (double *)calloc((size_t)n, sizeof(double));
    sp->d = (double *)calloc((size_t)n, sizeof(double));
    memcpy(sp->xs, xs, sizeof(double) * (size_t)n);
    memcpy(sp->ys, ys, sizeof(double) * (size_t)n);
    if (xs[0] < xs[1]) {
        sp->incr = 1;
        for (int i = 0; i < n - 1; i++) if (xs[i + 1] < xs[i]) sp->bad = 1;
    } else {
        sp->incr = 0;
        for (int i = 0; i < n - 1; i++) if (xs[i + 1] > xs[i]) sp->bad = 1;
    }
    sp->dx = (xs[n - 1] - xs[0]) / (double)(n - 1);
    /* calcY2s */
    if (n > 2) {
        int k = n - 2;
        double *as = (double *)malloc(sizeof(double) * (size_t)k * 4);
        double *bs = as + k, *cs = bs + k, *rs = cs + k;
        for (int i = 0; i < k; i++) {
            int j = i + 1;
            as[i] = (xs[j] - xs[j - 1]) / 6;
            bs[i] = (xs[j + 1] - xs[j - 1]) / 3;
            cs[i] = (xs[j + 1] - xs[j]) / 6;
            rs[i] = ((ys[j + 1] - ys[j]) / (xs[j + 1] - xs[j])) -
                    ((ys[j] - ys[j - 1]) / (xs[j] - xs[j - 1]));
        }
        if (tri_diag_at(as, bs, cs, rs, sp->y2s + 1, k) != 0) sp->bad = 1;
        free(as);
    }
    /* calcCoeffs */
    for (int i = 0; i < n - 1; i++) {
        double dx = xs[i + 1] - xs[i];
        sp->a[i] = (-sp->y2s[i] / 6 + sp->y2s[i + 1] / 6) / dx;
        sp->b[i] = sp->y2s[i] / 2;
        sp->c[i] = (ys[i + 1] - ys[i]) / dx +
                   dx * (-sp->y2s[i] / 3 - sp->y2s[i + 1] / 6);
        sp->d[i] = ys[i];
    }
    return sp;
}

void orc_spline_free(orc_spline *sp) {
    if (!sp) return;
    free(sp->xs); free(sp->ys); free(sp->y2s);
    free(sp->a); free(sp->b); free(sp->c); free(sp->d);
    free(sp);
}

/* math/interpolate/spline.go:176-200 */
static int spline_bsearch(const orc_spline *sp, double x) {
    int n = sp->n;
    int guess = (int)((x - sp->xs[0]) / sp->dx);
    if (guess >= 0 && guess < n - 1 &&
        ((sp->xs[guess] <= x) == sp->incr) &&
        ((sp->xs[guess + 1] >= x) == sp->incr))
        return guess;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) / 2;
        if (sp->incr == (x >= sp->xs[mid])) lo = mid; else hi = mid;
    }
    if (lo == n - 1) return -1;
    return lo;
}

/* math/interpolate/spline.go:80-97 */
int orc_spline_eval(const orc_spline *sp, double x, double *y) {
    if (!sp || sp->bad) return -1;
    int n = sp->n;
    if (((x <= sp->xs[0]) == sp->incr) || ((x >= sp->xs[n - 1]) == sp->incr)) {
        if (x == sp->xs[0]) { *y = sp->ys[0]; return 0; }
        if (x == sp->xs[n - 1]) { *y = sp->ys[n - 1]; return 0; }
        return -1;
    }
    int i = spline_bsearch(sp, x);
    if (i < 0) return -1;
    double dx = x - sp->xs[i];
    double a = sp->a[i], b = sp->b[i], c = sp->c[i], d = sp->d[i];
    *y = a * dx * dx * dx + b * dx * dx + c * dx + d;
    return 0;
}
