/* orc_shellprops.c -- oracle (TEST INFRASTRUCTURE): the Monte-Carlo shell
 * property integrals of `shellfish stats`, los/analyze/shell.go:21-32 (randomAngle,
 * cartesian), :58-68 (Volume), :93-111 (trisort), :113-199 (Axes), :201-226
 * (cosNorm), :229-237 (SurfaceArea), :268-283 (RadialRange) and BiCubic
 * (math/interpolate/cubic_interpolators.go:22-103), restated literally.
 *
 * PARITY UNPINNED: the reference draws its samples from Go's process-global
 * math/rand source (stdlib, not under /root/reference, never seeded by the
 * reference) and diagonalises the inertia tensor with gonum's mat64.Eigen
 * (un-vendored, gonum/matrix@c518dec07be9).  The sample stream is therefore an
 * input here (a callback returning rand.Float64()-like uniforms; the repo's
 * xorshift restatement serves as a stand-in), the eigen-decomposition is a
 * cyclic Jacobi iteration, and results can only be compared statistically.  ONE
 * reference-held row exists for that comparison: doc/quickstart.md:99-106,150-155
 * prints the coefficients of a halo and the `stats` row made of them; this
 * restatement reproduces its R_sp, Volume, Surface Area and axes within the scatter
 * of the 50 000-sample estimate (tests/test_shellprops.py).  The correction tables
 * are inputs too (the product interpolates the reference's own ellipse_grid values,
 * scripts/import_axis_tables.py; scripts/make_axis_tables.py recomputes them). */
#define _GNU_SOURCE
#include "shellfish_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* ---- a uniform stream: math/rand/xorshift.go:13-33 of the reference's own package */
void orc_stream_init(orc_stream *s, uint64_t seed) {
    s->x = 123456789u; s->y = 362436069u; s->z = 521288629u; s->w = (uint32_t)seed;
}

double orc_stream_next(void *state) {
    orc_stream *g = (orc_stream *)state;
    double res;
    do {
        uint32_t t = g->x ^ (g->x << 11);
        g->x = g->y; g->y = g->z; g->z = g->w;
        g->w = g->w ^ (g->w >> 19) ^ (t ^ (t >> 8));
        res = (double)(0xFFFFFFFFu - g->w) / (double)0xFFFFFFFFu;
    } while (res == 1.0);
    return res;
}

/* shell.go:21-24 */
static void random_angle(orc_uniform_fn fn, void *st, double *phi, double *theta) {
    double u = fn(st), v = fn(st);
    *phi = 2 * M_PI * u;
    *theta = acos(2 * v - 1);
}

/* shell.go:27-32 */
static void cartesian(double phi, double theta, double r, double *x, double *y, double *z) {
    double sinP = sin(phi), cosP = cos(phi);
    double sinT = sin(theta), cosT = cos(theta);
    *x = r * sinT * cosP; *y = r * sinT * sinP; *z = r * cosT;
}

static double shell(const double *cs, int order, double phi, double theta) {
    return orc_penna_func(cs, order, order, 2, phi, theta);
}

/* shell.go:59-68 */
double orc_shell_volume(const double *cs, int order, int samples, orc_uniform_fn fn, void *st) {
    double sum = 0.0;
    for (int i = 0; i < samples; i++) {
        double phi, theta;
        random_angle(fn, st, &phi, &theta);
        double r = shell(cs, order, phi, theta);
        sum += r * r * r;
    }
    double r3 = sum / (double)samples;
    return r3 * 4 * 1.04719755119659774615421446109316762806572313312504; /* math.Pi / 3, folded */
}

/* shell.go:201-226 */
double orc_cos_norm(const double *cs, int order, double phi, double theta) {
    double dp = 1e-3, dt = 1e-3;
    double x00, y00, z00, x01, y01, z01, x10, y10, z10, x11, y11, z11;
    double r00 = shell(cs, order, phi - dp, theta - dt);
    cartesian(phi - dp, theta - dt, r00, &x00, &y00, &z00);
    double r01 = shell(cs, order, phi - dp, theta + dt);
    cartesian(phi - dp, theta + dt, r01, &x01, &y01, &z01);
    double r10 = shell(cs, order, phi + dp, theta - dt);
    cartesian(phi + dp, theta - dt, r10, &x10, &y10, &z10);
    double r11 = shell(cs, order, phi + dp, theta + dt);
    cartesian(phi + dp, theta + dt, r11, &x11, &y11, &z11);

    double dxa = x00 - x11, dya = y00 - y11, dza = z00 - z11;
    double dxb = x01 - x10, dyb = y01 - y10, dzb = z01 - z10;

    double xn = dya * dzb - dza * dyb;
    double yn = dza * dxb - dxa * dzb;
    double zn = dxa * dyb - dya * dxb;
    double norm = sqrt(xn * xn + yn * yn + zn * zn);

    xn = xn / norm; yn = yn / norm; zn = zn / norm;
    double xl, yl, zl;
    cartesian(phi, theta, 1, &xl, &yl, &zl);
    return xl * xn + yl * yn + zl * zn;
}

/* shell.go:229-237 */
double orc_shell_surface_area(const double *cs, int order, int samples, orc_uniform_fn fn, void *st) {
    double sum = 0.0;
    for (int i = 0; i < samples; i++) {
        double phi, theta;
        random_angle(fn, st, &phi, &theta);
        double r = shell(cs, order, phi, theta);
        sum += r * r / orc_cos_norm(cs, order, phi, theta);
    }
    return sum / (double)samples * 4 * M_PI;
}

/* shell.go:268-283 */
void orc_shell_radial_range(const double *cs, int order, int samples, orc_uniform_fn fn, void *st,
                            double *low, double *high) {
    double phi, theta;
    random_angle(fn, st, &phi, &theta);
    *low = shell(cs, order, phi, theta);
    *high = *low;
    for (int i = 0; i < samples; i++) {
        random_angle(fn, st, &phi, &theta);
        double r = shell(cs, order, phi, theta);
        if (r > *high) *high = r;
        if (r < *low) *low = r;
    }
}

/* math/interpolate/cubic_interpolators.go:22-103: NewBiCubic + Eval.  Returns
 * -1 where Spline.Eval would panic (argument outside the knots). */
int orc_bicubic_eval(const double *xs, int nx, const double *ys, int ny, const double *vals,
                     double x, double y, double *out) {
    double *yVals = malloc(sizeof(double) * (size_t)ny);
    double *xSplineVals = malloc(sizeof(double) * (size_t)nx);
    int rc = 0;
    for (int xi = 0; xi < nx && rc == 0; xi++) {
        for (int yi = 0; yi < ny; yi++) yVals[yi] = vals[nx * yi + xi];
        orc_spline *sp = orc_spline_new(ys, yVals, ny);
        rc = orc_spline_eval(sp, y, &xSplineVals[xi]);
        orc_spline_free(sp);
    }
    if (rc == 0) {
        orc_spline *sp = orc_spline_new(xs, xSplineVals, nx);
        rc = orc_spline_eval(sp, x, out);
        orc_spline_free(sp);
    }
    free(yVals); free(xSplineVals);
    return rc;
}

/* shell.go:93-111 */
static void trisort(double x, double y, double z, double *a, double *b, double *c, int *aIdx) {
    double p, q;
    if (x > y && x > z) { *a = x; p = y; q = z; *aIdx = 0; }
    else if (y > x && y > z) { *a = y; p = z; q = x; *aIdx = 1; }
    else { *a = z; p = x; q = y; *aIdx = 2; }
    if (p > q) { *b = p; *c = q; } else { *b = q; *c = p; }
}

/* Eigenvalues (ascending) and eigenvectors (columns of v) of a symmetric 3x3
 * matrix by cyclic Jacobi -- stands in for mat64.Eigen, shell.go:150-157. */
static void sym_eigen3(double a[3][3], double w[3], double v[3][3]) {
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[i][j] = (i == j);
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
        if (off == 0.0) break;
        for (int p = 0; p < 2; p++) for (int q = p + 1; q < 3; q++) {
            if (a[p][q] == 0.0) continue;
            double phi = 0.5 * atan2(2 * a[p][q], a[q][q] - a[p][p]);
            double c = cos(phi), s = sin(phi);
            double r[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            r[p][p] = c; r[q][q] = c; r[p][q] = s; r[q][p] = -s;
            double t[3][3], u[3][3];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {   /* t = a r */
                t[i][j] = 0; for (int k = 0; k < 3; k++) t[i][j] += a[i][k] * r[k][j];
            }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {   /* a = r^T t */
                u[i][j] = 0; for (int k = 0; k < 3; k++) u[i][j] += r[k][i] * t[k][j];
            }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) a[i][j] = u[i][j];
            a[p][q] = a[q][p] = 0.5 * (a[p][q] + a[q][p]);
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {   /* v = v r */
                t[i][j] = 0; for (int k = 0; k < 3; k++) t[i][j] += v[i][k] * r[k][j];
            }
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[i][j] = t[i][j];
        }
    }
    int idx[3] = {0, 1, 2};
    for (int i = 0; i < 3; i++) for (int j = i + 1; j < 3; j++)
        if (a[idx[j]][idx[j]] < a[idx[i]][idx[i]]) { int t = idx[i]; idx[i] = idx[j]; idx[j] = t; }
    double vs[3][3];
    for (int k = 0; k < 3; k++) {
        w[k] = a[idx[k]][idx[k]];
        for (int i = 0; i < 3; i++) vs[i][k] = v[i][idx[k]];
    }
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) v[i][j] = vs[i][j];
}

/* shell.go:113-199.  tables may be NULL: the three ratios are then 1 (the
 * uncorrected axes make_grid.go tabulated).  Returns -1 where the reference
 * would panic (axis ratio outside the correction tables).  Eigenvalues are
 * taken in ascending order (gonum's order is not reproducible here); aVec is
 * the eigenvector of the smallest moment. */
int orc_shell_axes(const double *cs, int order, int samples, orc_uniform_fn fn, void *st,
                   const orc_axis_tables *tables, double *aOut, double *bOut, double *cOut,
                   double aVec[3]) {
    double nxx = 0.0, nyy = 0.0, nzz = 0.0;
    double nxy = 0.0, nyz = 0.0, nzx = 0.0;
    double nx = 0.0, ny = 0.0, nz = 0.0;
    double norm = 0.0;

    for (int i = 0; i < samples; i++) {
        double phi, theta, x, y, z;
        random_angle(fn, st, &phi, &theta);
        double r = shell(cs, order, phi, theta);
        double area = r * r / orc_cos_norm(cs, order, phi, theta);
        cartesian(phi, theta, r, &x, &y, &z);

        nxx += area * x * x;
        nyy += area * y * y;
        nzz += area * z * z;
        nxy += area * x * y;
        nyz += area * y * z;
        nzx += area * z * x;
        nx += area * x;
        ny += area * y;
        nz += area * z;

        norm += area;
    }

    nxx = nxx / norm; nyy = nyy / norm; nzz = nzz / norm;
    nxy = nxy / norm; nyz = nyz / norm; nzx = nzx / norm;
    nx = nx / norm; ny = ny / norm; nz = nz / norm;

    double mat[3][3] = {
        {nyy + nzz - ny * ny - nz * nz, -nxy + nx * ny, -nzx + nz * nx},
        {-nxy + nx * ny, nxx + nzz - nx * nx - nz * nz, -nyz + ny * nz},
        {-nzx + nz * nx, -nyz + ny * nz, nxx + nyy - nx * nx - ny * ny},
    };
    double vals[3], vecs[3][3];
    sym_eigen3(mat, vals, vecs);

    double Ix = vals[0], Iy = vals[1], Iz = vals[2];
    double ax2 = 3 * (Iy + Iz - Ix) / 2;
    double ay2 = 3 * Iy - ax2;
    double az2 = 3 * Iz - ax2;

    double a, b, c;
    int aIdx;
    trisort(sqrt(ax2), sqrt(ay2), sqrt(az2), &c, &b, &a, &aIdx);
    double ac = a / c, bc = b / c;

    double acRatio = 1, bcRatio = 1, cRatio = 1;
    if (tables) {
        if (orc_bicubic_eval(tables->acs, tables->n, tables->bcs, tables->n, tables->acGrid, ac, bc, &acRatio) ||
            orc_bicubic_eval(tables->acs, tables->n, tables->bcs, tables->n, tables->bcGrid, ac, bc, &bcRatio) ||
            orc_bicubic_eval(tables->acs, tables->n, tables->bcs, tables->n, tables->cGrid, ac, bc, &cRatio))
            return -1;
    }

    aVec[0] = vecs[0][0]; aVec[1] = vecs[1][0]; aVec[2] = vecs[2][0];
    norm = sqrt(aVec[0] * aVec[0] + aVec[1] * aVec[1] + aVec[2] * aVec[2]);
    aVec[0] = aVec[0] / norm; aVec[1] = aVec[1] / norm; aVec[2] = aVec[2] / norm;

    c = cRatio * c;
    *aOut = c; *bOut = bcRatio * bc * c; *cOut = acRatio * ac * c;
    return 0;
}

/* shell.go:295-335.  rs, fs: [bins].  Go's int() conversion of a float64 truncates
 * toward zero; NaN and out-of-range values become the most negative int64 on amd64. */
void orc_angular_fraction_profile(const double *cs, int order, int samples, int bins, double rMin, double rMax,
                                  orc_uniform_fn fn, void *st, double *rs, double *fs) {
    int64_t *ns = calloc((size_t)bins, sizeof(int64_t));
    double lrMin = log(rMin), lrMax = log(rMax);
    double dlr = (lrMax - lrMin) / (double)bins;

    for (int i = 0; i < bins; i++) rs[i] = exp(lrMin + ((double)i + 0.5) * dlr);

    for (int i = 0; i < samples; i++) {
        double phi, theta;
        random_angle(fn, st, &phi, &theta);
        double lr = log(shell(cs, order, phi, theta));
        double t = (lr - lrMin) / dlr;
        int64_t lri = (t != t || t >= 9.2e18 || t <= -9.2e18) ? INT64_MIN : (int64_t)t;
        if (lri < 0 || lri >= bins) continue;
        ns[lri]++;
    }

    /* reverse cumulative sum */
    for (int i = bins - 2; i >= 0; i--) ns[i] += ns[i + 1];

    for (int i = 0; i < bins; i++) fs[i] = (double)ns[i] / (double)ns[0];
    free(ns);
}
