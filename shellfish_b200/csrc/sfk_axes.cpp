// sfk_axes.cpp -- O(1)-per-halo closing arithmetic of the shell property
// integrals of `shellfish stats`: turns the per-halo sums of shell_props_kernel
// into the catalog columns of cmd/stats.go:343-349 (R_sp, Volume, SurfaceArea,
// axes, major-axis direction, RMin, RMax).  Follows Shell.Volume / SurfaceArea /
// Axes (los/analyze/shell.go:58-199) from the point where their sampling loops
// end.  Also builds the Gauss-Legendre rule the kernel samples on.
#include <algorithm>
#include <cmath>
#include <vector>

#include "sfk_internal.h"
#include "sfk_shellmath.cuh"

namespace sfk {

namespace {

// The correction tables Shell.Axes interpolates are data of the reference
// (los/analyze/ellipse_grid/grid.go); the product uses those very numbers
// (scripts/import_axis_tables.py).  sfk_axis_tables.inc holds the same grids recomputed from
// first principles (scripts/make_axis_tables.py): a cross-check kept by tests/test_shellprops.py.
#include "sfk_axis_tables_ref.inc"   // kRefAxisN, kRefAxisRatios, kRefACGrid, kRefBCGrid, kRefCGrid

const double kPi = 3.14159265358979323846264338327950288419716939937510582097494459;
// Go folds constant expressions exactly before rounding once.
const double kFourPiOver3 = 4.18879020478639098461685784437267051226289253250014;   // math.Pi*4/3
const double kPiOver3 = 1.04719755119659774615421446109316762806572313312504;       // math.Pi/3

// Natural cubic spline through increasing knots, math/interpolate/spline.go:
// calcY2s :203-226 (second derivatives, zero at both ends, TriDiagAt :266-295),
// calcCoeffs :228-237, Eval :80-97.
struct NaturalSpline {
    std::vector<double> xs, ys, a, b, c, d;
    void init(const double *x, const double *y, int n) {
        xs.assign(x, x + n); ys.assign(y, y + n);
        std::vector<double> y2(n, 0.0);
        const int m = n - 2;
        if (m > 0) {
            std::vector<double> as(m), bs(m), cs(m), rs(m), tmp(m, 0.0), out(m);
            for (int i = 0; i < m; i++) {
                const int j = i + 1;
                as[i] = (xs[j] - xs[j - 1]) / 6;
                bs[i] = (xs[j + 1] - xs[j - 1]) / 3;
                cs[i] = (xs[j + 1] - xs[j]) / 6;
                rs[i] = ((ys[j + 1] - ys[j]) / (xs[j + 1] - xs[j])) - ((ys[j] - ys[j - 1]) / (xs[j] - xs[j - 1]));
            }
            double beta = bs[0];
            out[0] = rs[0] / beta;
            for (int i = 1; i < m; i++) {
                tmp[i] = cs[i - 1] / beta;
                beta = bs[i] - as[i] * tmp[i];
                out[i] = (rs[i] - as[i] * out[i - 1]) / beta;
            }
            for (int i = m - 2; i >= 0; i--) out[i] -= tmp[i + 1] * out[i + 1];
            for (int i = 0; i < m; i++) y2[i + 1] = out[i];
        }
        a.resize(n - 1); b.resize(n - 1); c.resize(n - 1); d.resize(n - 1);
        for (int i = 0; i < n - 1; i++) {
            const double dx = xs[i + 1] - xs[i];
            a[i] = (-y2[i] / 6 + y2[i + 1] / 6) / dx;
            b[i] = y2[i] / 2;
            c[i] = (ys[i + 1] - ys[i]) / dx + dx * (-y2[i] / 3 - y2[i + 1] / 6);
            d[i] = ys[i];
        }
    }
    // Spline.Eval panics outside the table; the argument is clamped here instead
    // (see axes_from_moments).
    double eval(double x) const {
        const int n = static_cast<int>(xs.size());
        if (x <= xs[0]) return ys[0];
        if (x >= xs[n - 1]) return ys[n - 1];
        int lo = 0, hi = n - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) / 2;
            if (x >= xs[mid]) lo = mid; else hi = mid;
        }
        const double dx = x - xs[lo];
        return a[lo] * dx * dx * dx + b[lo] * dx * dx + c[lo] * dx + d[lo];
    }
};

// BiCubic, math/interpolate/cubic_interpolators.go:22-103: one spline in y per x knot
// (vals[nx*yi + xi]) built once (initSplines :71-90), then per evaluation a spline in x through
// their values at y (Eval :92-103).
struct BiCubic {
    std::vector<NaturalSpline> ySplines;
    const double *knots;
    int n;
    BiCubic(const double *kn, const double *vals, int count) : ySplines(count), knots(kn), n(count) {
        std::vector<double> col(n);
        for (int xi = 0; xi < n; xi++) {
            for (int yi = 0; yi < n; yi++) col[yi] = vals[n * yi + xi];
            ySplines[xi].init(knots, col.data(), n);
        }
    }
    double eval(double x, double y) const {
        std::vector<double> atY(n);
        for (int xi = 0; xi < n; xi++) atY[xi] = ySplines[xi].eval(y);
        NaturalSpline sp;
        sp.init(knots, atY.data(), n);
        return sp.eval(x);
    }
};

// Cyclic Jacobi for a symmetric 3x3 matrix; eigenvectors in the columns of v.
void jacobi3(double a[3][3], double w[3], double v[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) v[i][j] = i == j;
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
        const double diag = std::fabs(a[0][0]) + std::fabs(a[1][1]) + std::fabs(a[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                if (a[p][q] == 0) continue;
                const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
                const double cs = 1 / std::sqrt(t * t + 1), sn = t * cs;
                for (int k = 0; k < 3; k++) {   // A <- A J
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = cs * akp - sn * akq; a[k][q] = sn * akp + cs * akq;
                }
                for (int k = 0; k < 3; k++) {   // A <- J^T A
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = cs * apk - sn * aqk; a[q][k] = sn * apk + cs * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = cs * vkp - sn * vkq; v[k][q] = sn * vkp + cs * vkq;
                }
            }
    }
    for (int i = 0; i < 3; i++) w[i] = a[i][i];
}

// trisort, shell.go:93-111
void trisort(double x, double y, double z, double &a, double &b, double &c, int &aIdx) {
    double p, q;
    if (x > y && x > z) { a = x; p = y; q = z; aIdx = 0; }
    else if (y > x && y > z) { a = y; p = z; q = x; aIdx = 1; }
    else { a = z; p = x; q = y; aIdx = 2; }
    if (p > q) { b = p; c = q; } else { b = q; c = p; }
}

}  // namespace

// Gauss-Legendre nodes (ascending) and weights on [-1, 1] by Newton iteration on
// P_n; the weights sum to 2.
void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w) {
    x.assign(n, 0.0); w.assign(n, 0.0);
    for (int i = 0; i < (n + 1) / 2; i++) {
        double z = std::cos(kPi * (i + 0.75) / (n + 0.5));
        double pp = 1;
        for (int it = 0; it < 100; it++) {
            double p1 = 1, p2 = 0;
            for (int j = 0; j < n; j++) {
                const double p3 = p2;
                p2 = p1;
                p1 = ((2.0 * j + 1.0) * z * p2 - j * p3) / (j + 1);
            }
            pp = n * (z * p1 - p2) / (z * z - 1);
            const double dz = p1 / pp;
            z -= dz;
            if (std::fabs(dz) < 1e-16) break;
        }
        x[i] = -z; x[n - 1 - i] = z;
        w[i] = w[n - 1 - i] = 2 / ((1 - z * z) * pp * pp);
    }
}

// Shell.Axes from shell.go:140 on.  m = area-weighted means {xx, yy, zz, xy, yz,
// zx, x, y, z} (already divided by norm, :140-142).  out = {major, intermediate,
// minor, Ax, Ay, Az} in the order of the catalog columns (cmd/stats.go:345).
//
// Deliberate differences, both where the reference's result is not a function of
// its input (PARITY UNPINNED):
//  * the reference takes the eigenvalues in whatever order gonum's general
//    (non-symmetric) mat64.Eigen returns them and pairs ay2 / az2 with the wrong
//    eigenvector column (:164-166,186), so its aVec is the major axis only when
//    the smallest moment comes first.  {a, b, c} do not depend on that order;
//    here aVec is always the eigenvector of the smallest moment, i.e. the major
//    axis the column header promises, with the sign fixed so that its largest
//    component is positive.
//  * axis ratios below the first knot of the correction tables (0.2069) make
//    Spline.Eval panic in the reference; they are clamped to the table edge.
void axes_from_moments(const double m[9], double out[6]) {
    const double nxx = m[0], nyy = m[1], nzz = m[2], nxy = m[3], nyz = m[4], nzx = m[5];
    const double nx = m[6], ny = m[7], nz = m[8];
    double mat[3][3] = {
        {nyy + nzz - ny * ny - nz * nz, -nxy + nx * ny, -nzx + nz * nx},
        {-nxy + nx * ny, nxx + nzz - nx * nx - nz * nz, -nyz + ny * nz},
        {-nzx + nz * nx, -nyz + ny * nz, nxx + nyy - nx * nx - ny * ny},
    };
    double vals[3], vecs[3][3];
    jacobi3(mat, vals, vecs);
    // ascending moments: the first belongs to the major axis
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int i, int j) { return vals[i] < vals[j]; });
    const double Ix = vals[ord[0]], Iy = vals[ord[1]], Iz = vals[ord[2]];
    const double ax2 = 3 * (Iy + Iz - Ix) / 2;
    const double ay2 = 3 * Iy - ax2;
    const double az2 = 3 * Iz - ax2;
    double c, b, a;
    int aIdx;
    trisort(std::sqrt(ax2), std::sqrt(ay2), std::sqrt(az2), c, b, a, aIdx);
    (void)aIdx;
    const double ac = a / c, bc = b / c;
    // axisInterpolators, shell.go:356-374: built once per process
    static const BiCubic acTable(kRefAxisRatios, kRefACGrid, kRefAxisN), bcTable(kRefAxisRatios, kRefBCGrid, kRefAxisN),
        cTable(kRefAxisRatios, kRefCGrid, kRefAxisN);
    const double acRatio = acTable.eval(ac, bc);
    const double bcRatio = bcTable.eval(ac, bc);
    const double cRatio = cTable.eval(ac, bc);
    double v[3] = {vecs[0][ord[0]], vecs[1][ord[0]], vecs[2][ord[0]]};
    const double norm = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    int big = 0;
    for (int i = 1; i < 3; i++) if (std::fabs(v[i]) > std::fabs(v[big])) big = i;
    const double sgn = v[big] < 0 ? -1.0 : 1.0;
    for (int i = 0; i < 3; i++) out[3 + i] = sgn * v[i] / norm;
    c = cRatio * c;
    out[0] = c; out[1] = bcRatio * bc * c; out[2] = acRatio * ac * c;
}

// props = {R_sp, Volume, SurfaceArea, a, b, c, Ax, Ay, Az, RMin, RMax}: the float
// columns of the stats catalog after M_sp (cmd/stats.go:343-349).
void shell_props_finalize(const double *sums, double *props) {
    const double vol = sums[SS_R3] * 4 * kPiOver3;             // shell.go:66-67
    props[0] = go_pow(vol / kFourPiOver3, 0.33333);            // cmd/stats.go:279 (sic: 0.33333)
    props[1] = vol;
    props[2] = sums[SS_AREA] * 4 * kPi;                       // shell.go:236
    const double norm = sums[SS_AREA];
    double m[9];
    for (int i = 0; i < 9; i++) m[i] = sums[SS_XX + i] / norm;   // shell.go:140-142
    axes_from_moments(m, props + 3);
    props[9] = sums[SS_RMIN];
    props[10] = sums[SS_RMAX];
}

}  // namespace sfk
