/* orc_analyze.c -- oracle (TEST INFRASTRUCTURE): los/analyze of shellfish
 * v1.0.4: smooth.go, splashback_radius.go, kde.go, penna.go.
 * PARITY UNPINNED for everything in this file: the reference has no known
 * answers for it and cannot be run here.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -ffast-math). */
#include "shellfish_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GO_2PI 6.28318530717958647692528676655900576839433879875021
#define GO_PI 3.14159265358979323846264338327950288419716939937511

/* --------------------------------------------------------------- Smooth */

/* los/analyze/smooth.go:46-81 (kernel lookup is the caller's job) */
int orc_smooth(const double *valTaps, const double *derivTaps, int window,
               double *ys, int n, double *vals, double *derivs) {
    if (n <= window) return 0;
    for (int i = 0; i < n; i++) ys[i] = log(ys[i]);
    orc_convolve_extension(valTaps, window, ys, n, vals);
    orc_convolve_extension(derivTaps, window, ys, n, derivs);
    for (int i = 0; i < n; i++) ys[i] = exp(ys[i]);
    for (int i = 0; i < n; i++) vals[i] = exp(vals[i]);
    return 1;
}

/* los/analyze/splashback_radius.go:30-63 */
int orc_splashback_radius(const double *rs, const double *rhos,
                          const double *derivs, int n, double dLim, double *r) {
    if (n == 0) return 0;
    double rhoMin = rhos[0];
    double dMin = dLim;
    int iMin = -1;
    for (int i = 1; i < n - 1; i++) {
        if (rhos[i] < rhoMin) {
            rhoMin = rhos[i];
            if ((derivs[i] < derivs[i + 1] && derivs[i] < derivs[i - 1]) &&
                (derivs[i] < dMin)) {
                dMin = derivs[i];
                iMin = i;
            }
        }
    }
    if (iMin == -1) return 0;
    *r = rs[iMin];
    return 1;
}

/* ------------------------------------------------------------------ KDE */

#define KDE_N 100

static inline int go_int(double x) { return (int)(int64_t)x; }

/* los/analyze/kde.go:14-41 */
static orc_spline *gaussian_kde(const double *xs, int nx, double h, double low,
                                double high) {
    const int n = KDE_N;
    double dx = (high - low) / (double)(n - 1);
    double spXs[KDE_N], spYs[KDE_N];
    for (int i = 0; i < n; i++) spYs[i] = 0;
    for (int i = 0; i < n - 1; i++) spXs[i] = low + dx * (double)i;
    spXs[n - 1] = high;
    double maxDist = h * 3;
    for (int k = 0; k < nx; k++) {
        double x = xs[k];
        int lowIdx = go_int((x - maxDist - low) / dx);
        int highIdx = go_int((x + maxDist - low) / dx) + 1;
        if (lowIdx < 0) lowIdx = 0;
        if (highIdx >= n) highIdx = n - 1;
        for (int i = lowIdx; i <= highIdx; i++) {
            double udx = (spXs[i] - x) / h;
            spYs[i] += exp(-udx * udx);
        }
    }
    return orc_spline_new(spXs, spYs, n);
}

#define MAX_LEVELS 12

typedef struct {
    double h, low, high;
    int levels;                     /* number of splits */
    orc_spline **sp[MAX_LEVELS + 1]; /* spTree[level][node] */
    double *th[MAX_LEVELS + 1];      /* thTree */
    double **maxes[MAX_LEVELS + 1];  /* maxesTree[level][node][k] */
    int *nMaxes[MAX_LEVELS + 1];
    double *conn[MAX_LEVELS + 1];    /* connMaxes */
    int nConn;                       /* levels of conn filled so far */
    double spRs[KDE_N];
    int err;
} kde_tree;

static void kt_free(kde_tree *kt) {
    for (int l = 0; l <= kt->levels; l++) {
        int nodes = 1 << l;
        if (kt->sp[l]) {
            for (int j = 0; j < nodes; j++) orc_spline_free(kt->sp[l][j]);
            free(kt->sp[l]);
        }
        if (kt->maxes[l]) {
            for (int j = 0; j < nodes; j++) free(kt->maxes[l][j]);
            free(kt->maxes[l]);
        }
        free(kt->nMaxes[l]);
        free(kt->th[l]);
        free(kt->conn[l]);
    }
}

/* los/analyze/kde.go:274-283, 262-270 */
static double kt_finest_max(const kde_tree *kt, int idx, int level) {
    for (int i = 0; i <= level; i++) {
        double r = kt->conn[level - i][idx / (1 << i)];
        if (!isnan(r)) return r;
    }
    return NAN; /* Go: panic(":3") -- unreachable, level 0 is never NaN */
}

/* los/analyze/kde.go:287-312 + 325-331 (Radial): spline r(theta) at a level */
static orc_spline *kt_rfunc(const kde_tree *kt, int level) {
    int n = 1 << level;
    double *maxes = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) maxes[i] = kt_finest_max(kt, i, level);
    const double *ths = kt->th[level];
    int buf = 5;
    if (buf > n) buf = n;
    int m = 2 * buf + n;
    double *spThs = (double *)malloc(sizeof(double) * (size_t)m * 2);
    double *spMaxes = spThs + m;
    int j = n - buf;
    for (int i = 0; i < buf; i++) { spThs[i] = ths[j] - GO_2PI; spMaxes[i] = maxes[j]; j++; }
    j = 0;
    for (int i = buf; i < n + buf; i++) { spThs[i] = ths[j]; spMaxes[i] = maxes[j]; j++; }
    j = 0;
    for (int i = n + buf; i < n + 2 * buf; i++) { spThs[i] = ths[j] + GO_2PI; spMaxes[i] = maxes[j]; j++; }
    orc_spline *sp = orc_spline_new(spThs, spMaxes, m);
    free(spThs);
    free(maxes);
    return sp;
}

/* los/analyze/kde.go:193-206 */
static int local_spline_maxes(const double *xs, int n, const orc_spline *sp,
                              double *out, int *err) {
    double prev, curr, next;
    int k = 0;
    if (orc_spline_eval(sp, xs[0], &prev) || orc_spline_eval(sp, xs[1], &curr) ||
        orc_spline_eval(sp, xs[2], &next)) { *err = ORC_E_SPLINE; return 0; }
    if (curr > next && curr > prev) out[k++] = xs[1];
    for (int i = 2; i < n - 1; i++) {
        prev = curr; curr = next;
        if (orc_spline_eval(sp, xs[i + 1], &next)) { *err = ORC_E_SPLINE; return k; }
        if (curr > next && curr > prev) out[k++] = xs[i];
    }
    return k;
}

/* los/analyze/kde.go:66-110 NewKDETree + growTrees/findMaxes/connectMaxes */
static int kt_build(kde_tree *kt, const double *rs, const double *phis, int n,
                    int splits, double h) {
    memset(kt, 0, sizeof *kt);
    if (splits > MAX_LEVELS) return ORC_E_SPLINE;
    kt->levels = splits;
    if (n == 0) return ORC_E_NO_POINTS;
    kt->low = 0; kt->high = rs[0];
    for (int i = 0; i < n; i++) if (rs[i] > kt->high) kt->high = rs[i];
    const int rn = KDE_N;
    double dr = (kt->high - kt->low) / (double)rn;
    for (int i = 0; i < rn; i++) kt->spRs[i] = kt->low + dr * (double)i;
    kt->spRs[rn - 1] = kt->high;
    kt->h = h;

    kt->sp[0] = (orc_spline **)calloc(1, sizeof(orc_spline *));
    kt->sp[0][0] = gaussian_kde(rs, n, h, kt->low, kt->high);
    kt->th[0] = (double *)malloc(sizeof(double));
    kt->th[0][0] = GO_PI;

    /* growTrees + binByTheta, kde.go:140-174 */
    double *rBin = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int split = 0; split < splits; split++) {
        int bins = 1 << (1 + split);
        int l = split + 1;
        kt->sp[l] = (orc_spline **)calloc((size_t)bins, sizeof(orc_spline *));
        kt->th[l] = (double *)malloc(sizeof(double) * (size_t)bins);
        double dth = GO_2PI / (double)bins;
        for (int b = 0; b < bins; b++) {
            int m = 0;
            for (int i = 0; i < n; i++)
                if (go_int(phis[i] / dth) == b) rBin[m++] = rs[i];
            kt->sp[l][b] = gaussian_kde(rBin, m, h, kt->low, kt->high);
            kt->th[l][b] = GO_2PI * ((double)b + 0.5) / (double)bins;
        }
    }
    free(rBin);

    /* findMaxes, kde.go:178-189 */
    for (int l = 0; l <= splits; l++) {
        int nodes = 1 << l;
        kt->maxes[l] = (double **)calloc((size_t)nodes, sizeof(double *));
        kt->nMaxes[l] = (int *)calloc((size_t)nodes, sizeof(int));
        for (int j = 0; j < nodes; j++) {
            kt->maxes[l][j] = (double *)malloc(sizeof(double) * KDE_N);
            int err = 0;
            kt->nMaxes[l][j] = local_spline_maxes(kt->spRs, KDE_N, kt->sp[l][j],
                                                  kt->maxes[l][j], &err);
            if (err) return err;
        }
    }

    /* connectMaxes, kde.go:211-258 */
    if (kt->nMaxes[0][0] == 0) return ORC_E_NO_MAXIMUM;
    kt->conn[0] = (double *)malloc(sizeof(double));
    kt->conn[0][0] = kt->maxes[0][0][0];
    kt->nConn = 1;
    for (int split = 0; split < splits; split++) {
        int l = split + 1, nodes = 1 << l;
        const double *prevMaxes = kt->conn[split];
        double *currMaxes = (double *)malloc(sizeof(double) * (size_t)nodes);
        for (int i = 0; i < nodes; i++) currMaxes[i] = NAN;
        for (int node = 0; node < nodes; node++) {
            const double *nodeMaxes = kt->maxes[l][node];
            int nm = kt->nMaxes[l][node];
            double nodePrevMax = prevMaxes[node / 2];
            double connMax;
            if (isnan(nodePrevMax) || nm == 0) {
                connMax = NAN;
            } else {
                int connIdx = -1;
                double minDist = INFINITY;
                for (int i = 0; i < nm; i++) {
                    double dist = fabs(nodePrevMax - nodeMaxes[i]);
                    if (dist < minDist) { connIdx = i; minDist = dist; }
                }
                connMax = nodeMaxes[connIdx];
            }
            if (!isnan(connMax) && (split == 0 || fabs(connMax - nodePrevMax) < kt->h)) {
                currMaxes[node] = connMax;
            } else {
                for (int i = 0; i < nm; i++) {
                    orc_spline *rf = kt_rfunc(kt, split);
                    double spR;
                    int bad = orc_spline_eval(rf, kt->th[l][node], &spR);
                    orc_spline_free(rf);
                    if (bad) { free(currMaxes); return ORC_E_SPLINE; }
                    if (fabs(nodeMaxes[i] - spR) < kt->h) currMaxes[node] = nodeMaxes[i];
                }
            }
        }
        kt->conn[l] = currMaxes;
        kt->nConn = l + 1;
    }
    return 0;
}

/* los/analyze/penna.go:115-165 for one ring, with kde.go:355-369 FilterNearby */
int orc_filter_ring(const double *rs, const double *phis, int n, int levels,
                    double h0, double *fRs, double *fPhis, int *fIdx,
                    double *hUsed) {
    double factor = 1.0;
    int kept = 0;
    for (int it = 0; it < 10 && kept == 0; it++) {
        kde_tree kt;
        int rc = kt_build(&kt, rs, phis, n, levels, h0 * factor);
        if (rc != 0) { kt_free(&kt); return rc; }
        orc_spline *rf = kt_rfunc(&kt, levels);
        double dr = kt.h;
        for (int i = 0; i < n; i++) {
            double v;
            if (orc_spline_eval(rf, phis[i], &v)) {
                orc_spline_free(rf); kt_free(&kt); return ORC_E_SPLINE;
            }
            if (fabs(v - rs[i]) < dr / 2) {
                fRs[kept] = rs[i];
                fPhis[kept] = phis[i];
                if (fIdx) fIdx[kept] = i;
                kept++;
            }
        }
        if (hUsed) *hUsed = kt.h;
        orc_spline_free(rf);
        kt_free(&kt);
        factor *= 1.1;
    }
    return kept;
}

/* ---------------------------------------------------------------- Penna */

/* gonum/lapack/native Dgetf2 (unblocked, row-major; gonum/matrix pins
 * gonum/lapack e4cdc5a0bff9, go.mod:8): partial pivoting with the first
 * largest |a| in the column, scale by the reciprocal pivot, rank-1 update. */
static void dgetf2(double *a, int n, int *ipiv) {
    const double sfmin = 2.2250738585072014e-308;
    for (int j = 0; j < n; j++) {
        int jp = j;
        double amax = fabs(a[j * n + j]);
        for (int i = j + 1; i < n; i++) {
            double v = fabs(a[i * n + j]);
            if (v > amax) { amax = v; jp = i; }
        }
        ipiv[j] = jp;
        if (a[jp * n + j] != 0) {
            if (jp != j)
                for (int k = 0; k < n; k++) {
                    double t = a[j * n + k]; a[j * n + k] = a[jp * n + k]; a[jp * n + k] = t;
                }
            if (j < n - 1) {
                double aj = a[j * n + j];
                if (fabs(aj) >= sfmin) {
                    double inv = 1 / aj;
                    for (int i = j + 1; i < n; i++) a[i * n + j] *= inv;
                } else {
                    for (int i = j + 1; i < n; i++) a[i * n + j] = a[i * n + j] / aj;
                }
            }
        }
        if (j < n - 1) {
            /* Dger: A22 += -1 * x y^T, row by row */
            for (int i = j + 1; i < n; i++) {
                double tmp = -1 * a[i * n + j];
                if (tmp == 0) continue;
                for (int k = j + 1; k < n; k++) {
                    double p = tmp * a[j * n + k];
                    a[i * n + k] += p;
                }
            }
        }
    }
}

/* gonum/lapack/native Dtrti2(Upper, NonUnit) + Dgetri unblocked branch,
 * i.e. reference LAPACK DTRTI2/DGETRI stated for row-major storage. */
static void dgetri(double *a, int n, const int *ipiv) {
    double *work = (double *)calloc((size_t)n, sizeof(double));
    /* inv(U): for each column j, x := U(0:j,0:j) * U(0:j,j) then scale */
    for (int j = 0; j < n; j++) {
        double ajj = 1 / a[j * n + j];
        a[j * n + j] = ajj;
        ajj *= -1;
        /* Dtrmv upper, no-trans, non-unit on the leading j x j block applied
         * to the vector a[0:j, j] (stride n): x_i = sum_{k>=i} U_ik x_k */
        for (int i = 0; i < j; i++) {
            double tmp = 0;
            for (int k = i; k < j; k++) {
                double p = a[i * n + k] * a[k * n + j];
                tmp += p;
            }
            a[i * n + j] = tmp;
        }
        for (int i = 0; i < j; i++) a[i * n + j] *= ajj;
    }
    /* Solve inv(A)*L = inv(U) for inv(A), last column first */
    for (int j = n - 1; j >= 0; j--) {
        for (int i = j + 1; i < n; i++) {
            work[i] = a[i * n + j];
            a[i * n + j] = 0;
        }
        if (j < n - 1) {
            /* Dgemv: a[:,j] += -1 * A[:, j+1:n] * work[j+1:n] */
            for (int i = 0; i < n; i++) {
                double tmp = 0;
                for (int k = j + 1; k < n; k++) {
                    double p = a[i * n + k] * work[k];
                    tmp += p;
                }
                a[i * n + j] += -1 * tmp;
            }
        }
    }
    for (int j = n - 2; j >= 0; j--) {
        int jp = ipiv[j];
        if (jp != j)
            for (int i = 0; i < n; i++) {
                double t = a[i * n + j]; a[i * n + j] = a[i * n + jp]; a[i * n + jp] = t;
            }
    }
    free(work);
}

/* los/analyze/penna.go:14-58 + pinv :170-194 + mat.VecMult math/mat/mat.go:102 */
int orc_penna_coeffs(const double *xs, const double *ys, const double *zs,
                     int64_t N, int I, int J, int K, double *cs) {
    int P = I * J * K;
    double *rs = (double *)malloc(sizeof(double) * (size_t)N * 5);
    double *cosths = rs + N, *sinths = cosths + N, *cosphis = sinths + N,
           *sinphis = cosphis + N;
    for (int64_t i = 0; i < N; i++) {
        rs[i] = sqrt(xs[i] * xs[i] + ys[i] * ys[i] + zs[i] * zs[i]);
        cosths[i] = zs[i] / rs[i];
        sinths[i] = sqrt(1 - cosths[i] * cosths[i]);
        cosphis[i] = xs[i] / rs[i] / sinths[i];
        sinphis[i] = ys[i] / rs[i] / sinths[i];
    }
    /* M is P x N row-major (mat.NewMatrix(MVals, N, P): Width=N) */
    double *M = (double *)malloc(sizeof(double) * (size_t)P * (size_t)N);
    for (int64_t n = 0; n < N; n++) {
        int m = 0;
        for (int k = 0; k < K; k++) {
            double costh = orc_go_pow(cosths[n], (double)k);
            for (int j = 0; j < J; j++) {
                double sinphi = orc_go_pow(sinphis[n], (double)j);
                double cosphi = 1.0;
                for (int i = 0; i < I; i++) {
                    M[(int64_t)m * N + n] =
                        orc_go_pow(sinths[n], (double)(i + j)) * cosphi * costh * sinphi;
                    m++;
                    cosphi *= cosphis[n];
                }
            }
        }
    }
    /* out1 = M * M^T (gonum Dense.Mul: row-by-row axpy, l ascending) */
    double *G = (double *)calloc((size_t)P * (size_t)P, sizeof(double));
    for (int i = 0; i < P; i++)
        for (int64_t l = 0; l < N; l++) {
            double tmp = M[(int64_t)i * N + l];
            if (tmp == 0) continue;
            for (int j = 0; j < P; j++) {
                double p = tmp * M[(int64_t)j * N + l];
                G[i * P + j] += p;
            }
        }
    /* inv = Inverse(out1): Dgetrf (n < block size -> Dgetf2) + Dgetri */
    int *ipiv = (int *)malloc(sizeof(int) * (size_t)P);
    dgetf2(G, P, ipiv);
    int singular = 0;
    for (int i = 0; i < P; i++) if (G[i * P + i] == 0) singular = 1;
    if (!singular) dgetri(G, P, ipiv);
    /* out2 = M^T * inv (N x P), then cs = rs^T * out2 (mat.VecMult) */
    for (int i = 0; i < P; i++) cs[i] = 0;
    if (!singular) {
        double *out2 = (double *)calloc((size_t)N * (size_t)P, sizeof(double));
        for (int64_t n = 0; n < N; n++)
            for (int l = 0; l < P; l++) {
                double tmp = M[(int64_t)l * N + n];
                if (tmp == 0) continue;
                for (int j = 0; j < P; j++) {
                    double p = tmp * G[l * P + j];
                    out2[n * P + j] += p;
                }
            }
        for (int i = 0; i < P; i++) {
            double sum = 0.0;
            for (int64_t j = 0; j < N; j++) {
                double p = rs[j] * out2[i + (int64_t)P * j];
                sum += p;
            }
            cs[i] = sum;
        }
        free(out2);
    }
    free(ipiv); free(G); free(M); free(rs);
    return singular ? -1 : 0;
}

/* los/analyze/penna.go:62-82 */
double orc_penna_func(const double *cs, int I, int J, int K, double phi, double th) {
    int idx = 0;
    double sum = 0.0;
    double sinPhi = sin(phi), cosPhi = cos(phi);
    double sinTh = sin(th), cosTh = cos(th);
    for (int k = 0; k < K; k++) {
        double cosK = orc_go_pow(cosTh, (double)k);
        for (int j = 0; j < J; j++) {
            double sinJ = orc_go_pow(sinPhi, (double)j);
            for (int i = 0; i < I; i++) {
                double cosI = orc_go_pow(cosPhi, (double)i);
                double sinIJ = orc_go_pow(sinTh, (double)(i + j));
                sum += cs[idx] * sinIJ * cosK * sinJ * cosI;
                idx++;
            }
        }
    }
    return sum;
}
