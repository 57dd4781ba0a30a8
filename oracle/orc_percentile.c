/* orc_percentile.c -- oracle (TEST INFRASTRUCTURE): the PercentileProfile
 * output mode, cmd/shell.go:629-660, with the selection routine of
 * math/sort/median.go:8-103 and math/sort/quick.go:7-78 restated literally
 * (the partition scheme matters only for NaN inputs, which GetRhos never
 * produces, but it is what the reference runs).
 * Pinned by the property the reference's own test uses (sort_test.go:274-293:
 * NthLargest(xs, j) == sorted[len-j] for every j). */
#include "shellfish_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* quick.go:7-31: largest, middle, smallest */
static void sort3(double x, double y, double z, double *mx, double *md, double *mn) {
    if (x > y) {
        if (x > z) {
            if (y > z) { *mx = x; *md = y; *mn = z; } else { *mx = x; *md = z; *mn = y; }
        } else { *mx = z; *md = x; *mn = y; }
    } else {
        if (y > z) {
            if (x > z) { *mx = y; *md = x; *mn = z; } else { *mx = y; *md = z; *mn = x; }
        } else { *mx = z; *md = y; *mn = x; }
    }
}

/* quick.go:49-78 */
static int64_t partition(double *xs, int64_t n) {
    int64_t n2 = n / 2;
    double mx, md, mn, t;
    sort3(xs[0], xs[n2], xs[n - 1], &mx, &md, &mn);
    xs[0] = mn; xs[n2] = md; xs[n - 1] = mx;
    t = xs[1]; xs[1] = xs[n2]; xs[n2] = t;
    int64_t lo = 1, hi = n - 1;
    for (;;) {
        lo++;
        while (xs[lo] < md) lo++;
        hi--;
        while (xs[hi] > md) hi--;
        if (hi < lo) break;
        t = xs[lo]; xs[lo] = xs[hi]; xs[hi] = t;
    }
    t = xs[1]; xs[1] = xs[hi]; xs[hi] = t;
    return hi;
}

/* median.go:67-103 (1-indexed n; the Go recursion is a tail call on a sub-slice) */
static double nth_largest_inplace(double *xs, int64_t len, int64_t n) {
    for (;;) {
        if (len == 1) return xs[0];
        if (len == 2) {
            if (xs[0] > xs[1]) { double t = xs[0]; xs[0] = xs[1]; xs[1] = t; }
            return n <= 2 ? xs[2 - n] : NAN;   /* Go panics for n > 2 */
        }
        if (len == 3) {
            double mx, md, mn;
            sort3(xs[0], xs[1], xs[2], &mx, &md, &mn);
            return n == 1 ? mx : n == 2 ? md : n == 3 ? mn : NAN;
        }
        int64_t piv = partition(xs, len);
        int64_t nPiv = len - piv;
        if (nPiv > n) { xs += piv; len = nPiv; }
        else if (nPiv < n) { len = piv; n -= nPiv; }
        else return xs[piv];
    }
}

/* median.go:40-65: buf is scratch of the same length */
double orc_nth_largest(const double *xs, int64_t len, int64_t n, double *buf) {
    memcpy(buf, xs, sizeof(double) * (size_t)len);
    return nth_largest_inplace(buf, len, n);
}

/* median.go:8-22 */
double orc_percentile(const double *xs, int64_t len, double p, double *buf) {
    int64_t n = (int64_t)(p * (double)len);
    if (n == 0) n++;
    return orc_nth_largest(xs, len, n, buf);
}

/* calcPercentile, cmd/shell.go:629-660.  profiles: [nLos][bins] GetRhos of
 * every LoS (ring-major), out: [2*bins] = bin radii (GetRs, halo.go:360-370)
 * then the percentile density of every radial bin. */
void orc_calc_percentile(double rMin, double rMax, int bins, const double *profiles,
                         int64_t nLos, double percentile, double *out) {
    double dlr = (log(rMax) - log(rMin)) / (double)bins;
    double lrMin = log(rMin);
    for (int i = 0; i < bins; i++) out[i] = exp(lrMin + dlr * ((double)i + 0.5));
    double *rBuf = (double *)malloc(sizeof(double) * (size_t)nLos * 2);
    double *medBuf = rBuf + nLos;
    for (int ri = 0; ri < bins; ri++) {
        for (int64_t i = 0; i < nLos; i++) rBuf[i] = profiles[(size_t)i * (size_t)bins + ri];
        out[bins + ri] = orc_percentile(rBuf, nLos, percentile / 100, medBuf);
    }
    free(rBuf);
}
