/* orc_los.c -- oracle (TEST INFRASTRUCTURE): los.ProfileRing and los.Halo.
 * Restates los/profile_ring.go and los/halo.go of shellfish v1.0.4.
 * Build: gcc -std=c11 -O2 -ffp-contract=off (no -ffast-math). */
#include "shellfish_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GO_2PI 6.28318530717958647692528676655900576839433879875021

/* ---------------------------------------------------------- ProfileRing */

/* "Exact bins" variant (orc_set_exact_bins), NOT the reference's arithmetic: the same algorithm
 * with the bin sums carried exactly.  The reference adds rho*(1-rem) and rho*rem, two rounded
 * products whose sum is rho only up to an ulp, so two stretches of a profile with the same true
 * density (the plateau inside one particle's kernel, twice along a sparse line of sight) differ
 * by rounding noise, and SplashbackRadius' strict comparisons then decide on that noise.  The
 * CUDA path accumulates exact integers (rho - q, q).  This variant lets the tests separate the
 * two effects: it adds (rho - a, a) with a = rho*rem and keeps every bin as a double-double. */
static int g_exact_bins = 0;
void orc_set_exact_bins(int on) { g_exact_bins = on; }

struct orc_ring {
    double *derivs; /* [n][bins], LoS-major (profile_ring.go:35) */
    double *comp;   /* exact-bins variant: low words of the double-double sums, else NULL */
    int bins, n;
    double lowR, highR, dr;
    /* parity taps, owned by the halo (NULL when disabled) */
    uint32_t *tapInserts, *tapStart, *tapEnd;
    int64_t oob; /* writes Go would have rejected with an index panic */
};

/* los/profile_ring.go:75-82 */
static void ring_init(orc_ring *p, double lowR, double highR, int bins, int n) {
    memset(p, 0, sizeof *p);
    p->derivs = (double *)calloc((size_t)bins * (size_t)n, sizeof(double));
    p->comp = g_exact_bins ? (double *)calloc((size_t)bins * (size_t)n, sizeof(double)) : NULL;
    p->bins = bins;
    p->n = n;
    p->lowR = lowR;
    p->highR = highR;
    p->dr = (highR - lowR) / (double)bins;
}

orc_ring *orc_ring_new(double lowR, double highR, int bins, int n) {
    orc_ring *p = (orc_ring *)malloc(sizeof *p);
    ring_init(p, lowR, highR, bins, n);
    return p;
}

void orc_ring_free(orc_ring *p) {
    if (!p) return;
    free(p->derivs);
    free(p->comp);
    free(p);
}

const double *orc_ring_derivs(const orc_ring *p) { return p->derivs; }

/* One `p.derivs[k] += v` where k may be one past this LoS's row when modf
 * returns idx == bins (profile_ring.go:102-106 has no upper check).  Go only
 * panics when k runs off the whole slice. */
static inline void dd_add(double *hi, double *lo, double v) {   /* Knuth two-sum */
    double s = *hi + v;
    double bb = s - *hi;
    double e = (*hi - (s - bb)) + (v - bb);
    *hi = s;
    *lo += e;
}

static inline void ring_add(orc_ring *p, int64_t k, double v) {
    if (k < 0 || k >= (int64_t)p->bins * p->n) { p->oob++; return; }
    if (p->comp) dd_add(&p->derivs[k], &p->comp[k], v);
    else p->derivs[k] += v;
}

/* exact-bins variant of `derivs[idx] += sign*rho*(1-rem); derivs[idx+1] += sign*rho*rem` */
static inline void ring_add_split_exact(orc_ring *p, int64_t row, int idx, double rho, double rem, double sign) {
    double a = rho * rem;
    double s = rho - a;
    double bb = s - rho;
    double e = (rho - (s - bb)) + (-a - bb);   /* rho - a == s + e exactly */
    ring_add(p, row + idx, sign * s);
    ring_add(p, row + idx, sign * e);
    if (idx < p->bins - 1) ring_add(p, row + idx + 1, sign * a);
}

/* los/profile_ring.go:92-120 */
void orc_ring_insert(orc_ring *p, double start, double end, double rho, int i) {
    if (end <= p->lowR || start >= p->highR) return;
    int64_t row = (int64_t)i * p->bins;
    if (p->tapInserts) p->tapInserts[i]++;

    if (start > p->lowR) {
        double fidx;
        double rem = modf((start - p->lowR) / p->dr, &fidx);
        int idx = (int)fidx;
        if (p->comp) ring_add_split_exact(p, row, idx, rho, rem, 1.0);
        else {
        ring_add(p, row + idx, rho * (1 - rem));
        if (idx < p->bins - 1) ring_add(p, row + idx + 1, rho * rem);
        }
        if (p->tapStart && idx < p->bins) p->tapStart[row + idx]++;
    } else {
        ring_add(p, row, rho);
        if (p->tapStart) p->tapStart[row]++;
    }

    if (end < p->highR) {
        double fidx;
        double rem = modf((end - p->lowR) / p->dr, &fidx);
        int idx = (int)fidx;
        if (p->comp) ring_add_split_exact(p, row, idx, rho, rem, -1.0);
        else {
        ring_add(p, row + idx, -(rho * (1 - rem)));
        if (idx < p->bins - 1) ring_add(p, row + idx + 1, -(rho * rem));
        }
        if (p->tapEnd && idx < p->bins) p->tapEnd[row + idx]++;
    }
}

/* los/profile_ring.go:124-130 */
void orc_ring_retrieve(const orc_ring *p, int i, double *out) {
    double sum = 0, lo = 0;
    for (int j = 0; j < p->bins; j++) {
        if (p->comp) {
            dd_add(&sum, &lo, p->derivs[j + p->bins * i]);
            lo += p->comp[j + p->bins * i];
            out[j] = sum + lo;
        } else {
            sum += p->derivs[j + p->bins * i];
            out[j] = sum;
        }
    }
}

/* ----------------------------------------------------------------- Halo */

struct orc_halo {
    double origin[3];
    int rings, bins, n;
    double rMin, rMax;
    double *ringVecs; /* [n][2] = (cos, sin) */
    double *ringPhis;
    double dPhi;
    float *rots, *irots; /* [rings][9] */
    float *norms;        /* [rings][3] */
    orc_ring *profs;
    double defaultRho;
    uint32_t *tapInserts, *tapStart, *tapEnd;
};

/* los/halo.go:65-100 */
orc_halo *orc_halo_new(const float *norms, int rings, const double origin[3],
                       double rMin, double rMax, int bins, int n,
                       double defaultRho) {
    orc_halo *h = (orc_halo *)calloc(1, sizeof *h);
    memcpy(h->origin, origin, sizeof h->origin);
    h->rMin = rMin; h->rMax = rMax;
    h->rings = rings; h->bins = bins; h->n = n;
    h->norms = (float *)malloc(sizeof(float) * 3 * (size_t)rings);
    memcpy(h->norms, norms, sizeof(float) * 3 * (size_t)rings);
    h->defaultRho = defaultRho;
    const float zAxis[3] = {0, 0, 1};
    h->profs = (orc_ring *)calloc((size_t)rings, sizeof(orc_ring));
    h->rots = (float *)calloc((size_t)rings * 9, sizeof(float));
    h->irots = (float *)calloc((size_t)rings * 9, sizeof(float));
    for (int i = 0; i < rings; i++) {
        ring_init(&h->profs[i], log(rMin), log(rMax), bins, n);
        orc_euler_matrix_between(&h->norms[3 * i], zAxis, &h->rots[9 * i]);
        orc_euler_matrix_between(zAxis, &h->norms[3 * i], &h->irots[9 * i]);
    }
    h->ringPhis = (double *)malloc(sizeof(double) * (size_t)n);
    h->ringVecs = (double *)malloc(sizeof(double) * 2 * (size_t)n);
    for (int i = 0; i < n; i++) {
        h->ringPhis[i] = (double)i / (double)n * GO_2PI;
        /* math.Sincos: sin and cos of the same argument */
        h->ringVecs[2 * i + 1] = sin(h->ringPhis[i]);
        h->ringVecs[2 * i + 0] = cos(h->ringPhis[i]);
    }
    h->dPhi = 1 / (double)n * GO_2PI;
    return h;
}

void orc_halo_free(orc_halo *h) {
    if (!h) return;
    for (int i = 0; i < h->rings; i++) { free(h->profs[i].derivs); free(h->profs[i].comp); }
    free(h->profs); free(h->rots); free(h->irots); free(h->norms);
    free(h->ringPhis); free(h->ringVecs);
    free(h->tapInserts); free(h->tapStart); free(h->tapEnd);
    free(h);
}

void orc_halo_enable_taps(orc_halo *h) {
    size_t rn = (size_t)h->rings * (size_t)h->n;
    if (h->tapInserts) return;
    h->tapInserts = (uint32_t *)calloc(rn, sizeof(uint32_t));
    h->tapStart = (uint32_t *)calloc(rn * (size_t)h->bins, sizeof(uint32_t));
    h->tapEnd = (uint32_t *)calloc(rn * (size_t)h->bins, sizeof(uint32_t));
    for (int r = 0; r < h->rings; r++) {
        h->profs[r].tapInserts = h->tapInserts + (size_t)r * h->n;
        h->profs[r].tapStart = h->tapStart + (size_t)r * h->n * h->bins;
        h->profs[r].tapEnd = h->tapEnd + (size_t)r * h->n * h->bins;
    }
}

const uint32_t *orc_halo_tap_inserts(const orc_halo *h) { return h->tapInserts; }
const uint32_t *orc_halo_tap_start(const orc_halo *h) { return h->tapStart; }
const uint32_t *orc_halo_tap_end(const orc_halo *h) { return h->tapEnd; }
const double *orc_halo_derivs(const orc_halo *h, int ring) { return h->profs[ring].derivs; }
const float *orc_halo_rot(const orc_halo *h, int ring) { return h->rots + 9 * ring; }
const float *orc_halo_irot(const orc_halo *h, int ring) { return h->irots + 9 * ring; }

/* los/halo.go:129-140 + profile_ring.go:44-52 */
void orc_halo_join(orc_halo *dst, const orc_halo *src) {
    size_t m = (size_t)dst->n * (size_t)dst->bins;
    for (int r = 0; r < dst->rings; r++) {
        double *a = dst->profs[r].derivs;
        const double *b = src->profs[r].derivs;
        if (dst->profs[r].comp && src->profs[r].comp) {
            for (size_t i = 0; i < m; i++) {
                dd_add(&a[i], &dst->profs[r].comp[i], b[i]);
                dst->profs[r].comp[i] += src->profs[r].comp[i];
            }
        } else
        for (size_t i = 0; i < m; i++) a[i] += b[i];
    }
    if (dst->tapInserts && src->tapInserts) {
        size_t rn = (size_t)dst->rings * (size_t)dst->n;
        for (size_t i = 0; i < rn; i++) dst->tapInserts[i] += src->tapInserts[i];
        for (size_t i = 0; i < rn * (size_t)dst->bins; i++) {
            dst->tapStart[i] += src->tapStart[i];
            dst->tapEnd[i] += src->tapEnd[i];
        }
    }
}

/* los/halo.go:107-123 on an equally sized worker: zero its grid */
void orc_halo_clear(orc_halo *h) {
    size_t m = (size_t)h->n * (size_t)h->bins;
    for (int r = 0; r < h->rings; r++) {
        memset(h->profs[r].derivs, 0, m * sizeof(double));
        if (h->profs[r].comp) memset(h->profs[r].comp, 0, m * sizeof(double));
    }
    if (h->tapInserts) {
        size_t rn = (size_t)h->rings * (size_t)h->n;
        memset(h->tapInserts, 0, rn * sizeof(uint32_t));
        memset(h->tapStart, 0, rn * (size_t)h->bins * sizeof(uint32_t));
        memset(h->tapEnd, 0, rn * (size_t)h->bins * sizeof(uint32_t));
    }
}

/* los/halo.go:168-197 -- mutates xs in place, float32 throughout */
void orc_halo_transform(const orc_halo *h, float *xs, int64_t n, double totalWidth) {
    float x0 = (float)h->origin[0], y0 = (float)h->origin[1], z0 = (float)h->origin[2];
    float tw = (float)totalWidth;
    float tw2 = tw / 2;
    for (int64_t i = 0; i < n; i++) {
        float x = xs[3 * i], y = xs[3 * i + 1], z = xs[3 * i + 2];
        float dx = x - x0, dy = y - y0, dz = z - z0;
        if (dx > tw2) xs[3 * i] -= tw; else if (dx < -tw2) xs[3 * i] += tw;
        if (dy > tw2) xs[3 * i + 1] -= tw; else if (dy < -tw2) xs[3 * i + 1] += tw;
        if (dz > tw2) xs[3 * i + 2] -= tw; else if (dz < -tw2) xs[3 * i + 2] += tw;
    }
}

/* los/halo.go:147-164 */
void orc_halo_intersect(const orc_halo *h, const float *xs, int64_t n, double r,
                        uint8_t *intr) {
    double rMin = h->rMin - r, rMax = h->rMax + r;
    if (rMin < 0) rMin = 0;
    float rMin2 = (float)(rMin * rMin), rMax2 = (float)(rMax * rMax);
    float x0 = (float)h->origin[0], y0 = (float)h->origin[1], z0 = (float)h->origin[2];
    for (int64_t i = 0; i < n; i++) {
        float x = xs[3 * i] - x0, y = xs[3 * i + 1] - y0, z = xs[3 * i + 2] - z0;
        float xx = x * x, yy = y * y, zz = z * z;
        float r2 = (xx + yy) + zz;
        intr[i] = (r2 > rMin2 && r2 < rMax2) ? 1 : 0;
    }
}

/* los/halo.go:218-224 */
int orc_halo_sphere_intersect_ring(const orc_halo *h, const float vec[3],
                                   double radius, int ring) {
    const float *norm = &h->norms[3 * ring];
    float a = norm[0] * vec[0], b = norm[1] * vec[1], c = norm[2] * vec[2];
    float d = (a + b) + c;
    double dot = (double)d;
    return dot < radius && dot > -radius;
}

/* los/halo.go:315-345 */
double orc_half_angular_width(double dist2, double r2) {
    return asin(sqrt(r2 / dist2));
}

void orc_two_val_intr_dist(double dist2, double rad2, double b, double *lo, double *hi) {
    double b2 = b * b;
    double midDist = sqrt(dist2 - b2);
    double diff = sqrt(rad2 - b2);
    *lo = midDist - diff;
    *hi = midDist + diff;
}

double orc_one_val_intr_dist(double dist2, double rad2, double b, double dir) {
    double b2 = b * b;
    double radMidDist = sqrt(rad2 - b2);
    double cMidDist = sqrt(dist2 - b2);
    if (dir > 0) return radMidDist + cMidDist;
    return radMidDist - cMidDist;
}

/* Go's int(float64) on amd64 is CVTTSD2SQ: truncation toward zero. */
static inline int go_int(double x) { return (int)(int64_t)x; }

/* los/halo.go:284-310 */
void orc_halo_idx_range(const orc_halo *h, double phiLo, double phiHi, int out[4]) {
    if (phiHi > GO_2PI) {
        out[0] = go_int(phiLo / h->dPhi);
        out[1] = h->n;
        out[2] = 0;
        out[3] = go_int((phiHi - GO_2PI) / h->dPhi) + 1;
    } else if (phiLo < 0) {
        out[0] = go_int((phiLo + GO_2PI) / h->dPhi);
        out[1] = h->n;
        out[2] = 0;
        out[3] = go_int(phiHi / h->dPhi) + 1;
    } else {
        out[0] = go_int(phiLo / h->dPhi);
        out[1] = go_int(phiHi / h->dPhi) + 1;
        out[2] = 0;
        out[3] = 0;
    }
}

/* los/halo.go:228-277 */
void orc_halo_insert_to_ring(orc_halo *h, const float vecIn[3], double radius,
                             double rho, int ring) {
    float vec[3] = {vecIn[0], vecIn[1], vecIn[2]};
    orc_rotate_vec(vec, &h->rots[9 * ring]);
    orc_ring *prof = &h->profs[ring];

    double cx = (double)vec[0], cy = (double)vec[1], cz = (double)vec[2];
    double projDist2 = cx * cx + cy * cy;
    double projRad2 = radius * radius - cz * cz;
    if (projRad2 < 0) projRad2 = 0;

    if (projRad2 > projDist2) {
        for (int i = 0; i < h->n; i++) {
            double c = h->ringVecs[2 * i], s = h->ringVecs[2 * i + 1];
            double b = cy * c - cx * s;
            double dir = cx * c + cy * s;
            double rHi = orc_one_val_intr_dist(projDist2, projRad2, b, dir);
            orc_ring_insert(prof, -INFINITY, log(rHi), rho, i);
        }
    } else {
        double alpha = orc_half_angular_width(projDist2, projRad2);
        double projPhi = atan2(cy, cx);
        double phiStart = projPhi - alpha, phiEnd = projPhi + alpha;
        int r[4];
        orc_halo_idx_range(h, phiStart, phiEnd, r);
        for (int pass = 0; pass < 2; pass++) {
            for (int i = r[2 * pass]; i < r[2 * pass + 1]; i++) {
                double c = h->ringVecs[2 * i], s = h->ringVecs[2 * i + 1];
                double b = cy * c - cx * s;
                double rLo, rHi;
                orc_two_val_intr_dist(projDist2, projRad2, b, &rLo, &rHi);
                if (isnan(rLo) || isnan(rHi)) continue;
                orc_ring_insert(prof, log(rLo), log(rHi), rho, i);
            }
        }
    }
}

/* los/halo.go:201-215 */
void orc_halo_insert(orc_halo *h, const float vecIn[3], double radius, double rho) {
    float vec[3];
    vec[0] = vecIn[0] - (float)h->origin[0];
    vec[1] = vecIn[1] - (float)h->origin[1];
    vec[2] = vecIn[2] - (float)h->origin[2];
    for (int ring = 0; ring < h->rings; ring++) {
        if (orc_halo_sphere_intersect_ring(h, vec, radius, ring))
            orc_halo_insert_to_ring(h, vec, radius, rho, ring);
    }
}

/* los/halo.go:350-357 */
void orc_halo_get_rhos(const orc_halo *h, int ring, int los, double *buf) {
    orc_ring_retrieve(&h->profs[ring], los, buf);
    for (int i = 0; i < h->bins; i++)
        if (buf[i] < h->defaultRho) buf[i] = h->defaultRho;
}

/* los/halo.go:360-370 */
void orc_halo_get_rs(const orc_halo *h, double *buf) {
    double dlr = (log(h->rMax) - log(h->rMin)) / (double)h->bins;
    double lrMin = log(h->rMin);
    for (int i = 0; i < h->bins; i++) buf[i] = exp(lrMin + dlr * ((double)i + 0.5));
}

/* los/halo.go:442-467 */
static double wrap_dist(double x1, double x2, double width) {
    double dist = x1 - x2;
    if (dist > width / 2) return dist - width;
    if (dist < width / -2) return dist + width;
    return dist;
}

static int in_range(double x, double r, double low, double width, double tw) {
    return (x + r > low && x - r < low + width) ||
           (wrap_dist(x, low, tw) > -r && wrap_dist(x, low + width, tw) < r);
}

int orc_halo_sheet_intersect(const orc_halo *h, const float origin[3],
                             const float width[3], double tw) {
    return in_range(h->origin[0], h->rMax, (double)origin[0], (double)width[0], tw) &&
           in_range(h->origin[1], h->rMax, (double)origin[1], (double)width[1], tw) &&
           in_range(h->origin[2], h->rMax, (double)origin[2], (double)width[2], tw);
}

/* los/halo.go:471-475 */
void orc_halo_plane_to_volume(const orc_halo *h, int ring, double px, double py,
                              double out[3]) {
    float v[3] = {(float)px, (float)py, 0};
    orc_rotate_vec(v, &h->irots[9 * ring]);
    out[0] = (double)v[0]; out[1] = (double)v[1]; out[2] = (double)v[2];
}

/* accessors used by orc_analyze.c / orc_shell.c */
int orc_halo_rings(const orc_halo *h) { return h->rings; }
int orc_halo_bins(const orc_halo *h) { return h->bins; }
int orc_halo_n(const orc_halo *h) { return h->n; }
double orc_halo_rmax(const orc_halo *h) { return h->rMax; }
double orc_halo_phi(const orc_halo *h, int i) { return h->ringPhis[i]; }
int64_t orc_halo_oob(const orc_halo *h) {
    int64_t s = 0;
    for (int r = 0; r < h->rings; r++) s += h->profs[r].oob;
    return s;
}

/* Halo.Split, los/halo.go:107-123 + ProfileRing.Split, profile_ring.go:57-70.
 * First use of a worker (shape mismatch) is a full Init from h. */
orc_halo *orc_halo_split_new(const orc_halo *h) {
    return orc_halo_new(h->norms, h->rings, h->origin, h->rMin, h->rMax,
                        h->bins, h->n, h->defaultRho);
}

/* Re-use of an equally shaped worker copies norms, rots, origin, rMin, rMax
 * and each ring's lowR/highR -- but NOT irots, defaultRho or the ring's dr
 * (`p2.dr = p2.dr`, profile_ring.go:64): a worker keeps the dr of the first
 * halo it ever served.  Its grid is zeroed. */
void orc_halo_split_reuse(orc_halo *w, const orc_halo *h) {
    memcpy(w->norms, h->norms, sizeof(float) * 3 * (size_t)h->rings);
    memcpy(w->rots, h->rots, sizeof(float) * 9 * (size_t)h->rings);
    memcpy(w->origin, h->origin, sizeof w->origin);
    w->rMin = h->rMin; w->rMax = h->rMax;
    for (int r = 0; r < h->rings; r++) {
        w->profs[r].lowR = h->profs[r].lowR;
        w->profs[r].highR = h->profs[r].highR;
    }
    orc_halo_clear(w);
}
