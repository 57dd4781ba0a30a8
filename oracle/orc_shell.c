/* orc_shell.c -- oracle (TEST INFRASTRUCTURE): the `shellfish shell` loop of
 * cmd/shell.go:288-627 run on in-memory particle blocks, including the
 * strided-worker / Split / Join structure (cmd/shell.go:423-500) on pthreads.
 * PARITY UNPINNED end to end (no reference golden catalog exists).
 * Build: gcc -std=c11 -O2 -ffp-contract=off -pthread (no -ffast-math). */
#include "shellfish_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* Go constant 4*math.Pi/3, folded exactly before rounding (cmd/shell.go:485,549) */
#define GO_4PI_3 4.18879020478639098461685784437267051226289253250014

int orc_halo_rings(const orc_halo *h);
int orc_halo_bins(const orc_halo *h);
int orc_halo_n(const orc_halo *h);
double orc_halo_rmax(const orc_halo *h);
double orc_halo_phi(const orc_halo *h, int i);
orc_halo *orc_halo_split_new(const orc_halo *h);
void orc_halo_split_reuse(orc_halo *w, const orc_halo *h);

/* los/analyze/smooth.go:9-12,84-96: kernels cached by window only, so the
 * derivative kernel keeps the dx of the first profile ever smoothed. */
#define MAX_CACHE 8
static struct { int window; double *k, *kd; } g_cache[MAX_CACHE];
static int g_ncache = 0;
static pthread_mutex_t g_cacheMu = PTHREAD_MUTEX_INITIALIZER;

void orc_reset_kernel_cache(void) {
    pthread_mutex_lock(&g_cacheMu);
    for (int i = 0; i < g_ncache; i++) { free(g_cache[i].k); free(g_cache[i].kd); }
    g_ncache = 0;
    pthread_mutex_unlock(&g_cacheMu);
}

static int get_smoothing_kernel(int window, double dx, const double **k, const double **kd) {
    pthread_mutex_lock(&g_cacheMu);
    for (int i = 0; i < g_ncache; i++)
        if (g_cache[i].window == window) {
            *k = g_cache[i].k; *kd = g_cache[i].kd;
            pthread_mutex_unlock(&g_cacheMu);
            return 0;
        }
    if (g_ncache == MAX_CACHE) { pthread_mutex_unlock(&g_cacheMu); return -1; }
    double *a = (double *)malloc(sizeof(double) * (size_t)window);
    double *b = (double *)malloc(sizeof(double) * (size_t)window);
    if (orc_savgol_kernel(4, window, 0, 0, a) || orc_savgol_kernel(4, window, 1, dx, b)) {
        free(a); free(b);
        pthread_mutex_unlock(&g_cacheMu);
        return -1;
    }
    g_cache[g_ncache].window = window;
    g_cache[g_ncache].k = a; g_cache[g_ncache].kd = b;
    g_ncache++;
    *k = a; *kd = b;
    pthread_mutex_unlock(&g_cacheMu);
    return 0;
}

/* ------------------------------------------------- worker fan-out (Insert) */

typedef struct {
    orc_halo *h;
    const float *xs, *ms;
    const uint8_t *intr;
    int64_t n;
    int offset, workers;
    int64_t sf3;
    double rad, sphVol, rhoM;
} worker_arg;

/* cmd/shell.go:479-500 chanLoadSphereVec */
static void *worker_main(void *p) {
    worker_arg *w = (worker_arg *)p;
    int64_t skip = (int64_t)w->workers * w->sf3;
    for (int64_t i = (int64_t)w->offset * w->sf3; i < w->n; i += skip) {
        if (w->intr[i]) {
            double rho = ((double)w->ms[i] * (double)w->sf3 / w->sphVol) / w->rhoM;
            orc_halo_insert(w->h, &w->xs[3 * i], w->rad, rho);
        }
    }
    return NULL;
}

/* ---------------------------------------------------- analysis of one halo */

/* cmd/shell.go:613-627 calcCoeffs = ring_buffer.go:57-93 per ring, then
 * penna.go:115-165 FilterPoints and :88-108 PennaVolumeFit. */
static int calc_coeffs(orc_halo *h, const orc_shell_config *c, double *cs,
                       double *rspOut, double *profOut, int64_t *nFiltered) {
    int rings = orc_halo_rings(h), n = orc_halo_n(h), bins = orc_halo_bins(h);
    int window = (int)c->smoothingWindow;
    int order = (int)c->order;
    double *profRs = (double *)malloc(sizeof(double) * (size_t)bins * 4);
    double *profRhos = profRs + bins, *smoothRhos = profRhos + bins,
           *smoothDerivs = smoothRhos + bins;
    size_t cap = (size_t)rings * (size_t)n;
    /* the RingBuffers of loop(), cmd/shell.go:294-297: Rs/Phis/Oks per ring */
    double *Rs = (double *)malloc(sizeof(double) * cap * 2);
    double *Phis = Rs + cap;
    uint8_t *oks = (uint8_t *)malloc(cap);
    double *vRs = (double *)malloc(sizeof(double) * (size_t)n * 4);
    double *vPhis = vRs + n, *fRs = vPhis + n, *fPhis = fRs + n;
    double *fX = (double *)malloc(sizeof(double) * cap * 3);
    double *fY = fX + cap, *fZ = fY + cap;
    int64_t nPts = 0;
    int rc = 0;

    orc_halo_get_rs(h, profRs);
    /* calcCoeffs first runs Clear + Splashback on every ring (:614-617) ... */
    for (int ring = 0; ring < rings && rc == 0; ring++) {
        for (int i = 0; i < n; i++) {
            size_t at = (size_t)ring * n + i;
            oks[at] = 0; Rs[at] = 0; Phis[at] = 0;
            orc_halo_get_rhos(h, ring, i, profRhos);
            if (profOut) memcpy(profOut + at * bins, profRhos, sizeof(double) * (size_t)bins);
            if (rspOut) rspOut[at] = NAN;
            if (bins <= window) continue; /* Smooth returns ok=false */
            double dx = log(profRs[1]) - log(profRs[0]);
            const double *k, *kd;
            if (get_smoothing_kernel(window, dx, &k, &kd)) { rc = -1; break; }
            orc_smooth(k, kd, window, profRhos, bins, smoothRhos, smoothDerivs);
            double r;
            if (!orc_splashback_radius(profRs, smoothRhos, smoothDerivs, bins,
                                       c->losSlopeCutoff, &r))
                continue;
            oks[at] = 1;
            Rs[at] = r;
            Phis[at] = orc_halo_phi(h, i);
            if (Phis[at] < 0) Phis[at] += 3.14159265358979323846264338327950288;
            if (rspOut) rspOut[at] = r;
        }
    }
    /* ... then FilterPoints walks the rings in order (penna.go:119-162) */
    for (int ring = 0; ring < rings && rc == 0; ring++) {
        int nv = 0;
        for (int i = 0; i < n; i++) {
            size_t at = (size_t)ring * n + i;
            if (oks[at]) { vRs[nv] = Rs[at]; vPhis[nv] = Phis[at]; nv++; }
        }
        int kept = orc_filter_ring(vRs, vPhis, nv, (int)c->levels,
                                   orc_halo_rmax(h) / c->eta, fRs, fPhis, NULL, NULL);
        if (kept < 0) { rc = kept; break; }
        /* fXs = fRs*cos, fYs = fRs*sin; PlaneToVolume (via float32) */
        for (int i = 0; i < kept; i++) {
            double px = fRs[i] * cos(fPhis[i]), py = fRs[i] * sin(fPhis[i]);
            double v[3];
            orc_halo_plane_to_volume(h, ring, px, py, v);
            fX[nPts] = v[0]; fY[nPts] = v[1]; fZ[nPts] = v[2];
            nPts++;
        }
    }
    if (nFiltered) *nFiltered = nPts;
    if (rc == 0) {
        if (nPts == 0) rc = ORC_E_NO_POINTS;
        else if (orc_penna_coeffs(fX, fY, fZ, nPts, order, order, 2, cs)) rc = -1;
    }
    free(fX); free(vRs); free(oks); free(Rs); free(profRs);
    return rc;
}

/* ---------------------------------------------------------------- the loop */

int orc_shell_run(const orc_shell_config *c, float minMass, int64_t nHalo,
                  const double *hx, const double *hy, const double *hz,
                  const double *hr, int64_t nBlock, const orc_block *blocks,
                  double *coeffs, uint8_t *ok, int32_t *status,
                  const orc_taps *taps) {
    if (nBlock <= 0 || nHalo <= 0) return -1;
    int rings = (int)c->rings, n = (int)c->spokes, bins = (int)c->radialBins;
    int P2 = (int)(c->order * c->order * 2);
    int workers = c->threads > 0 ? (int)c->threads : 1;
    int64_t sf = c->subsampleFactor, sf3 = sf * sf * sf;
    size_t rn = (size_t)rings * (size_t)n, rnb = rn * (size_t)bins;
    int wantTaps = taps && (taps->nInsert || taps->startEdges || taps->endEdges ||
                            taps->nIntersections);

    /* createHalos, cmd/shell.go:530-564: cosmology of the first block */
    const orc_block *hd0 = &blocks[0];
    float *norms = (float *)malloc(sizeof(float) * 3 * (size_t)rings);
    orc_halo **halos = (orc_halo **)calloc((size_t)nHalo, sizeof(orc_halo *));
    for (int64_t i = 0; i < nHalo; i++) {
        ok[i] = 0; status[i] = 0;
        for (int k = 0; k < P2; k++) coeffs[i * P2 + k] = 0;
        double r = hr[i];
        if (r <= 0) continue; /* reference leaves a nil *Halo here */
        orc_norm_vecs(c->seed, rings, norms);
        double origin[3] = {hx[i], hy[i], hz[i]};
        double rMax = r * c->rMaxMult, rMin = r * c->rMinMult;
        double rad = r * c->rKernelMult;
        double sphVol = GO_4PI_3 * rad * rad * rad;
        double rhoM = orc_rho_average(hd0->H100 * 100, hd0->OmegaM, hd0->OmegaL, hd0->Z);
        double rho = (((double)minMass * (double)sf3) / sphVol) / rhoM;
        double defaultRho = rho * c->backgroundRhoMult;
        halos[i] = orc_halo_new(norms, rings, origin, rMin, rMax, bins, n, defaultRho);
        if (wantTaps) orc_halo_enable_taps(halos[i]);
        if (taps && taps->nCandidates) taps->nCandidates[i] = 0;
    }
    free(norms);

    /* worker halos (sphBuf.sphWorkers), cmd/shell.go:315-320 */
    orc_halo **wh = (orc_halo **)calloc((size_t)workers, sizeof(orc_halo *));
    pthread_t *tids = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)workers);
    worker_arg *wargs = (worker_arg *)malloc(sizeof(worker_arg) * (size_t)workers);

    /* sphereLoop, cmd/shell.go:376-415 */
    for (int64_t b = 0; b < nBlock; b++) {
        const orc_block *hd = &blocks[b];
        int any = 0;
        for (int64_t i = 0; i < nHalo && !any; i++)
            if (halos[i] && orc_halo_sheet_intersect(halos[i], hd->Origin, hd->Width, hd->TotalWidth))
                any = 1;
        if (!any) continue;
        /* buf.Read gives a fresh slice that Transform then mutates */
        float *xs = (float *)malloc(sizeof(float) * 3 * (size_t)(hd->N > 0 ? hd->N : 1));
        memcpy(xs, hd->xs, sizeof(float) * 3 * (size_t)hd->N);
        uint8_t *intr = (uint8_t *)malloc((size_t)(hd->N > 0 ? hd->N : 1));
        for (int64_t i = 0; i < nHalo; i++) {
            orc_halo *h = halos[i];
            if (!h || !orc_halo_sheet_intersect(h, hd->Origin, hd->Width, hd->TotalWidth))
                continue;
            /* loadSphereVecs, cmd/shell.go:423-465 */
            orc_halo_transform(h, xs, hd->N, hd->TotalWidth);
            double rad = orc_halo_rmax(h) * c->rKernelMult / c->rMaxMult;
            orc_halo_intersect(h, xs, hd->N, rad, intr);
            if (taps && taps->nCandidates)
                for (int64_t k = 0; k < hd->N; k += sf3) taps->nCandidates[i] += intr[k];
            double sphVol = GO_4PI_3 * rad * rad * rad;
            double rhoM = orc_rho_average(hd->H100 * 100, hd->OmegaM, hd->OmegaL, hd->Z);
            /* h.Split(sphWorkers), los/halo.go:107-123 */
            for (int w = 0; w < workers - 1; w++) {
                if (!wh[w]) {
                    wh[w] = orc_halo_split_new(h);
                    if (wantTaps) orc_halo_enable_taps(wh[w]);
                } else {
                    orc_halo_split_reuse(wh[w], h);
                }
            }
            for (int w = 0; w < workers; w++) {
                wargs[w].h = (w < workers - 1) ? wh[w] : h;
                wargs[w].xs = xs; wargs[w].ms = hd->ms; wargs[w].intr = intr;
                wargs[w].n = hd->N; wargs[w].offset = w; wargs[w].workers = workers;
                wargs[w].sf3 = sf3; wargs[w].rad = rad; wargs[w].sphVol = sphVol;
                wargs[w].rhoM = rhoM;
            }
            for (int w = 0; w < workers - 1; w++)
                pthread_create(&tids[w], NULL, worker_main, &wargs[w]);
            worker_main(&wargs[workers - 1]);
            for (int w = 0; w < workers - 1; w++) pthread_join(tids[w], NULL);
            for (int w = 0; w < workers - 1; w++) orc_halo_join(h, wh[w]);
        }
        free(intr);
        free(xs);
    }

    /* haloAnalysis, cmd/shell.go:502-528 */
    for (int64_t i = 0; i < nHalo; i++) {
        orc_halo *h = halos[i];
        if (!h) { status[i] = ORC_E_NO_POINTS; continue; }
        if (taps) {
            if (taps->nInsert)
                memcpy(taps->nInsert + (size_t)i * rn, orc_halo_tap_inserts(h), rn * sizeof(uint32_t));
            if (taps->startEdges)
                memcpy(taps->startEdges + (size_t)i * rnb, orc_halo_tap_start(h), rnb * sizeof(uint32_t));
            if (taps->endEdges)
                memcpy(taps->endEdges + (size_t)i * rnb, orc_halo_tap_end(h), rnb * sizeof(uint32_t));
            if (taps->nIntersections) {
                const uint32_t *t = orc_halo_tap_inserts(h);
                int64_t s = 0;
                for (size_t k = 0; k < rn; k++) s += t[k];
                taps->nIntersections[i] = s;
            }
        }
        int rc = calc_coeffs(h, c, coeffs + i * P2,
                             taps && taps->rsp ? taps->rsp + (size_t)i * rn : NULL,
                             taps && taps->profiles ? taps->profiles + (size_t)i * rnb : NULL,
                             taps && taps->nFiltered ? &taps->nFiltered[i] : NULL);
        status[i] = rc;
        ok[i] = rc == 0;
    }

    for (int w = 0; w < workers; w++) orc_halo_free(wh[w]);
    for (int64_t i = 0; i < nHalo; i++) orc_halo_free(halos[i]);
    free(wargs); free(tids); free(wh); free(halos);
    return 0;
}
