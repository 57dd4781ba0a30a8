/* shellfish_oracle.h -- CPU restatement of the `shellfish shell` line-of-sight
 * hot path of phil-mansfield/shellfish v1.0.4.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load it.  The shipped library (shellfish_b200/libsfk.so) never links,
 * loads or calls anything in oracle/.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Arithmetic contract: IEEE binary32/binary64,
 * round-to-nearest, evaluated operation by operation with NO fused
 * multiply-add (Go on amd64/GOAMD64=v1 does not contract); build with
 * -ffp-contract=off and without -ffast-math.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   pinned by the reference's own golden vectors (tests/test_oracle_golden.py):
 *     ProfileRing Insert/Retrieve (los/profile_ring_test.go:34-93),
 *     insertToRing (los/halo_test.go:28-43), halfAngularWidth /
 *     twoValIntrDist / oneValIntrDist / idxRange / sphereIntersectRing
 *     (los/geom_test.go), Savitzky-Golay taps
 *     (math/interpolate/filter_test.go:40-51), Euler rotations
 *     (los/geom/rotation_test.go:25-79) -- math/mat's tests are benchmarks only, the LU
 *     restatement is cross-checked against numpy instead;
 *     statistically, by the one documented `shell | stats` row (doc/quickstart.md):
 *     Shell.Volume / SurfaceArea / Axes and the R_sp formula (orc_shellprops.c).
 *   PARITY UNPINNED (no reference test holds a known answer and the Go
 *   toolchain is absent, so the reference cannot be run here): Smooth output,
 *   SplashbackRadius, the KDE filter, the Penna fit (gonum mat64 is an
 *   un-vendored dependency whose published LAPACK algorithms are restated),
 *   and every end-to-end `shell` result.  Those are cross-checked against
 *   independent scipy/numpy formulations instead (tests/test_oracle_xcheck.py).
 */
#ifndef SHELLFISH_ORACLE_H
#define SHELLFISH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- math/rand/xorshift.go:13-33, math/rand/rand.go:91-96, cmd/shell.go:566-592 */
void orc_norm_vecs(uint64_t seed, int rings, float *out /* [rings][3] */);

/* ---- los/geom/rotation.go:26-81, math/mat/mat32.go:53-74 */
void orc_spherical_angles(float x, float y, float z, float *phi, float *theta);
void orc_euler_matrix(float phi, float theta, float psi, float out[9]);
void orc_euler_matrix_between(const float v1[3], const float v2[3], float out[9]);
void orc_rotate_vec(float v[3], const float m[9]);

/* ---- Go stdlib math.Pow (src/math/pow.go), see orc_math.c */
double orc_go_pow(double x, double y);

/* ---- cosmo/param.go:12-41, cosmo/const.go:3-8 */
double orc_rho_average(double H0, double omegaM, double omegaL, double z);

/* ---- math/mat/mat.go:224-348 (Crout LU with implicit scaling) */
/* Solves a x = b in place for square n x n row-major a (a is overwritten with
 * the LU factors).  Returns 0, or -1 on an all-zero row ("Singular Matrix."). */
int orc_lu_solve(double *a, int n, double *b);
int orc_lu_invert(double *a, int n, double *out);

/* ---- math/interpolate/filter.go:218-318 */
/* ld = 0: smoothing taps; ld = 1: first-derivative taps scaled by 1/dx. */
int orc_savgol_kernel(int order, int width, int ld, double dx, double *taps);
/* filter.go:104-161 with the Extension boundary rule. */
void orc_convolve_extension(const double *taps, int width, const double *xs,
                            int n, double *out);

/* ---- math/interpolate/spline.go */
typedef struct orc_spline orc_spline;
orc_spline *orc_spline_new(const double *xs, const double *ys, int n);
void orc_spline_free(orc_spline *sp);
/* Returns 0 and writes *y, or -1 where Go would panic (out of bounds). */
int orc_spline_eval(const orc_spline *sp, double x, double *y);

/* ---- los/profile_ring.go */
typedef struct orc_ring orc_ring;
/* Process-wide switch, read when a ring is created: 0 = the reference's float64 bins (default),
 * 1 = the "exact bins" variant documented in orc_los.c (test infrastructure for tie analysis). */
void orc_set_exact_bins(int on);
orc_ring *orc_ring_new(double lowR, double highR, int bins, int n);
void orc_ring_free(orc_ring *p);
void orc_ring_insert(orc_ring *p, double start, double end, double rho, int i);
void orc_ring_retrieve(const orc_ring *p, int i, double *out);
const double *orc_ring_derivs(const orc_ring *p);

/* ---- los/halo.go */
typedef struct orc_halo orc_halo;
orc_halo *orc_halo_new(const float *norms, int rings, const double origin[3],
                       double rMin, double rMax, int bins, int n,
                       double defaultRho);
void orc_halo_free(orc_halo *h);
/* Enables the integer parity taps (insert counts and edge counts). */
void orc_halo_enable_taps(orc_halo *h);
void orc_halo_transform(const orc_halo *h, float *xs /* [n][3] */, int64_t n,
                        double totalWidth);
void orc_halo_intersect(const orc_halo *h, const float *xs, int64_t n,
                        double rad, uint8_t *intr);
void orc_halo_insert(orc_halo *h, const float vec[3], double radius, double rho);
void orc_halo_insert_to_ring(orc_halo *h, const float vec[3], double radius,
                             double rho, int ring);
int orc_halo_sphere_intersect_ring(const orc_halo *h, const float vec[3],
                                   double radius, int ring);
void orc_halo_idx_range(const orc_halo *h, double phiLo, double phiHi,
                        int out[4]);
void orc_halo_get_rhos(const orc_halo *h, int ring, int los, double *buf);
void orc_halo_get_rs(const orc_halo *h, double *buf);
int orc_halo_sheet_intersect(const orc_halo *h, const float origin[3],
                             const float width[3], double totalWidth);
void orc_halo_plane_to_volume(const orc_halo *h, int ring, double px,
                              double py, double out[3]);
/* Adds the grids of `src` into `dst` (Halo.Join, los/halo.go:129-140). */
void orc_halo_join(orc_halo *dst, const orc_halo *src);
/* Zeroes the grid (ProfileRing.Split on an already-sized ring, :57-70). */
void orc_halo_clear(orc_halo *h);
const double *orc_halo_derivs(const orc_halo *h, int ring);
const float *orc_halo_rot(const orc_halo *h, int ring);
const float *orc_halo_irot(const orc_halo *h, int ring);
/* taps: n_insert [rings][n]; start/end edge counts [rings][n][bins] */
const uint32_t *orc_halo_tap_inserts(const orc_halo *h);
const uint32_t *orc_halo_tap_start(const orc_halo *h);
const uint32_t *orc_halo_tap_end(const orc_halo *h);

double orc_half_angular_width(double dist2, double r2);
void orc_two_val_intr_dist(double dist2, double rad2, double b, double *lo,
                           double *hi);
double orc_one_val_intr_dist(double dist2, double rad2, double b, double dir);

/* ---- los/analyze/smooth.go:46-96, splashback_radius.go:30-63 */
/* valTaps/derivTaps are the cached kernels (smooth.go:84-96). ys is restored
 * on return like the reference does (log then exp round trip, :68-75). */
int orc_smooth(const double *valTaps, const double *derivTaps, int window,
               double *ys, int n, double *vals, double *derivs);
int orc_splashback_radius(const double *rs, const double *rhos,
                          const double *derivs, int n, double dLim, double *r);

/* ---- los/analyze/kde.go + penna.go:115-165 for ONE ring.
 * rs/phis: valid points in LoS order.  Outputs filtered (r, phi) and indices.
 * Returns number kept (>=0) or a negative ORC_E_* code where Go panics. */
#define ORC_E_NO_POINTS (-2)   /* NewKDETree panic, kde.go:72-81 */
#define ORC_E_NO_MAXIMUM (-3)  /* maxesTree[0][0][0] index panic, kde.go:212 */
#define ORC_E_SPLINE (-4)      /* Spline panic (unsorted table / out of range) */
int orc_filter_ring(const double *rs, const double *phis, int n, int levels,
                    double h0, double *fRs, double *fPhis, int *fIdx,
                    double *hUsed);

/* ---- los/analyze/penna.go:14-58,170-194 (gonum Dgetrf/Dgetri restated) */
int orc_penna_coeffs(const double *xs, const double *ys, const double *zs,
                     int64_t N, int I, int J, int K, double *cs);
double orc_penna_func(const double *cs, int I, int J, int K, double phi,
                      double th);

/* ---- cmd/shell.go:288-627: the whole `shell` loop on in-memory blocks. */
typedef struct {
    int64_t radialBins, spokes, rings;
    double rMaxMult, rMinMult, rKernelMult;
    double eta;
    int64_t order, smoothingWindow, levels, subsampleFactor;
    double losSlopeCutoff, backgroundRhoMult;
    uint64_t seed;    /* cmd/cmd.go:658 randSeed, pinned by the caller */
    int64_t threads;  /* GlobalConfig.Threads (<=0 -> 1 here) */
} orc_shell_config;

typedef struct {
    /* io.Header, io/io.go:71-83 */
    double Z, OmegaM, OmegaL, H100;
    int64_t N;
    double TotalWidth;
    float Origin[3], Width[3];
    /* block contents as buf.Read returns them (io/io.go:58-67) */
    const float *xs; /* [N][3] */
    const float *ms; /* [N] */
} orc_block;

typedef struct {
    /* all optional (may be NULL) */
    double *rsp;       /* [halo][rings][spokes] splashback radius per LoS, NaN if !ok */
    double *profiles;  /* [halo][rings][spokes][bins] GetRhos output */
    uint32_t *nInsert; /* [halo][rings][spokes] */
    uint32_t *startEdges; /* [halo][rings][spokes][bins] */
    uint32_t *endEdges;   /* [halo][rings][spokes][bins] */
    int64_t *nFiltered;   /* [halo] points that reached the Penna fit */
    int64_t *nCandidates; /* [halo] particles passing Halo.Intersect */
    int64_t *nIntersections; /* [halo] ProfileRing.Insert calls past the early-out */
} orc_taps;

/* Runs one snapshot: halos (x,y,z,r200m)[nHalo], blocks[nBlock].
 * coeffs: [nHalo][2*order*order]; ok: [nHalo] (0 = failed/skipped).
 * status: [nHalo] 0 or ORC_E_*.  Returns 0 or a negative error. */
int orc_shell_run(const orc_shell_config *cfg, float minMass, int64_t nHalo,
                  const double *hx, const double *hy, const double *hz,
                  const double *hr, int64_t nBlock, const orc_block *blocks,
                  double *coeffs, uint8_t *ok, int32_t *status,
                  const orc_taps *taps);

/* ---- PercentileProfile mode, cmd/shell.go:629-660 + math/sort/median.go */
double orc_nth_largest(const double *xs, int64_t len, int64_t n, double *buf);
double orc_percentile(const double *xs, int64_t len, double p, double *buf);
void orc_calc_percentile(double rMin, double rMax, int bins, const double *profiles,
                         int64_t nLos, double percentile, double *out);

/* ---- `stats` M_sp pass, cmd/stats.go:429-543,685-692; los/analyze/shell.go:349 */
int orc_shell_contains(const double *cs, int order, double x, double y, double z);
double orc_mass_contained(double totalWidth, const float *xs, const float *ms, int64_t n,
                          const double *cs, int order, const float c[3],
                          double rLow, double rHigh, int workers);
int64_t orc_shell_particles(double totalWidth, const float *xs, int64_t n, const double *cs, int order,
                            const float c[3], float sphereR, double rLow, double rHigh, double shellWidth,
                            int64_t *out);
int orc_sphere_sheet_intersect(const float c[3], float r, const float origin[3],
                               const float width[3], double tw);

/* ---- `stats` shell property integrals, los/analyze/shell.go:21-32,58-283 (orc_shellprops.c) */
typedef double (*orc_uniform_fn)(void *state);    /* one rand.Float64() */
typedef struct { uint32_t w, x, y, z; } orc_stream;   /* math/rand/xorshift.go */
void orc_stream_init(orc_stream *s, uint64_t seed);
double orc_stream_next(void *state);
typedef struct {
    int n;                                  /* knots per axis */
    const double *acs, *bcs;                /* grid.ACRatios, grid.BCRatios */
    const double *acGrid, *bcGrid, *cGrid;  /* [n*n], vals[n*yi + xi] */
} orc_axis_tables;
double orc_shell_volume(const double *cs, int order, int samples, orc_uniform_fn fn, void *st);
double orc_cos_norm(const double *cs, int order, double phi, double theta);
double orc_shell_surface_area(const double *cs, int order, int samples, orc_uniform_fn fn, void *st);
void orc_shell_radial_range(const double *cs, int order, int samples, orc_uniform_fn fn, void *st,
                            double *low, double *high);
int orc_bicubic_eval(const double *xs, int nx, const double *ys, int ny, const double *vals,
                     double x, double y, double *out);
void orc_angular_fraction_profile(const double *cs, int order, int samples, int bins, double rMin, double rMax,
                                  orc_uniform_fn fn, void *st, double *rs, double *fs);
int orc_shell_axes(const double *cs, int order, int samples, orc_uniform_fn fn, void *st,
                   const orc_axis_tables *tables, double *a, double *b, double *c, double aVec[3]);

/* Resets the process-wide Savitzky-Golay kernel cache (smooth.go:9-12). */
void orc_reset_kernel_cache(void);

#ifdef __cplusplus
}
#endif
#endif
