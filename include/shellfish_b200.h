/* shellfish_b200.h -- C ABI of the B200 (sm_100a) line-of-sight shell finder.
 *
 * Drop-in boundary for the hot path of `shellfish shell`
 * (phil-mansfield/shellfish v1.0.4).  The reference has no FFI layer; its seam
 * is the set of los.Halo / analyze.* calls made from cmd/shell.go.  Each entry
 * point below names the reference code it replaces; INTEGRATION.md shows the
 * cgo stub that binds it from cmd/shell.go.
 *
 * Conventions: plain pointers and sizes, no CUDA or torch types.  Every call
 * returns 0 on success or a negative SFK_E_* code; sfk_last_error() gives the
 * message the host prints before "Shellfish terminating." (shellfish.go:441).
 * Host input buffers are copied before the call returns (cgo may not retain
 * Go pointers; io.VectorBuffer reuses its slices, io/io.go:57-67).  One
 * context drives one GPU and is not re-entrant (cmd/shell.go's loop is a
 * single goroutine).  There is no CPU fallback: sfk_create fails when no
 * sm_100 device is present.
 */
#ifndef SHELLFISH_B200_H
#define SHELLFISH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFK_OK 0
#define SFK_E_INVALID (-1)   /* bad argument / call order */
#define SFK_E_CUDA (-2)      /* CUDA runtime error */
#define SFK_E_NODEVICE (-3)  /* no usable sm_100 GPU */
#define SFK_E_NOMEM (-4)
#define SFK_E_UNSUPPORTED (-5)

/* per-halo status written by sfk_finish_snapshot */
#define SFK_HALO_OK 0
#define SFK_HALO_SKIPPED 1     /* r200m <= 0: cmd/shell.go:535-538 leaves a nil halo */
#define SFK_HALO_NO_POINTS 2   /* a ring had no valid LoS: NewKDETree panic, kde.go:72-81 */
#define SFK_HALO_NO_MAXIMUM 3  /* level-0 KDE without maximum: index panic, kde.go:212 */
#define SFK_HALO_SPLINE 4      /* Spline table/bounds panic, spline.go:56-96 */
#define SFK_HALO_SINGULAR 5    /* singular normal equations in the Penna fit */
#define SFK_HALO_NOT_OWNED 6   /* left to another rank by sfk_set_owned */

/* sfk_params.flags */
#define SFK_FLAG_TAPS 1u /* keep parity taps: profiles, integer edge counts */
/* A/B and tests: evaluate the two Savitzky-Golay filters of Smooth (smooth.go:46-81) with
 * the literal 2 x SmoothingWindow tap loop of ConvolveAt (filter.go:104-161) and libm's log.
 * The default evaluates them through five sliding polynomial moments (same values to
 * ~1e-13, same R_sp) whenever RadialBins = 256, 29 <= SmoothingWindow <= 224 and the background
 * density is positive; every other configuration always takes the tap loop. */
#define SFK_FLAG_LOS_TAPLOOP 2u
/* Tests only: build the ring-frame hit records exactly as insertToRing / idxRange do
 * (los/halo.go:241-310: all n LoS for a centre-containing circle, float64 asin / atan2,
 * nothing dropped), i.e. without the exact pruning of the product path, so that a test
 * can show pruned and unpruned integer bins are identical. */
#define SFK_FLAG_NO_PRUNE 4u
/* Tests only: keep every halo's int64 fixed-point bin grid for sfk_get_grid. */
#define SFK_FLAG_KEEP_GRID 8u

typedef struct sfk_ctx sfk_ctx;

/* ShellConfig, cmd/shell.go:24-35; defaults :105-119; validation :143-187.
 * seed replaces cmd/cmd.go:658 randSeed (wall clock in the reference). */
typedef struct {
    int64_t radial_bins, spokes, rings;
    double r_max_mult, r_min_mult, r_kernel_mult;
    double eta;
    int64_t order, smoothing_window, levels, subsample_factor;
    double los_slope_cutoff, background_rho_mult;
    int32_t percentile_profile; /* cmd/shell.go:28: output calcPercentile rows instead of coefficients */
    double percentile;          /* in [0, 100], cmd/shell.go:119,656 */
    uint64_t seed;
    int32_t device; /* CUDA ordinal */
    uint32_t flags; /* SFK_FLAG_* */
} sfk_params;

/* io.CosmologyHeader, io/io.go:71-76 */
typedef struct {
    double z, omega_m, omega_l, h100;
} sfk_cosmo;

/* Work counters of the last finished snapshot. */
typedef struct {
    int64_t n_halos;
    int64_t n_candidates;    /* particles passing Halo.Intersect, los/halo.go:147 */
    int64_t n_ring_hits;     /* sphereIntersectRing successes, los/halo.go:218 */
    int64_t n_intersections; /* ProfileRing.Insert calls past the early-out, profile_ring.go:92-95 */
    int64_t n_points;        /* filtered points that reached the Penna fit */
    double ms_cull, ms_hits, ms_bin, ms_los, ms_filter, ms_fit; /* CUDA-event time per stage */
    int64_t bin_launches;    /* launches of the binning kernel */
    int64_t launches;        /* kernel launches of this library since begin_snapshot */
    int64_t n_los_literal;   /* lines of sight with long plateaus, smoothed by the literal tap loop
                                instead of sliding moments (exact ties; 0 for well-sampled halos) */
} sfk_stats;

/* Number of CUDA devices visible to the process (0 when there is none). */
int sfk_device_count(void);

/* Fills *p with the defaults of cmd/shell.go:105-119. */
void sfk_default_params(sfk_params *p);

/* ShellConfig.ReadConfig + validate (cmd/shell.go:102-187) and the per-run
 * set-up of loop() (cmd/shell.go:288-320): ring normals (normVecs :566),
 * rotation matrices and LoS vectors (Halo.Init, los/halo.go:65-100). */
int sfk_create(sfk_ctx **out, const sfk_params *params);
void sfk_destroy(sfk_ctx *ctx);
const char *sfk_last_error(const sfk_ctx *ctx);

/* Start of one iteration of the snapshot loop, cmd/shell.go:322-344.  cosmo
 * is the header of the first block of the first snapshot (hds[0], :305,:340)
 * and min_mass is buf.MinMass() (:309). */
int sfk_begin_snapshot(sfk_ctx *ctx, const sfk_cosmo *cosmo, float min_mass);

/* createHalos, cmd/shell.go:530-564.  May be called once per snapshot. */
int sfk_add_halos(sfk_ctx *ctx, int64_t n, const double *x, const double *y,
                  const double *z, const double *r200m);

/* Multi-GPU sharding: owned is [n_halo]; halos with owned[i] == 0 are left to
 * another rank.  They still take part in Halo.Transform's cumulative in-place
 * wrap of each block (los/halo.go:168-197, called per halo at
 * cmd/shell.go:441), so the halos this rank does own see bit-identical
 * positions to a single-process run.  Call after sfk_add_halos, before the
 * first block. */
int sfk_set_owned(sfk_ctx *ctx, int64_t n, const uint8_t *owned);

/* binIntersections / Halo.SheetIntersect (cmd/shell.go:599-611,
 * los/halo.go:442-467) for one block header.  hit may be NULL; *n_hit
 * receives the number of halos touching the block. */
int sfk_block_halo_mask(sfk_ctx *ctx, const float origin[3], const float width[3],
                        double total_width, uint8_t *hit, int64_t *n_hit);

/* Body of sphereLoop for one buf.Read (cmd/shell.go:399-410): for every halo
 * touching the block, in halo order, Transform + Intersect
 * (los/halo.go:168-197,147-164) and queueing of the surviving particles with
 * the weight of chanLoadSphereVec (cmd/shell.go:479-500).  xyz is [n][3],
 * mass is [n] -- or NULL when every particle of the block has the mass given to
 * sfk_begin_snapshot (buf.MinMass(): what the LGadget-2 and gotetra readers fill ms with,
 * io/lgadget2.go:228, io/gotetra.go:43), which saves a quarter of the transfer; both are
 * host pointers, consumed when the call returns (the copy runs on
 * a second stream into one of two staging buffers and overlaps the cull of the previous
 * block, so page-locked host memory pays). */
int sfk_push_block(sfk_ctx *ctx, const float *xyz, const float *mass, int64_t n,
                   const sfk_cosmo *cosmo, double total_width,
                   const float origin[3], const float width[3]);

/* sfk_push_block for a reader that owns several page-locked buffers (a ring of
 * io.VectorBuffer slices allocated once, or a whole snapshot held in memory): the call only
 * QUEUES the copy, so the copies of consecutive blocks run back to back on the bus instead of
 * one per call round trip.  xyz / mass must stay unchanged until sfk_push_sync or
 * sfk_finish_snapshot returns.  Same results as sfk_push_block. */
int sfk_push_block_async(sfk_ctx *ctx, const float *xyz, const float *mass, int64_t n,
                         const sfk_cosmo *cosmo, double total_width,
                         const float origin[3], const float width[3]);
/* Returns once every buffer handed to sfk_push_block_async has been copied (they may be reused). */
int sfk_push_sync(sfk_ctx *ctx);

/* Same with xyz/mass already in device memory of ctx's GPU (used to time the
 * kernels with inputs resident in HBM).  The buffers must be complete on the device when
 * they are handed over (work queued before the first device push of a snapshot is waited
 * for once) and must stay valid and unchanged until sfk_finish_snapshot returns: the cull
 * of a block is queued on the context's stream, not finished, when the call returns. */
int sfk_push_block_device(sfk_ctx *ctx, const float *d_xyz, const float *d_mass,
                          int64_t n, const sfk_cosmo *cosmo, double total_width,
                          const float origin[3], const float width[3]);

/* Halo.Insert for all queued particles (los/halo.go:201-277,
 * profile_ring.go:92-120) followed by haloAnalysis (cmd/shell.go:502-528):
 * RingBuffer.Splashback, FilterPoints, PennaVolumeFit.  coeffs is
 * [n_halo][sfk_row_length()] in catalog column order, status is [n_halo]
 * SFK_HALO_*; either may be NULL.  With percentile_profile set, a row is
 * calcPercentile's output instead (cmd/shell.go:629-660): radial_bins bin
 * radii (Halo.GetRs) followed by radial_bins densities, each the
 * percentile-th value over all rings*spokes profiles (msort.Percentile). */
int sfk_finish_snapshot(sfk_ctx *ctx, double *coeffs, int32_t *status);

/* Values per halo in the result table: 2*order*order (cmd/shell.go:223), or
 * 2*radial_bins in percentile mode (cmd/shell.go:632). */
int64_t sfk_row_length(const sfk_ctx *ctx);

/* ---- split-halo mode ------------------------------------------------------
 * One giant halo whose particles are spread over several GPUs (each rank
 * pushes its own part of the blocks).  The private grids of Halo.Split/Join
 * (los/halo.go:107-140) become one exchange step: the caller all-reduces three
 * device buffers with its own collective library (NCCL through
 * torch.distributed in this repo) between the calls:
 *   sfk_split_count  -> all-reduce SFK_SPLIT_HIT_COUNTS (int32, SUM) and
 *                       SFK_SPLIT_RHO_MAX (int64 bit patterns, MAX)
 *   sfk_split_bin    -> all-reduce SFK_SPLIT_GRID (int64, SUM)
 *   sfk_split_finish -> haloAnalysis on the summed grid (same result on every rank)
 * The grid is int64 fixed point, so the sum is exact and independent of the
 * number of GPUs. */
#define SFK_SPLIT_HIT_COUNTS 0
#define SFK_SPLIT_RHO_MAX 1
#define SFK_SPLIT_GRID 2
int sfk_split_count(sfk_ctx *ctx);
int sfk_split_bin(sfk_ctx *ctx);
int sfk_split_buffer(sfk_ctx *ctx, int32_t which, void **d_ptr, int64_t *count);
int sfk_split_finish(sfk_ctx *ctx, double *coeffs, int32_t *status);

/* The same exchange inside the library, over NCCL (opened with dlopen on first use: a
 * libnccl.so.2 already in the process, $SFK_NCCL_LIB, or the loader path).  One context per
 * GPU, one rank per context:
 *   sfk_comm_unique_id  on one rank; the SFK_COMM_ID_BYTES are handed to the others by the host
 *   sfk_comm_init       on every rank (collective; ncclCommInitRank on the context's GPU)
 *   sfk_split_finish_comm  on every rank after it pushed its share of the blocks (collective):
 *     all-reduce of the per-ring hit counts and of rho_max -> private grids (Halo.Split) ->
 *     reduce-scatter(SUM) of the int64 grid by (halo, ring) row (Halo.Join): rank g receives
 *     rows g*ceil(rows/G) ... -> RingBuffer.Splashback + FilterPoints on those rings only ->
 *     all-gather of R_sp and of the filtered points -> PennaVolumeFit of all points on every
 *     rank, in the single-GPU summation order.  Coefficients, status and taps are identical on
 *     every rank and bitwise equal to a single-GPU run.  sfk_stats.n_intersections is the
 *     halo-wide count. */
#define SFK_COMM_ID_BYTES 128
int sfk_comm_unique_id(void *id128);
int sfk_comm_init(sfk_ctx *ctx, const void *id128, int32_t rank, int32_t nranks);
int sfk_split_finish_comm(sfk_ctx *ctx, double *coeffs, int32_t *status);

/* ---- `shellfish stats`, splashback mass M_sp ---------------------------------
 * massContained (cmd/stats.go:451-543) for every halo of a snapshot: the mass of
 * the particles inside the Penna-Dines surface R(phi, theta) given by a halo's
 * coefficient row (Shell.Contains, los/analyze/shell.go:349-354), with the
 * reference's float32 distance chain and its periodic `wrap` (cmd/stats.go:
 * 429-436, which shifts by half the box).  Call order, mirroring Run()
 * (cmd/stats.go:268-341): begin -> push every block for which
 * sfk_sphere_block_hit is true for at least one halo (binSphereIntersections,
 * :694) -> finish.  Every pushed block is tested against every halo, as
 * cmd/stats.go:322-327 does.
 *   centers  [n_halo][3]  float32(X, Y, Z) of the halo table (boundingSpheres, :407-421)
 *   coeffs   [n_halo][2*order*order]  the P_ijk columns of the `shell` catalog
 *   r_low, r_high  [n_halo]  Shell.RadialRange of each halo (rangeSp, :445-449) */
int sfk_mass_begin(sfk_ctx *ctx, int64_t n_halo, int32_t order, const float *centers,
                   const double *coeffs, const double *r_low, const double *r_high);
int sfk_mass_push_block(sfk_ctx *ctx, const float *xyz, const float *mass, int64_t n,
                        double total_width);
/* Device buffers, as for sfk_push_block_device: complete on the device when handed over (work
 * queued before the first device push of a pass is waited for once), valid until sfk_mass_finish. */
int sfk_mass_push_block_device(sfk_ctx *ctx, const float *d_xyz, const float *d_mass,
                               int64_t n, double total_width);
int sfk_mass_finish(sfk_ctx *ctx, double *masses);
/* ShellParticleFile (cmd/stats.go:326-341, appendShellParticlesChan :545-597): after
 * sfk_mass_begin, sfk_mass_set_band switches the collection on -- shell_width is ShellWidth
 * (!= 0; negative: every particle inside the surface), radii[i] = float32(R200m) of halo i
 * (geom.Sphere.R), r_low / r_high as in sfk_mass_begin.  After every sfk_mass_push_block[_device]
 * sfk_mass_take_band returns the block's selected (halo, particle) pairs as halo << 32 |
 * index-in-block, sorted (the reference's order with one worker); the array belongs to the
 * context and lasts until the next push. */
int sfk_mass_set_band(sfk_ctx *ctx, double shell_width, const float *radii, const double *r_low,
                      const double *r_high);
int sfk_mass_take_band(sfk_ctx *ctx, int64_t *n, const int64_t **pairs);
/* sphereSheetIntersect, cmd/stats.go:685-692 (radius = float32(R200m)); returns 0 or 1. */
int sfk_sphere_block_hit(const float center[3], float radius, const float origin[3],
                         const float width[3], double total_width);

/* ---- `shellfish stats`, shell property integrals ------------------------------
 * Shell.Volume / SurfaceArea / Axes / RadialRange (los/analyze/shell.go:58-283)
 * of the Penna-Dines surface of every halo, as cmd/stats.go:274-289 computes
 * them.  The reference samples MonteCarloSamples random solid angles per
 * quantity from Go's unseeded global math/rand; this call integrates the same
 * quantities over a fixed Gauss-Legendre(cos theta) x uniform(phi) rule of
 * SFK_SHELL_RULE_THETA x SFK_SHELL_RULE_PHI nodes, so it is deterministic and
 * agrees with the reference within the reference's own sampling noise.
 *   coeffs [n_halo][2*order*order]   P_ijk columns of the `shell` catalog
 *   props  [n_halo][SFK_SHELL_PROPS] = R_sp, Volume, SurfaceArea, a (major),
 *          b, c (minor), Ax, Ay, Az, RMin, RMax -- the float columns of the
 *          stats catalog after M_sp (cmd/stats.go:343-349); RMin / RMax are what
 *          rangeSp (:445-449) feeds to massContained as r_low / r_high. */
#define SFK_SHELL_PROPS 11
#define SFK_SHELL_RULE_THETA 128
#define SFK_SHELL_RULE_PHI 256
#define SFK_SHELL_SUMS 13
int sfk_shell_props(sfk_ctx *ctx, int64_t n_halo, int32_t order, const double *coeffs, double *props);
/* Parity tap: the raw per-halo sums of the last sfk_shell_props call,
 * [n_halo][SFK_SHELL_SUMS] = mean r^3, mean area (r^2 / cosNorm), area-weighted
 * means of xx yy zz xy yz zx x y z (not yet divided by the mean area), min r,
 * max r. */
int sfk_shell_sums(sfk_ctx *ctx, int64_t n_halo, double *sums);
/* Shell.AngularFractionProfile (los/analyze/shell.go:295-335), the quantity
 * `shellfish prof` prints for ProfileType = angular-fraction (cmd/prof.go:628-661):
 * for each halo the fraction of solid angle in which the shell lies beyond radius
 * r, in `bins` log-radial bins of [r_min, r_max] = R200m * [RMinMult, RMaxMult].
 * Instead of `Samples` random directions the histogram counts one direction per
 * equal-area stratum, SFK_FRACTION_STRATA_THETA midpoints in cos(theta) x
 * SFK_FRACTION_STRATA_PHI in phi (262 144 directions; the default Samples is 50 000).
 *   rs, fs  [n_halo][bins]  bin centres and fractions */
#define SFK_FRACTION_STRATA_THETA 1024
#define SFK_FRACTION_STRATA_PHI 256
int sfk_angular_fraction(sfk_ctx *ctx, int64_t n_halo, int32_t order, const double *coeffs,
                         const double *r_min, const double *r_max, int32_t bins, double *rs, double *fs);
/* Shell.Axes from los/analyze/shell.go:140 on (inertia tensor, eigenvalues,
 * axis-ratio correction tables): moments = area-weighted means {xx, yy, zz, xy,
 * yz, zx, x, y, z}, out = {a, b, c, Ax, Ay, Az}.  Host arithmetic, no context. */
int sfk_axes_from_moments(const double moments[9], double out[6]);

int sfk_get_stats(const sfk_ctx *ctx, sfk_stats *out);

/* The cudaStream_t (as void*) every kernel of this context is launched on, so
 * a caller can record its own CUDA events around the calls. */
void *sfk_stream(const sfk_ctx *ctx);

/* ---- parity taps (valid after sfk_finish_snapshot) ---------------------- */

/* Splashback radius per LoS, [rings][spokes], NaN where RingBuffer.Oks is
 * false (ring_buffer.go:57-93).  Always available. */
int sfk_get_rsp(sfk_ctx *ctx, int64_t halo, double *r);
/* Number of particles queued for the halo (passing Halo.Intersect). */
int sfk_get_candidate_count(sfk_ctx *ctx, int64_t halo, int64_t *n);
/* ProfileRing.Insert calls of the halo that passed the early-out (profile_ring.go:92-95):
 * the halo's share of sfk_stats.n_intersections.  Always available. */
int sfk_get_intersection_count(sfk_ctx *ctx, int64_t halo, int64_t *n);
/* Halo.GetRhos for every LoS: [rings][spokes][bins].  Needs SFK_FLAG_TAPS. */
int sfk_get_profiles(sfk_ctx *ctx, int64_t halo, double *rho);
/* Integer taps, need SFK_FLAG_TAPS: inserts per LoS [rings][spokes] and the
 * number of start / end edges landing in each bin [rings][spokes][bins]. */
int sfk_get_counts(sfk_ctx *ctx, int64_t halo, uint32_t *n_insert,
                   uint32_t *start_edges, uint32_t *end_edges);
/* The halo's bin grid as the binning kernel left it: ProfileRing.derivs
 * (profile_ring.go:34-40) of every ring, [rings][spokes][bins], as int64 multiples of
 * *quantum (either output may be NULL).  Needs SFK_FLAG_KEEP_GRID. */
int sfk_get_grid(sfk_ctx *ctx, int64_t halo, int64_t *grid, double *quantum);
/* Host-side tables the context was built with: normals [rings][3], rotation
 * matrices rot / irot [rings][9]. */
int sfk_get_ring_tables(const sfk_ctx *ctx, float *norms, float *rot, float *irot);

#ifdef __cplusplus
}
#endif
#endif
