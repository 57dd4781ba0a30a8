// sfk_kernels.cuh -- launch wrappers of the sm_100a kernels (sfk_kernels.cu).
#pragma once

#include <cuda_runtime.h>

#include "sfk_internal.h"

namespace sfk {

struct CullArgs {
    const float *xyz;        // [n][3] block positions (device)
    const float *mass;       // [n], or null: every particle has uniformMass
    float uniformMass;       // buf.MinMass(), what the LGadget-2 / gotetra readers fill ms with
    int64_t n;
    int64_t sf3;             // SubsampleFactor^3, cmd/shell.go:490-493
    const int32_t *haloList; // halos touching the block, in halo order
    int nList;
    const HaloDev *halos;
    float tw;                // float32(TotalWidth), los/halo.go:172
    double rhoM;             // mean density of this block's cosmology
    // one pass: survivors are appended to one list (they carry their halo index) and
    // bucketed by halo at sfk_finish_snapshot
    int tma;                      // sf3 == 1 and both arrays 16-byte aligned: full tiles are staged by TMA
    Cand *cands;
    int64_t cap;                  // slots of cands
    unsigned long long *total;    // [1] candidates appended so far
    uint32_t *haloCount;          // [nHalo] candidates per halo
};

// Buckets the appended candidates by halo: sorted[offset[halo] + k] for the k-th arrival.
cudaError_t launch_bucket(const Cand *cands, int64_t n, const int64_t *offset, uint32_t *cursor, Cand *sorted,
                          cudaStream_t s);

struct BinArgs {
    const HitRec *recs;
    const int64_t *hitOffset;   // [nb * rings + 1] first slot of (halo, ring); the list owns the slots up to the next offset
    const uint32_t *hitCount;   // [nb][rings][2] records written from the front / from the back of the list (hits_kernel)
    const HaloDev *halos;
    const int32_t *batchHalos;  // [nb] halo index of each batch slot
    const double *ringCos, *ringSin;
    unsigned long long *grid;   // [nb][rings][spokes][bins] int64 fixed point
    unsigned long long *nIntersections; // [nHalo]
    uint32_t *tapInsert, *tapStart, *tapEnd; // TAPS only, indexed by halo
    const double *binInvG;      // [bins+1] 1/exp(k*dr): reciprocal bin edges in units of rMin
    double invDr60;             // 1/(60 dr), see bin_of
    float cK;                   // bins per octave: ln 2 / dr
    int rings, spokes, bins, chunks;
    // chunks == 1: tiles are stored, not added; edges falling into the next LoS's bin 0 go here
    unsigned long long *ovfList;   // [ovfCap][2] (grid cell, value)
    unsigned *ovfCount;            // [2] entries of this launch, entries ever dropped
    unsigned ovfCap;
};
// true when launch_bin writes every cell of the grid itself (no memset needed)
bool bin_stores_whole_grid(const BinArgs &a);

struct LosArgs {
    const unsigned long long *grid;
    const HaloDev *halos;
    const int32_t *batchHalos;
    const double *valTaps, *derivTaps;
    const double *logTab;   // [128][2] fast_log table (1/c, -ln(1/c)), sfk_tables.cpp
    double *rsp;        // [nHalo][rings][spokes]
    double *profiles;   // TAPS: [nHalo][rings][spokes][bins], else null
    int rings, spokes, bins, window;
    double dLim;
    // lines of sight [losBegin, losBegin + losCount) of the batch's flat (slot, ring, LoS) order;
    // rspBySlot: rsp is indexed by that flat order instead of by halo (ring-partitioned analysis)
    int64_t losBegin, losCount;
    int rspBySlot;
    // Lines of sight whose ln(rho) holds a long run of identical values (>= flatLanes whole lanes of
    // 8 bins: a profile clamped to the background, or without any kernel edge, over half a
    // smoothing window).  There the reference's outputs at consecutive bins are bit-identical --
    // identical window contents -- and its strict comparisons see exact ties, which the sliding
    // moments cannot reproduce; the moments kernel lists such lines and the literal tap loop
    // redoes them (launch_los).  Sparse halos only: none in configs[1].
    long long *flatList;         // [nb * rings * spokes]
    unsigned long long *flatCount;   // [1]
    unsigned long long *flatTotal;   // [1] lines handed over since sfk_begin_snapshot
    int listMode;                // tap-loop pass over flatList instead of [losBegin, losBegin + losCount)
    int flatLanes;               // run length, in whole lanes, that marks a line for the tap loop
};

struct FilterArgs {
    const double *rsp;
    const HaloDev *halos;
    const int32_t *batchHalos;
    double *fR;          // [nb][rings][spokes] kept radii
    int32_t *fIdx;       // [nb][rings][spokes] LoS index of each kept point
    int32_t *fCount;     // [nb][rings]
    int32_t *status;     // [nHalo]
    int rings, spokes, levels;
    double eta, dPhiUnused;
    const double *ringPhi;
    // (slot, ring) rows [rowBegin, rowBegin + rowCount) of the batch; rspBySlot as in LosArgs
    int rowBegin, rowCount, rspBySlot;
};

struct FitArgs {
    const double *fR;
    const int32_t *fIdx, *fCount;
    const int32_t *batchHalos;
    const float *irots;
    const double *ringCos, *ringSin;
    double *coeffs;  // [nHalo][P2]
    int32_t *status; // [nHalo]
    long long *nPoints; // [nHalo]
    // scratch (fit_scratch_doubles): design rows + r [nb][rings*spokes][P2+1], partial Gram
    // matrices [nb][chunks][P2*P2], inverse [nb][P2*P2], partial coefficients [nb][chunks][P2]
    double *rows, *gramPart, *gramInv, *csPart;
    int rings, spokes, order;
};

// PercentileProfile mode, cmd/shell.go:629-660
struct PercentileArgs {
    unsigned long long *grid;   // [nb][rings*spokes][bins] int64 derivative bins, prefix-summed in place
    const HaloDev *halos;
    const int32_t *batchHalos;
    double *rows;               // [nHalo][2*bins]; this stage fills [bins, 2*bins)
    int64_t nLos;               // rings*spokes
    int bins;
    long long nth;              // 1-indexed n of msort.NthLargest, math/sort/median.go:15-21
};

cudaError_t launch_percentile(const PercentileArgs &a, int nb, cudaStream_t s);

// `stats` M_sp pass, cmd/stats.go:451-543
struct MassArgs {
    const float *xyz, *mass;     // one particle block
    int64_t n;
    float tw2;                   // float32(TotalWidth) / 2, cmd/stats.go:518
    const float *centers;        // [nHalo][3] float32 sphere centres, cmd/stats.go:411-415
    const float *low2, *high2;   // float32(rLow*rLow), float32(rHigh*rHigh), :522
    const double *coeffs;        // [nHalo][P2]
    double *masses;              // [nHalo], accumulated across blocks
    unsigned *bboxKeys;          // [6] scratch: the block's coordinate range
    int *list, *listCount;       // [nHalo] + [1] scratch: halos this block can reach
    int nHalo, order, P2;
    // ShellParticleFile pass (launch_band): low2 / high2 are the widened band, cmd/stats.go:556-560
    const double *bandDelta;     // [nHalo] float64(sphere.R) * ShellWidth (0 when ShellWidth < 0)
    int bandInside;              // ShellWidth < 0: everything inside the surface
    long long *bandList;         // [bandCap] halo << 32 | particle index
    unsigned long long *bandCount, bandCap;
};
cudaError_t launch_mass(const MassArgs &a, cudaStream_t s);
cudaError_t launch_band(const MassArgs &a, cudaStream_t s);

// `stats` shell property integrals, los/analyze/shell.go:58-283
struct ShellPropArgs {
    const double *coeffs;        // [nHalo][2*order*order]
    const double *glX, *glW;     // [nTheta] Gauss-Legendre nodes / weights in cos(theta)
    double *sums;                // [nHalo][SS_COUNT], see sfk_shellmath.cuh
    int nHalo, order, nTheta, nPhi;
};
cudaError_t launch_shell_props(const ShellPropArgs &a, cudaStream_t s);

// `prof` angular-fraction profile, los/analyze/shell.go:295-335
struct ShellFractionArgs {
    const double *coeffs;        // [nHalo][2*order*order]
    const double *rMin, *rMax;   // [nHalo] profile range, cmd/prof.go:641-642
    double *fs;                  // [nHalo][bins]
    int nHalo, order, bins, nTheta, nPhi;
};
cudaError_t launch_shell_fraction(const ShellFractionArgs &a, cudaStream_t s);
cudaError_t launch_cull(const CullArgs &a, cudaStream_t s);
cudaError_t launch_hit_count(const Piece *pieces, int nPieces, const Cand *cands, const HaloDev *halos,
                             const float *norms, int rings, uint32_t *hitCount,
                             unsigned long long *rhoMaxBits, cudaStream_t s);
cudaError_t launch_hits(const Piece *pieces, int nPieces, const Cand *cands,
                        const HaloDev *halos, const float *norms, const float *rots,
                        const double *ringCos, const double *ringSin,
                        int rings, int spokes, double dPhi, const int64_t *hitOffset,
                        uint32_t *hitCursor, const int32_t *haloSlot, HitRec *recs, bool prune, cudaStream_t s);
cudaError_t launch_bin(const BinArgs &a, int nb, bool taps, cudaStream_t s);
struct SgPoly;   // sfk_sgmath.cuh
bool los_moments_supported(const LosArgs &a);
// sg != null: sliding-moment Savitzky-Golay where the shape allows it; null: the literal tap loop
cudaError_t launch_los(const LosArgs &a, const SgPoly *sg, int nb, cudaStream_t s);
cudaError_t launch_filter(const FilterArgs &a, int nb, cudaStream_t s);
cudaError_t launch_fit(const FitArgs &a, int nb, cudaStream_t s);
size_t bin_smem_bytes(int bins, bool taps);
int bin_tile_los(int bins, bool taps);   // LoS per CTA of the instantiation launch_bin picks
size_t fit_scratch_doubles(int nb, int rings, int spokes, int order, size_t *rows, size_t *gramPart, size_t *gramInv,
                           size_t *csPart);

}  // namespace sfk
