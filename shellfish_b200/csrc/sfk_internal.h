// sfk_internal.h -- shared declarations of the sm_100a shell-finder library.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/shellfish_b200.h"

namespace sfk {

// LoS per binning tile of the narrow instantiations (sfk_bin.cu; the product instantiation takes 64):
// TILE / 2 warps per CTA, every warp filters the ring's hit records into a private ring and bins them.
constexpr int kTileLos = 32;
constexpr int kKdeKnots = 100;   // rn and the KDE knot count, kde.go:70,99
constexpr int kMaxLevels = 6;    // KDE tree depth supported on the device
constexpr int kMaxOrder = 4;     // Penna order (2*order^2 <= 32 coefficients)
constexpr int kMaxRings = 1024;

// Per-halo constants, built on the host in float64 exactly as
// createHalos / Halo.Init / loadSphereVecs do (cmd/shell.go:530-564,423-465).
struct HaloDev {
    float x0, y0, z0;   // float32(origin), los/halo.go:156,169-171
    float lo2, hi2;     // Intersect bounds, los/halo.go:148-152
    int32_t valid;      // r200m > 0
    int32_t owned;      // analysed by this context (else only replayed in Transform)
    double rad;         // kernel radius RMax*RKernelMult/RMaxMult, cmd/shell.go:442
    double rad2;        // radius*radius, los/halo.go:236
    double sphVol;      // 4*pi/3*rad^3, cmd/shell.go:485
    double lowR, highR, dr;  // ProfileRing.Init, profile_ring.go:75-82
    double lrMin, dlr;       // GetRs, los/halo.go:365-366
    double rMax, rMin, invRMin;
    double defaultRho;
    double scale;       // fixed-point scale 2^e of the bin accumulators
    double quantum;     // 2^-e
};

// A particle that passed Halo.Intersect, already shifted to the halo frame.
struct Cand {
    float vx, vy, vz;   // wrapped position - float32(origin)
    int32_t halo;
    double rho;         // (m*sf^3/sphVol)/rhoM, cmd/shell.go:494-495
};
static_assert(sizeof(Cand) == 24, "Cand layout");

// One (particle, ring) pair that passed sphereIntersectRing, in ring frame.
struct alignas(16) HitRec {
    float cx, cy;       // RotateVec output (float32) in the ring plane, los/halo.go:231
    uint32_t r01, r23;  // idxRange result packed lo | hi << 16: [iLo1,iHi1), [iLo2,iHi2), :284-310;
                        // bit 31 of r23: the projected circle contains the centre, :241
    double rhoS;        // rho * scale
    double projRad2;    // max(radius^2 - cz^2, 0), :235-238 (cz itself is needed nowhere else)
};
constexpr uint32_t kHitCaseA = 0x80000000u;
static_assert(sizeof(HitRec) == 32, "HitRec layout");

// Work item of the hit kernels: <= 256 candidates of one halo.
struct Piece {
    int64_t start;      // index into the candidate array
    int32_t count;
    int32_t halo;
};

// Host-built constant tables (once per context).
struct Tables {
    std::vector<float> norms, rots, irots;     // [rings][3], [rings][9] x2
    std::vector<double> ringCos, ringSin, ringPhi;  // [spokes]
    double dPhi;
    // log-radial bin edges in units of rMin: binInvG[k] = 1/exp(k*dr0),
    // dr0 = ln(RMaxMult/RMinMult)/RadialBins (profile_ring.go:81 for every halo)
    std::vector<double> binInvG;
    double dr0;
};

void build_tables(const sfk_params &p, Tables &t);
double rho_average(double H0, double omegaM, double omegaL, double z);
// Savitzky-Golay value / first-derivative taps (order 4), smooth.go:84-96.
bool savgol_taps(int order, int width, int ld, double dx, std::vector<double> &taps);
double go_pow(double x, double y);
struct SgPoly;   // sfk_sgmath.cuh
void build_log_table(std::vector<double> &tab);
bool savgol_poly(const std::vector<double> &taps0, const std::vector<double> &taps1, SgPoly &out);
// `stats` shell properties (sfk_axes.cpp): quadrature rule and per-halo closing arithmetic.
void gauss_legendre(int n, std::vector<double> &x, std::vector<double> &w);
void axes_from_moments(const double m[9], double out[6]);
void shell_props_finalize(const double *sums, double *props);

}  // namespace sfk
