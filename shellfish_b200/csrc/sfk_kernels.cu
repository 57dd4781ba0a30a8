// sfk_kernels.cu -- hand-written sm_100a kernels of the shell-finder hot path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false.  FMA
// contraction is off for the whole file so that every float32 / float64
// expression rounds operation by operation, as the reference does on amd64;
// the float32 decision chain (wrap, cull, ring test, rotation) must do so for
// the integer taps to be bit-exact.
//
//   cull_kernel    Halo.Transform + Halo.Intersect     los/halo.go:147-197
//   hits_kernel    sphereIntersectRing + ring geometry los/halo.go:218-258,284-320
//   bin_kernel     per-LoS chords + ProfileRing.Insert los/halo.go:241-277, profile_ring.go:92-120
//   los_kernel     GetRhos + Smooth + SplashbackRadius halo.go:350, smooth.go:46, splashback_radius.go:30
//   filter_kernel  FilterPoints (KDE tree)             penna.go:115-165, kde.go
//   fit_*_kernel   PennaVolumeFit                      penna.go:14-108,170-194
#include <cstdint>
#include <cuda_runtime.h>

#include "sfk_kernels.cuh"
#include "sfk_device.cuh"
#include "sfk_sgmath.cuh"

namespace sfk {

namespace {

constexpr double kTwoPi = 6.28318530717958647692528676655900576839433879875021;
constexpr double kPi = 3.14159265358979323846264338327950288419716939937511;
constexpr double kTrigEps = 2e-5;   // see make_hit

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// Go's int(float64) on amd64 (CVTTSD2SQ): truncation toward zero.
__device__ __forceinline__ long long go_int(double x) { return __double2ll_rz(x); }

// ----------------------------------------------------------------- cull --

constexpr int kCullThreads = 256;
constexpr int kCullHaloChunk = 64;

struct CullHalo {
    float x0, y0, z0, lo2, hi2;
    double sphVol;
};

// One thread per (subsampled) particle.  The reference wraps the block's
// positions in place, halo after halo (Halo.Transform mutates sphBuf.xs,
// cmd/shell.go:441), so the position a halo sees depends on every earlier
// halo of the block: the loop over the block's halo list is sequential per
// particle and carries x, y, z in registers.  Survivors are appended to one
// list with a warp-aggregated slot claim; the bins are exact integers, so the
// order of a halo's candidates cannot change a result.
//
// A full tile of 256 particles (3 KB of positions + 1 KB of masses, contiguous in the block's
// arrays) is staged in shared memory by two TMA bulk copies (cp.async.bulk, completion on an
// mbarrier) issued by one thread: the [n][3] layout would otherwise cost three stride-3 scalar
// loads per thread.  The last, partial tile of a block and subsampled passes load directly.
__global__ void __launch_bounds__(kCullThreads) cull_kernel(CullArgs a) {
    __shared__ CullHalo sh[kCullHaloChunk];
    __shared__ int shIdx[kCullHaloChunk];
    __shared__ __align__(16) float sXyz[3 * kCullThreads];
    __shared__ __align__(16) float sMass[kCullThreads];
    __shared__ __align__(8) uint64_t bar;
    const int64_t tile0 = static_cast<int64_t>(blockIdx.x) * kCullThreads;
    const int64_t j = tile0 + threadIdx.x;
    const int64_t i = j * a.sf3;
    const bool live = i < a.n;
    float x = 0.f, y = 0.f, z = 0.f;
    double mScaled = 0.0;
    if (a.tma && tile0 + kCullThreads <= a.n) {   // uniform over the CTA
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_barrier_init();
            mbar_expect_tx(&bar, (a.mass ? 16 : 12) * kCullThreads);
            tma_load_1d(sXyz, a.xyz + 3 * tile0, 12 * kCullThreads, &bar);
            if (a.mass) tma_load_1d(sMass, a.mass + tile0, 4 * kCullThreads, &bar);
        }
        __syncthreads();   // the barrier is initialised before anybody polls it
        mbar_wait(&bar, 0);
        x = sXyz[3 * threadIdx.x];       // stride 3 is coprime to the 32 banks: conflict-free
        y = sXyz[3 * threadIdx.x + 1];
        z = sXyz[3 * threadIdx.x + 2];
        mScaled = static_cast<double>(a.mass ? sMass[threadIdx.x] : a.uniformMass);   // sf3 == 1
    } else if (live) {
        x = a.xyz[3 * i];
        y = a.xyz[3 * i + 1];
        z = a.xyz[3 * i + 2];
        // float64(ms[i]) * float64(sf^3), cmd/shell.go:494
        mScaled = static_cast<double>(a.mass ? a.mass[i] : a.uniformMass) * static_cast<double>(a.sf3);
    }
    const float tw = a.tw;
    const float tw2 = tw / 2;
    for (int base = 0; base < a.nList; base += kCullHaloChunk) {
        const int m = min(kCullHaloChunk, a.nList - base);
        __syncthreads();
        if (threadIdx.x < m) {
            const int hi = a.haloList[base + threadIdx.x];
            const HaloDev &h = a.halos[hi];
            sh[threadIdx.x] = CullHalo{h.x0, h.y0, h.z0, h.owned ? h.lo2 : INFINITY, h.hi2, h.sphVol};
            shIdx[threadIdx.x] = hi;
        }
        __syncthreads();
        for (int k = 0; k < m; k++) {
            const CullHalo h = sh[k];
            // Halo.Transform, los/halo.go:175-195
            float dx = x - h.x0, dy = y - h.y0, dz = z - h.z0;
            if (dx > tw2) x -= tw; else if (dx < -tw2) x += tw;
            if (dy > tw2) y -= tw; else if (dy < -tw2) y += tw;
            if (dz > tw2) z -= tw; else if (dz < -tw2) z += tw;
            // Halo.Intersect, los/halo.go:158-163
            dx = x - h.x0; dy = y - h.y0; dz = z - h.z0;
            const float r2 = (dx * dx + dy * dy) + dz * dz;
            const bool keep = live && r2 > h.lo2 && r2 < h.hi2;
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (mask == 0) continue;
            const int leader = __ffs(mask) - 1;
            const int halo = shIdx[k];
            unsigned long long off = 0;
            if (lane_id() == leader) {
                off = atomicAdd(a.total, static_cast<unsigned long long>(__popc(mask)));
                atomicAdd(&a.haloCount[halo], static_cast<uint32_t>(__popc(mask)));
            }
            off = __shfl_sync(0xffffffffu, off, leader);
            if (keep) {
                const long long at = static_cast<long long>(off) + __popc(mask & ((1u << lane_id()) - 1u));
                if (at < a.cap) {   // the host sizes the list from an upper bound: never false
                    Cand c;
                    c.vx = dx; c.vy = dy; c.vz = dz;   // Insert: vec - float32(origin)
                    c.halo = halo;
                    c.rho = (mScaled / h.sphVol) / a.rhoM;
                    a.cands[at] = c;
                }
            }
        }
    }
}

// Counting sort of the candidate list by halo (offsets from the per-halo counts of cull_kernel).
__global__ void __launch_bounds__(256) bucket_kernel(const Cand *cands, int64_t n, const int64_t *offset,
                                                     uint32_t *cursor, Cand *sorted) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * 256 + threadIdx.x;
    const bool live = i < n;
    Cand c;
    c.halo = -1 - static_cast<int>(lane_id());
    if (live) c = cands[i];
    // candidates arrive in runs of one halo: one cursor claim per (warp, halo)
    const unsigned peers = __match_any_sync(0xffffffffu, c.halo);
    if (!live) return;
    const int leader = __ffs(peers) - 1;
    uint32_t base = 0;
    if (static_cast<int>(lane_id()) == leader) base = atomicAdd(&cursor[c.halo], static_cast<uint32_t>(__popc(peers)));
    base = __shfl_sync(peers, base, leader);
    sorted[offset[c.halo] + base + __popc(peers & ((1u << lane_id()) - 1u))] = c;
}

// ----------------------------------------------------------------- hits --

constexpr int kHitThreads = 256;

__device__ __forceinline__ void rotate_vec(const float *m, float vx, float vy, float vz,
                                           float &ox, float &oy, float &oz) {
    // los/geom/rotation.go:60-65, left-to-right float32
    ox = (m[0] * vx + m[1] * vy) + m[2] * vz;
    oy = (m[3] * vx + m[4] * vy) + m[5] * vz;
    oz = (m[6] * vx + m[7] * vy) + m[8] * vz;
}

__device__ __forceinline__ uint16_t clamp_idx(long long v, int n) {
    return static_cast<uint16_t>(v < 0 ? 0 : (v > n ? n : v));
}

// Does any line of sight of a centre-containing hit run into the reference's
// NaN path?  oneValIntrDist takes sqrt(dist2 - b*b) (los/halo.go:337-346); for
// a LoS perpendicular to the projected position the radicand is ~0 and can
// round below zero, the distance becomes NaN and ProfileRing.Insert then adds
// rho to bin 0 and never subtracts it (profile_ring.go:93-110).  Only the LoS
// nearest to phiC +- pi/2 can qualify; they are evaluated literally.
__device__ bool has_nan_los(double cx, double cy, double projDist2, double phiC, double dPhi, double invDPhi,
                            int n, const double *ringCos, const double *ringSin) {
    bool any = false;
    for (int sgn = -1; sgn <= 1; sgn += 2) {
        double ph = phiC + sgn * (kPi / 2);
        if (ph < 0) ph += kTwoPi;
        const double t = ph * invDPhi, ft = floor(t);   // approximate on purpose: four neighbours are tested
        // b^2 can round above dist2 only for a LoS within ~3e-8 rad of the perpendicular; phiC
        // is good to 1e-6, so unless a LoS lies within 1e-5 rad there is nothing to test
        if (fmin(t - ft, ft + 1.0 - t) * dPhi > 1e-5) continue;
        const int i0 = static_cast<int>(ft);   // ph in [0, 2.5 pi): 0 <= i0 < 2n
        for (int d = -1; d <= 2; d++) {
            int i = i0 + d;   // in [-1, 2n + 2)
            if (i < 0) i += n;
            if (i >= n) i -= n;
            if (i >= n) i -= n;
            const double b = cy * ringCos[i] - cx * ringSin[i];
            const double b2 = b * b;
            if (projDist2 - b2 < 0) any = true;
        }
    }
    return any;
}

// Ring-frame record of one (particle, ring) hit: insertToRing up to the LoS
// loops (los/halo.go:231-258) and idxRange (:284-310).  Returns false for a
// hit that provably reaches no bin of any LoS -- the whole chord range
// [dist - prad, dist + prad] lies below rMin or above rMax, so every
// ProfileRing.Insert would take the early-out of profile_ring.go:93-95 -- and
// for centre-containing hits narrows the LoS range (all n in the reference) to
// a superset of the LoS whose far intersection exceeds rMin.  Both prunings
// only drop work the reference does to no effect; the per-LoS evaluation in
// bin_kernel stays literal.
//
// PRUNE == false (SFK_FLAG_NO_PRUNE, tests only) is the reference verbatim: all n LoS for a
// centre-containing hit, float64 asin / atan2 and idxRange as written for the others, nothing
// dropped.  tests/test_gpu_parity.py shows that both instantiations give identical integer bins.
struct HitTmp {   // make_hit_impl's working record
    float cx, cy, cz;
    uint32_t caseA, r01, r23;
    double rhoS;
};

template <bool PRUNE>
__device__ bool make_hit_impl(const float *rot, const Cand &c, const HaloDev &h, double dPhi, double invDPhi, int n,
                              const double *ringCos, const double *ringSin, HitTmp &r) {
    rotate_vec(rot, c.vx, c.vy, c.vz, r.cx, r.cy, r.cz);
    const double cx = r.cx, cy = r.cy, cz = r.cz;
    const double projDist2 = cx * cx + cy * cy;
    double projRad2 = h.rad2 - cz * cz;
    if (projRad2 < 0) projRad2 = 0;
    r.rhoS = c.rho * h.scale;
    if (!PRUNE) {
        r.caseA = projRad2 > projDist2 ? 1u : 0u;
        r.r01 = static_cast<uint32_t>(n) << 16; r.r23 = 0;
        if (r.caseA) return true;
        const double alpha = asin(sqrt(projRad2 / projDist2));   // halfAngularWidth, halo.go:315-318
        const double projPhi = atan2(cy, cx);
        const double phiLo = projPhi - alpha, phiHi = projPhi + alpha;
        long long i0, i1, i2, i3;   // idxRange, halo.go:284-310
        if (phiHi > kTwoPi) {
            i0 = go_int(phiLo / dPhi); i1 = n; i2 = 0; i3 = go_int((phiHi - kTwoPi) / dPhi) + 1;
        } else if (phiLo < 0) {
            i0 = go_int((phiLo + kTwoPi) / dPhi); i1 = n; i2 = 0; i3 = go_int(phiHi / dPhi) + 1;
        } else {
            i0 = go_int(phiLo / dPhi); i1 = go_int(phiHi / dPhi) + 1; i2 = 0; i3 = 0;
        }
        r.r01 = clamp_idx(i0, n) | static_cast<uint32_t>(clamp_idx(i1, n)) << 16;
        r.r23 = clamp_idx(i2, n) | static_cast<uint32_t>(clamp_idx(i3, n)) << 16;
        return true;
    }
    // D, P only enter comparisons guarded by 1e-9 relative margins or LoS-spacing guards: the
    // <= 1 ulp branch-free root serves
    const double D = sqrt_pos(projDist2), P = sqrt_pos(projRad2);
    const double rMinSafe = h.rMin * (1.0 - 1e-9), rMaxSafe = h.rMax * (1.0 + 1e-9);
    if (projRad2 > projDist2) {
        // circle contains the centre: Insert(-Inf, ln rHi) on every LoS, halo.go:241-250
        r.caseA = 1;
        r.r01 = static_cast<uint32_t>(n) << 16; r.r23 = 0;
        const double phiC = static_cast<double>(atan2f(r.cy, r.cx));   // only used with >= 1-LoS guards
        const bool dead = D + P <= rMinSafe;
        const double cosT = (h.rMin * h.rMin + projDist2 - projRad2) / (2.0 * h.rMin * D);
        const bool full = !(D > 1e-9 * h.rMin) || !(cosT > -1.0);
        if (!dead && full) return true;
        if (has_nan_los(cx, cy, projDist2, phiC, dPhi, invDPhi, n, ringCos, ringSin)) return true;
        if (dead) return false;
        // rHi(theta) > rMin  <=>  cos(theta) > cosT, theta = angle(LoS, projected position)
        const double theta = static_cast<double>(acosf(static_cast<float>(fmin(cosT, 1.0)))) + 2.0 * dPhi + 1e-5;
        if (theta >= kPi) return true;
        double a = phiC - theta;
        a -= kTwoPi * floor(a * (1.0 / kTwoPi));
        if (a < 0) a = 0;
        // guarded by theta's two LoS spacings: the reciprocal is exact enough
        const long long iLo = min(static_cast<long long>(n - 1), __double2ll_rd(a * invDPhi));
        const long long iHi = __double2ll_rd((a + 2.0 * theta) * invDPhi) + 1;
        if (iHi <= n) {
            r.r01 = static_cast<uint32_t>(iLo) | static_cast<uint32_t>(iHi) << 16;
        } else if (iHi - n < iLo) {
            r.r01 = static_cast<uint32_t>(iLo) | static_cast<uint32_t>(n) << 16;
            r.r23 = static_cast<uint32_t>(iHi - n) << 16;
        }
        return true;
    }
    r.caseA = 0;
    if (D + P <= rMinSafe || D - P >= rMaxSafe) return false;
    // halfAngularWidth and Atan2 (halo.go:252-253, 315-318) only feed idxRange's int()
    // truncations and the guarded trims below, so float32 libm (error < 3e-6 rad for a width
    // below asin(sqrt(0.98))) decides them unless a truncated value lies within kTrigEps of a
    // LoS boundary or the branch point phiLo = 0; those few hits take the float64 functions.
    double alpha, projPhi;
    bool literal;   // alpha, projPhi are the reference's float64 values
    {
        const float qf = static_cast<float>(projRad2) / static_cast<float>(projDist2);
        bool fast = qf < 0.98f;
        if (fast) {
            alpha = static_cast<double>(asinf(sqrtf(qf)));
            projPhi = static_cast<double>(atan2f(r.cy, r.cx));
            const double lo = projPhi - alpha, hi = projPhi + alpha;
            const double tlo = (lo < 0 ? lo + kTwoPi : lo) * invDPhi, thi = hi * invDPhi;
            const double tol = kTrigEps / dPhi;
            if (fabs(lo) < kTrigEps || fabs(tlo - rint(tlo)) < tol || fabs(thi - rint(thi)) < tol) fast = false;
        }
        if (!fast) {
            alpha = asin(sqrt(projRad2 / projDist2));
            projPhi = atan2(cy, cx);
        }
        literal = !fast;
    }
    // Radial trimming.  Along a LoS at angle delta from the projected position the chord is
    // [rLo, rHi](delta), rLo increasing and rHi decreasing in |delta| up to the tangent length
    // T = sqrt(D^2 - P^2) at delta = alpha.  Where the projected circle crosses the sphere r = R
    // the crossing is at cos(delta*) = (D^2 + R^2 - P^2)/(2 D R); it is the NEAR end of the chord
    // when T > R and the far end when T < R.  So for T > rMax every LoS beyond delta*(rMax) has
    // rLo >= rMax, for T < rMin every LoS beyond delta*(rMin) has rHi <= rMin, and
    // ProfileRing.Insert takes its early-out (profile_ring.go:93-95).  With a guard of two LoS
    // spacings the relative margin on the compared radius is >= 2 dPhi^2.
    double half = alpha;        // half-width handed to idxRange
    double edge = alpha + 2.0 * dPhi;   // first angle past which a LoS is provably dead
    {
        const double T2 = projDist2 - projRad2;
        double dStar = alpha;
        // float32 acos: its error (< 4e-4 rad even where cos -> 1) is inside the 2 dPhi guard
        if (T2 > rMaxSafe * rMaxSafe && D - P < h.rMax)
            dStar = acosf(fminf(fmaxf(static_cast<float>((projDist2 + h.rMax * h.rMax - projRad2) / (2.0 * D * h.rMax)), -1.f), 1.f));
        else if (T2 < rMinSafe * rMinSafe && D + P > h.rMin)
            dStar = acosf(fminf(fmaxf(static_cast<float>((projDist2 + h.rMin * h.rMin - projRad2) / (2.0 * D * h.rMin)), -1.f), 1.f));
        if (dStar + 2.0 * dPhi + 1e-3 < alpha) { half = dStar + 2.0 * dPhi + 1e-3; edge = half; literal = false; }
    }
    const double phiLo = projPhi - half, phiHi = projPhi + half;
    // idxRange, halo.go:284-310 (literal when half == alpha; otherwise a subset of the literal
    // ranges that still holds every LoS within `half` of the projected position)
    // The reference divides by dPhi; where the angles are not its float64 values anyway (float32
    // trig, kept away from LoS boundaries by kTrigEps, or a guarded trim) the reciprocal serves.
    long long i0, i1, i2, i3;
    if (literal) {
        if (phiHi > kTwoPi) {
            i0 = go_int(phiLo / dPhi); i1 = n; i2 = 0; i3 = go_int((phiHi - kTwoPi) / dPhi) + 1;
        } else if (phiLo < 0) {
            i0 = go_int((phiLo + kTwoPi) / dPhi); i1 = n; i2 = 0; i3 = go_int(phiHi / dPhi) + 1;
        } else {
            i0 = go_int(phiLo / dPhi); i1 = go_int(phiHi / dPhi) + 1; i2 = 0; i3 = 0;
        }
    } else if (phiLo < 0) {
        i0 = go_int((phiLo + kTwoPi) * invDPhi); i1 = n; i2 = 0; i3 = go_int(phiHi * invDPhi) + 1;
    } else {
        i0 = go_int(phiLo * invDPhi); i1 = go_int(phiHi * invDPhi) + 1; i2 = 0; i3 = 0;
    }
    r.r01 = clamp_idx(i0, n) | static_cast<uint32_t>(clamp_idx(i1, n)) << 16;
    r.r23 = clamp_idx(i2, n) | static_cast<uint32_t>(clamp_idx(i3, n)) << 16;
    if (phiLo < 0 && i3 <= 0) {
        // atan2 put the whole wedge below zero, so idxRange scans from its start all
        // the way to LoS n-1 (halo.go:297-303) and the reference drops the LoS past
        // the wedge one by one through the NaN test of :262.  A LoS at angle
        // delta from the projected position has b^2 = dist2 sin^2(delta), so it is
        // NaN whenever alpha < delta < pi - alpha; with a guard of two LoS spacings
        // on both sides the margin on sin(delta) is >= 2 sin^2(dPhi), nine orders of
        // magnitude above the rounding of the literal test.  Those LoS are cut out;
        // the ones near the anti-direction (delta > pi - alpha), where the reference
        // does insert (twoValIntrDist only sees b^2, so the chord is that of the
        // mirrored LoS pi - delta), stay in the second range.
        const double phiC = projPhi + kTwoPi;
        const double a1 = phiC + edge, a2 = phiC + kPi - edge;
        if (a1 < a2) {
            const long long deadLo = __double2ll_rd(a1 * invDPhi) + 1;   // first dead LoS
            const long long deadHi = __double2ll_rd(a2 * invDPhi);       // last dead LoS
            if (deadLo < n && deadLo > i0 && deadHi >= deadLo) {
                r.r01 = clamp_idx(i0, n) | static_cast<uint32_t>(clamp_idx(deadLo, n)) << 16;
                r.r23 = deadHi + 1 < n ? (clamp_idx(deadHi + 1, n) | static_cast<uint32_t>(n) << 16) : 0u;
            }
        }
    }
    return true;
}

template <bool PRUNE>
__device__ __forceinline__ bool make_hit(const float *rot, const Cand &c, const HaloDev &h, double dPhi, double invDPhi,
                                         int n, const double *ringCos, const double *ringSin, HitRec &out) {
    HitTmp r;
    if (!make_hit_impl<PRUNE>(rot, c, h, dPhi, invDPhi, n, ringCos, ringSin, r)) return false;
    out.cx = r.cx; out.cy = r.cy;
    out.r01 = r.r01;
    out.r23 = r.r23 | (r.caseA ? kHitCaseA : 0u);
    out.rhoS = r.rhoS;
    const double cz = r.cz;
    double projRad2 = h.rad2 - cz * cz;   // as in make_hit_impl and insertToRing, halo.go:235-238
    if (projRad2 < 0) projRad2 = 0;
    out.projRad2 = projRad2;
    return true;
}

// Raw hit counts: one CTA per Piece (<= 256 candidates of one halo).  Counts
// sphereIntersectRing successes per (halo, ring) -- an upper bound of the
// records hits_kernel will write -- and the halo's largest rho.
__global__ void __launch_bounds__(kHitThreads)
hit_count_kernel(const Piece *pieces, const Cand *cands, const HaloDev *halos, const float *norms,
                 int rings, uint32_t *hitCount, unsigned long long *rhoMaxBits) {
    extern __shared__ unsigned char smemRaw[];
    float *sNorm = reinterpret_cast<float *>(smemRaw);                     // [rings][3]
    uint32_t *sCnt = reinterpret_cast<uint32_t *>(sNorm + 3 * rings);      // [rings]
    const Piece pc = pieces[blockIdx.x];
    for (int k = threadIdx.x; k < 3 * rings; k += kHitThreads) sNorm[k] = norms[k];
    for (int k = threadIdx.x; k < rings; k += kHitThreads) sCnt[k] = 0;
    __syncthreads();
    const bool live = threadIdx.x < pc.count;
    Cand c;
    c.vx = c.vy = c.vz = 0.f; c.halo = pc.halo; c.rho = 0.0;
    if (live) c = cands[pc.start + threadIdx.x];
    const double rad = halos[pc.halo].rad;
    for (int ring = 0; ring < rings; ring++) {
        // sphereIntersectRing, los/halo.go:218-224 (dot in float32)
        const float d = (sNorm[3 * ring] * c.vx + sNorm[3 * ring + 1] * c.vy) + sNorm[3 * ring + 2] * c.vz;
        const double dot = static_cast<double>(d);
        const bool hit = live && dot < rad && dot > -rad;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask && lane_id() == 0) atomicAdd(&sCnt[ring], __popc(mask));
    }
    __syncthreads();
    for (int k = threadIdx.x; k < rings; k += kHitThreads)
        if (sCnt[k]) atomicAdd(&hitCount[static_cast<size_t>(pc.halo) * rings + k], sCnt[k]);
    // rho > 0: IEEE doubles order like unsigned integers
    unsigned long long bits = live ? static_cast<unsigned long long>(__double_as_longlong(c.rho)) : 0ull;
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, o);
        bits = other > bits ? other : bits;
    }
    if (lane_id() == 0 && bits) atomicMax(&rhoMaxBits[pc.halo], bits);
}

constexpr int kHitRingChunk = 32;   // rings per queue pass: 256 x 32 entries of 4 B = 32 KB

// Ring-frame records: one CTA per Piece.  Phase 1 tests every (candidate,
// ring) pair of a ring chunk and queues the hits in shared memory; phase 2
// walks the queue with all lanes busy (the hit rate is ~1/3, so computing the
// geometry in place would idle two lanes out of three), builds the record and
// claims a slot of the (halo, ring) list.
// The queue has two ends.  Whether the projected circle contains the centre
// (insertToRing's first branch) is |v|^2 < rad^2 up to the rounding of the
// rotation, i.e. a property of the PARTICLE, and the two branches of make_hit
// share almost no code; so hits of particles inside the kernel radius grow from
// the bottom of the queue and all others from the top, and a warp of phase 2
// runs one branch (the few misfiled entries at |v| ~ rad only cost divergence).
template <bool PRUNE>
__global__ void __launch_bounds__(kHitThreads, 4)
hits_kernel(const Piece *pieces, const Cand *cands, const HaloDev *halos, const float *norms,
            const float *rots, const double *ringCos, const double *ringSin, int rings, int spokes,
            double dPhi, const int64_t *hitOffset, uint32_t *hitCursor, const int32_t *haloSlot,
            HitRec *recs) {
    extern __shared__ __align__(16) unsigned char smemRaw[];
    Cand *sCand = reinterpret_cast<Cand *>(smemRaw);                          // [256]
    float *sNorm = reinterpret_cast<float *>(sCand + kHitThreads);            // [rings][3]
    uint32_t *queue = reinterpret_cast<uint32_t *>(sNorm + 3 * rings);        // [256 * kHitRingChunk]
    __shared__ uint32_t qCount[2];
    constexpr uint32_t qCap = kHitThreads * kHitRingChunk;
    const Piece pc = pieces[blockIdx.x];
    const HaloDev &h = halos[pc.halo];
    const int slot = haloSlot[pc.halo];
    for (int k = threadIdx.x; k < 3 * rings; k += kHitThreads) sNorm[k] = norms[k];
    const bool live = threadIdx.x < pc.count;
    Cand c;
    c.vx = c.vy = c.vz = 0.f; c.halo = pc.halo; c.rho = 0.0;
    if (live) c = cands[pc.start + threadIdx.x];
    sCand[threadIdx.x] = c;
    const double rad = h.rad;
    const double invDPhi = 1.0 / dPhi;
    const bool inner = static_cast<double>((c.vx * c.vx + c.vy * c.vy) + c.vz * c.vz) < h.rad2;
    for (int r0 = 0; r0 < rings; r0 += kHitRingChunk) {
        const int r1 = min(rings, r0 + kHitRingChunk);
        if (threadIdx.x < 2) qCount[threadIdx.x] = 0;
        __syncthreads();
        for (int ring = r0; ring < r1; ring++) {
            const float d = (sNorm[3 * ring] * c.vx + sNorm[3 * ring + 1] * c.vy) + sNorm[3 * ring + 2] * c.vz;
            const double dot = static_cast<double>(d);
            const bool hit = live && dot < rad && dot > -rad;
            const unsigned mask = __ballot_sync(0xffffffffu, hit);
            if (mask == 0) continue;
            const unsigned maskIn = __ballot_sync(0xffffffffu, hit && inner);
            const unsigned mine = inner ? maskIn : mask & ~maskIn;   // the hits of my end of the queue
            // lanes 0 and 1 claim the slots of the bottom and top end
            uint32_t base = 0;
            if (lane_id() < 2) {
                const int cnt = __popc(lane_id() == 0 ? maskIn : mask & ~maskIn);
                if (cnt) base = atomicAdd(&qCount[lane_id()], cnt);
            }
            base = __shfl_sync(0xffffffffu, base, inner ? 0 : 1);
            if (hit) {
                const uint32_t at = base + __popc(mine & ((1u << lane_id()) - 1u));
                queue[inner ? at : qCap - 1 - at] = threadIdx.x | (ring << 8);
            }
        }
        __syncthreads();
        for (int end = 0; end < 2; end++) {
        const uint32_t nq = qCount[end];
        for (uint32_t e0 = threadIdx.x & ~31u; e0 < nq; e0 += kHitThreads) {
            const uint32_t e = e0 + lane_id();
            const uint32_t q = e < nq ? queue[end == 0 ? e : qCap - 1 - e] : 0u;
            const int ring = static_cast<int>(q >> 8);
            HitRec rec;
            const bool alive = e < nq && make_hit<PRUNE>(rots + 9 * ring, sCand[q & 0xffu], h, dPhi, invDPhi, spokes,
                                                  ringCos, ringSin, rec);
            // A (halo, ring) list is filled from both ends: circles that contain the centre or whose
            // tangent length stays below rMin (rLo <= sqrt(dist^2 - prad^2) <= rMin on every LoS: each
            // chord starts in bin 0) from the front, the others from the back, so that a warp of
            // bin_kernel mostly runs ONE branch of ProfileRing.Insert's start edge.  The class is a
            // scheduling hint in float32: every record is evaluated literally wherever it lands.
            const bool outer = alive && !(rec.r23 & kHitCaseA) &&
                               (rec.cx * rec.cx + rec.cy * rec.cy) - static_cast<float>(rec.projRad2) >
                                   static_cast<float>(h.rMin * h.rMin);
            // one slot claim per (warp, ring, end): queue order keeps a ring's hits together
            const unsigned peers = __match_any_sync(0xffffffffu, alive ? 2 * ring + (outer ? 1 : 0)
                                                                       : -1 - static_cast<int>(lane_id()));
            if (alive) {
                const size_t hr = static_cast<size_t>(slot) * rings + ring;
                const int leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (static_cast<int>(lane_id()) == leader) base = atomicAdd(&hitCursor[2 * hr + (outer ? 1 : 0)], __popc(peers));
                base = __shfl_sync(peers, base, leader) + __popc(peers & ((1u << lane_id()) - 1u));
                recs[outer ? hitOffset[hr + 1] - 1 - base : hitOffset[hr] + base] = rec;
            }
        }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------ los --

constexpr int kLosWarps = 8;
// columns of the padded ln(rho) rows of los_kernel_fast<8> (256 bins): positions
// 0 .. B + window + 8 must fit in 8 rows, so the fast path serves windows up to 224 (the moments path shifts by up to 7)
constexpr int kLosFastRow = 64;
constexpr int kLosFastMaxWindow = 8 * kLosFastRow - 256 - 32;

__device__ __forceinline__ double nan_f64() { return __longlong_as_double(0x7ff8000000000000ll); }

// ln(x) for a normal x > 0: x = 2^e m, m in [1, 2); the top 7 mantissa bits pick
// c = 1 + (idx + 1/2)/128 from a table of (1/c rounded, -ln of that rounded value), so that
// ln m = tab.y + log1p(m * tab.x - 1) holds exactly, and |m * tab.x - 1| < 2^-8 leaves a
// degree-6 series (truncation 2e-18).  Absolute error ~2e-16, as libm's log for |ln x| > 1;
// used only where the value decides nothing beyond tie level (the smoothed profile).
__device__ __forceinline__ double fast_log(double x, const double2 *tab) {
    const int hi = __double2hiint(x);
    const int e = (hi >> 20) - 1023;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
    const double2 t = tab[(hi >> 13) & 127];
    const double r = fma(m, t.x, -1.0);
    double L = fma(r, -1.0 / 6.0, 0.2);
    L = fma(L, r, -0.25);
    L = fma(L, r, 1.0 / 3.0);
    L = fma(L, r, -0.5);
    L = fma(L, r, 1.0);
    return fma(static_cast<double>(e), 0.693147180559945309417232121458, fma(L, r, t.y));
}

// Fast path of los_kernel for bins == 32 * PER: every lane keeps its PER
// consecutive bins in registers.
//
// Smooth (smooth.go:46-81) returns vals = exp(K0 * ln rho) and derivs = K1 * ln rho, and
// SplashbackRadius (splashback_radius.go:30-58) only COMPARES vals with each other, so the
// moments path takes the running minimum on K0 * ln rho itself: exp is monotone, NaN and +-Inf
// order the same way, and the 256 exp() per line of sight are gone.  (Lines where smoothed
// values can differ by rounding noise only are not decided there, see the flat-run hand-over.)
//
// MOMENTS == true (default whenever every ln rho is finite, i.e. rho_bg > 0): the two
// Savitzky-Golay FIRs are evaluated through the five sliding polynomial moments of
// sfk_sgmath.cuh -- one direct pass over the window for the lane's first output (taps +-k
// paired: 7 operations per two taps), PER - 1 binomial shifts for the rest -- and ln() is
// fast_log.  Same values as the tap loop to ~1e-13 (tests/test_sgmoments.py), same R_sp.
// MOMENTS == false: ConvolveAt verbatim (filter.go:104-161) as a register sliding window:
// per tap one shared-memory read of ln(rho) per lane (conflict-free permuted layout) and one
// broadcast read of the tap pair feed 2*PER FMAs; libm log, so that the -Inf / NaN
// propagation of Gadget-2's MinMass() == 0 stays sample-exact.
// The steepest-slope search is the running-minimum scan of splashback_radius.go:30-58 done
// as a warp prefix-min plus an arg-min.
template <int PER, bool MOMENTS>
__global__ void __launch_bounds__(kLosWarps * 32) los_kernel_fast(LosArgs a, int nb, SgPoly sg) {
    const int i0 = (threadIdx.x & 31) * PER;
    extern __shared__ double sm[];
    constexpr int B = 32 * PER;
    const int window = a.window;
    const int center = window / 2;
    // ln(rho) with the Extension boundary (filter.go:104-161) materialised: sample i sits at
    // position p = i + padL, padL = center + 1, the positions below / above hold sample 0 / B-1.
    // Position p is stored at row p % PER, column p / PER, so the FIR below reads row jj at
    // column (lane + const): conflict-free and with no index arithmetic.
    // The moments path shifts everything by `shift` positions so that the window centre of a
    // lane's first output sits on row 0 (see the anchor below).
    const int shift = MOMENTS ? (PER - ((1 + center) % PER)) % PER : 0;
    const int padL = center + 1 + shift;
    constexpr int rowLen = kLosFastRow;
    double2 *taps = reinterpret_cast<double2 *>(sm);                       // tap loop: [window] (value, deriv)
    double *upow = sm;                                                     // moments: [center + 1][4] u, u^2, u^3, u^4 of tap pair k
    const int tabOff = MOMENTS ? 4 * (center + 1) : 2 * window;
    const double2 *logTab = reinterpret_cast<const double2 *>(sm + tabOff);   // moments: [128]
    double *ys = sm + tabOff + (MOMENTS ? 256 : 0) + (threadIdx.x >> 5) * (PER * rowLen);
    if (MOMENTS) {
        for (int k = threadIdx.x; k <= center; k += blockDim.x) {
            const double u = static_cast<double>(k) / static_cast<double>(center);
            upow[4 * k] = u; upow[4 * k + 1] = u * u; upow[4 * k + 2] = u * u * u; upow[4 * k + 3] = (u * u) * (u * u);
        }
        for (int k = threadIdx.x; k < 256; k += blockDim.x) sm[tabOff + k] = a.logTab[k];
    } else {
        for (int k = threadIdx.x; k < window; k += blockDim.x) taps[k] = make_double2(a.valTaps[k], a.derivTaps[k]);
    }
    __syncthreads();
    const int lane = lane_id();
    const int64_t gw0 = static_cast<int64_t>(blockIdx.x) * kLosWarps + (threadIdx.x >> 5);
    const int64_t perHalo = static_cast<int64_t>(a.rings) * a.spokes;
    int64_t gw;
    if (!MOMENTS && a.listMode) {   // second pass: the lines the moments kernel handed over
        if (gw0 >= static_cast<int64_t>(*a.flatCount)) return;
        gw = a.flatList[gw0];
        if (gw0 == 0 && lane_id() == 0) atomicAdd(a.flatTotal, *a.flatCount);
    } else {
        if (gw0 >= a.losCount) return;
        gw = a.losBegin + gw0;
    }
    const int slot = static_cast<int>(gw / perHalo);
    const int64_t rl = gw - slot * perHalo;
    const int haloIdx = a.batchHalos[slot];
    const HaloDev &h = a.halos[haloIdx];
    const longlong2 *g = reinterpret_cast<const longlong2 *>(a.grid + (static_cast<size_t>(slot) * perHalo + rl) * B);

    // Retrieve + GetRhos: exact int64 prefix sum, clamp, ln
    long long d[PER];
#pragma unroll
    for (int c = 0; c < PER; c += 2) {
        const longlong2 v = g[(lane * PER + c) / 2];
        d[c] = v.x; d[c + 1] = v.y;
    }
    long long run = 0;
#pragma unroll
    for (int c = 0; c < PER; c++) run += d[c];
    long long incl = run;
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    long long acc = incl - run;
    const double q = h.quantum, bg = h.defaultRho;
    double yFirst = 0.0, yLast = 0.0;
    bool flatLane = true;   // my PER values are all equal
#pragma unroll
    for (int c = 0; c < PER; c++) {
        acc += d[c];
        double rho = static_cast<double>(acc) * q;
        if (rho < bg) rho = bg;
        if (a.profiles) a.profiles[(static_cast<size_t>(haloIdx) * perHalo + rl) * B + lane * PER + c] = rho;
        const double y = MOMENTS ? fast_log(rho, logTab) : log(rho);
        const int p = lane * PER + c + padL;
        ys[(p % PER) * rowLen + p / PER] = y;
        if (c == 0) yFirst = y;
        if (MOMENTS && c > 0) flatLane = flatLane && y == yLast;
        yLast = y;
    }
    if (MOMENTS) {
        // Exact ties.  Two outputs see identical window contents once center + 2 consecutive values
        // are equal (at an array end the Extension boundary repeats the end sample); such a run
        // covers >= (center - 6) / PER whole lanes, each flat and equal to the next lane's first
        // value.  Sparse lines of sight also repeat whole step patterns -- the same two plateau
        // levels around an edge at the same fraction of a bin -- whose smoothed derivatives are
        // equal in exact arithmetic and ordered by rounding alone.  Both need long plateaus, so a
        // line is handed to the literal instantiation as soon as it holds flatLanes <= 2 whole
        // flat lanes in a row (>= 16 equal bins); profiles of >= 1e4-particle halos have none.
        const double nextFirst = __shfl_down_sync(0xffffffffu, yFirst, 1);
        const unsigned fm = __ballot_sync(0xffffffffu, flatLane && (lane == 31 || yLast == nextFirst));
        unsigned run = fm;
        for (int k = 1; k < a.flatLanes; k++) run &= fm >> k;
        if (run != 0 && lane == 0) a.flatList[atomicAdd(a.flatCount, 1ull)] = gw;
    }
    // Extension: repeat the end samples
    yFirst = __shfl_sync(0xffffffffu, yFirst, 0);
    yLast = __shfl_sync(0xffffffffu, yLast, 31);
    for (int e = lane; e < padL; e += 32) ys[(e % PER) * rowLen + e / PER] = yFirst;
    for (int e = B + padL + lane; e < PER * rowLen; e += 32) ys[(e % PER) * rowLen + e / PER] = yLast;
    __syncwarp();

    double *out = a.rspBySlot ? a.rsp + gw : a.rsp + static_cast<size_t>(haloIdx) * perHalo + rl;
    if (B <= window) {   // Smooth: ok = false, smooth.go:51-53
        if (lane == 0) *out = nan_f64();
        return;
    }
    // Output i0 + c reads positions lane*PER + shift + c + 1 + j, j = 0 .. window-1 (tap j, sample
    // i0 + c - center + j); position lane*PER + q sits at row q % PER, column lane + q / PER.
    const double *yl = ys + lane;
    double s0[PER], s1[PER];
    if (MOMENTS) {
        // Anchor: the five moments of output c = 0.  Its window centre is position qc = shift + 1 +
        // center, a multiple of PER by the choice of `shift`, i.e. row 0 of column c0.  Taps +-k are
        // paired (7 operations per two taps); with k = PER*m + j the sample at qc + k sits at row j,
        // column c0 + m, and the one at qc - k at row (PER - j) % PER, column c0 - m - (j > 0): the
        // rows are compile-time offsets and only the columns move (no index arithmetic per tap).
        const int c0 = (shift + 1 + center) / PER;
        double M[5] = {yl[c0], 0.0, 0.0, 0.0, 0.0};
        for (int m = 0; m * PER <= center; m++) {
            const double *yp = yl + c0 + m, *ym = yl + c0 - m;
            const double *ub = upow + 4 * PER * m;
#pragma unroll
            for (int j = 0; j < PER; j++) {
                const int k = m * PER + j;
                if ((j > 0 || m > 0) && k <= center) {
                    // (u, u^2) from the table, u^3 and u^4 formed here: the pass is bound by the
                    // shared-memory pipe, the FP64 pipe has room
                    const double2 u12 = *reinterpret_cast<const double2 *>(ub + 4 * j);
                    const double up[4] = {u12.x, u12.y, u12.x * u12.y, u12.y * u12.y};
                    sg_accumulate_pair(M, yp[j * rowLen], ym[((PER - j) % PER) * rowLen - (j > 0 ? 1 : 0)], up);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < PER; c++) {
            sg_outputs(M, sg, s0[c], s1[c]);
            if (c < PER - 1) {
                const int qo = shift + c + 1, qi = shift + c + 1 + window;
                sg_advance(M, sg, yl[(qo % PER) * rowLen + qo / PER], yl[(qi % PER) * rowLen + qi / PER]);
            }
        }
    } else {
        // ConvolveAt with Extension (filter.go:104-161), taps ascending: w[m % PER] holds
        // position lane*PER + 1 + m for m = j .. j + PER - 1.
        double w[PER];
#pragma unroll
        for (int c = 0; c < PER; c++) { s0[c] = 0.0; s1[c] = 0.0; }
#pragma unroll
        for (int m = 0; m < PER - 1; m++) w[m] = yl[(1 + m) * rowLen];
        int jb = 0;
        for (; jb + PER <= window; jb += PER) {
            const double *yj = yl + jb / PER + 1;   // position lane*PER + jb + PER + jj: row jj, column lane + jb/PER + 1
#pragma unroll
            for (int jj = 0; jj < PER; jj++) {
                w[(jj + PER - 1) % PER] = yj[jj * rowLen];
                const double2 tp = taps[jb + jj];
#pragma unroll
                for (int c = 0; c < PER; c++) {
                    const double x = w[(c + jj) % PER];
                    s0[c] = s0[c] + x * tp.x;   // multiply, round, add: Go on amd64 does not fuse,
                    s1[c] = s1[c] + x * tp.y;   // and equal-step ties are decided by these roundings
                }
            }
        }
        {   // the last window % PER taps
            const double *yj = yl + jb / PER + 1;
#pragma unroll
            for (int jj = 0; jj < PER; jj++) {
                if (jb + jj < window) {
                    w[(jj + PER - 1) % PER] = yj[jj * rowLen];
                    const double2 tp = taps[jb + jj];
#pragma unroll
                    for (int c = 0; c < PER; c++) {
                        const double x = w[(c + jj) % PER];
                        s0[c] = s0[c] + x * tp.x;
                        s1[c] = s1[c] + x * tp.y;
                    }
                }
            }
        }
    }
    // s0 = ln vals, s1 = d ln rho / d ln r.  The literal instantiation forms vals = exp(s0) as Smooth
    // does: exp is monotone but not injective in floating point -- for |s0| < 1 several neighbouring
    // doubles share one exp -- so where smoothed values differ by rounding noise only (the flat
    // windows this instantiation is handed) the reference sees exact ties that a comparison of the
    // logarithms would break.
    if (!MOMENTS) {
#pragma unroll
        for (int c = 0; c < PER; c++) s0[c] = exp(s0[c]);
    }

    // SplashbackRadius: rhoMin starts at vals[0]; i = 1 .. B-2 is a candidate when
    // vals[i] < min(vals[0..i-1]) and derivs[i] is a strict local minimum below dLim.
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    const double v0 = __shfl_sync(0xffffffffu, s0[0], 0);
    double lmin = inf;
#pragma unroll
    for (int c = 0; c < PER; c++) { const double v = s0[c]; if (v < lmin) lmin = v; }
    double excl = lmin;   // inclusive prefix min over lanes, then shifted
    for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, excl, o);
        if (lane >= o && up < excl) excl = up;
    }
    excl = __shfl_up_sync(0xffffffffu, excl, 1);
    if (lane == 0) excl = inf;
    const double dPrev = __shfl_up_sync(0xffffffffu, s1[PER - 1], 1);
    const double dNext = __shfl_down_sync(0xffffffffu, s1[0], 1);
    double runMin = excl, best = a.dLim;
    int bestI = -1;
#pragma unroll
    for (int c = 0; c < PER; c++) {
        const int i = i0 + c;
        const double v = s0[c];
        if (i == 0) { runMin = v; continue; }           // rhoMin := rhos[0]
        if (v < runMin) {
            runMin = v;
            const double dm = c == 0 ? dPrev : s1[c == 0 ? 0 : c - 1];
            const double dp = c == PER - 1 ? dNext : s1[c == PER - 1 ? c : c + 1];
            const double dd = s1[c];
            if (i < B - 1 && dd < dp && dd < dm && dd < best) { best = dd; bestI = i; }
        }
    }
    // vals[0] NaN poisons the running minimum: no candidate anywhere
    if (v0 != v0) bestI = -1;
    // arg-min over lanes, earliest index on ties (sequential update semantics)
    for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bestI, o);
        const bool take = oi >= 0 && (bestI < 0 || ob < best || (ob == best && oi < bestI));
        if (take) { best = ob; bestI = oi; }
    }
    if (lane == 0)
        *out = bestI < 0 ? nan_f64() : exp(h.lrMin + h.dlr * (static_cast<double>(bestI) + 0.5));
}

// One warp per line of sight: prefix sum of the bin deltas (exact, in int64),
// clamp to the background density, ln, two Savitzky-Golay FIRs with edge
// extension, exp, steepest-slope search.
__global__ void __launch_bounds__(kLosWarps * 32) los_kernel(LosArgs a, int nb) {
    extern __shared__ double sm[];
    const int bins = a.bins, window = a.window;
    double *taps0 = sm;                 // [window]
    double *taps1 = taps0 + window;     // [window]
    double *wbuf = taps1 + window + (threadIdx.x >> 5) * 3 * bins;
    double *ys = wbuf, *vals = wbuf + bins, *ders = wbuf + 2 * bins;
    for (int k = threadIdx.x; k < window; k += blockDim.x) {
        taps0[k] = a.valTaps[k];
        taps1[k] = a.derivTaps[k];
    }
    __syncthreads();
    const int lane = lane_id();
    const int64_t gw0 = static_cast<int64_t>(blockIdx.x) * kLosWarps + (threadIdx.x >> 5);
    const int64_t perHalo = static_cast<int64_t>(a.rings) * a.spokes;
    if (gw0 >= a.losCount) return;
    const int64_t gw = a.losBegin + gw0;
    const int slot = static_cast<int>(gw / perHalo);
    const int64_t rl = gw - slot * perHalo;   // ring*spokes + los
    const int haloIdx = a.batchHalos[slot];
    const HaloDev &h = a.halos[haloIdx];
    const unsigned long long *g = a.grid + (static_cast<size_t>(slot) * perHalo + rl) * bins;

    // Retrieve + GetRhos (profile_ring.go:124-130, halo.go:350-357).  Each lane
    // owns a contiguous run of bins; the carry between lanes is a warp scan.
    const int per = (bins + 31) / 32;
    const int b0 = lane * per, b1 = min(bins, b0 + per);
    long long run = 0;
    for (int k = b0; k < b1; k++) run += static_cast<long long>(g[k]);
    long long incl = run;
    for (int o = 1; o < 32; o <<= 1) {
        const long long up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    long long acc = incl - run;
    const double q = h.quantum, bg = h.defaultRho;
    for (int k = b0; k < b1; k++) {
        acc += static_cast<long long>(g[k]);
        double rho = static_cast<double>(acc) * q;
        if (rho < bg) rho = bg;
        if (a.profiles) a.profiles[(static_cast<size_t>(haloIdx) * perHalo + rl) * bins + k] = rho;
        ys[k] = log(rho);
    }
    __syncwarp();

    double *out = a.rspBySlot ? a.rsp + gw : a.rsp + static_cast<size_t>(haloIdx) * perHalo + rl;
    if (bins <= window) {  // Smooth: ok = false, smooth.go:51-53
        if (lane == 0) *out = __longlong_as_double(0x7ff8000000000000ll);
        return;
    }
    // ConvolveAt with Extension, filter.go:104-161: taps ascending, mul then add
    const int center = window / 2;
    for (int i = lane; i < bins; i += 32) {
        double s0 = 0.0, s1 = 0.0;
        for (int j = 0; j < window; j++) {
            int idx = i + j - center;
            idx = idx < 0 ? 0 : (idx >= bins ? bins - 1 : idx);
            const double x = ys[idx];
            s0 += x * taps0[j];
            s1 += x * taps1[j];
        }
        vals[i] = exp(s0);   // Smooth's vals, smooth.go:75
        ders[i] = s1;
    }
    __syncwarp();
    // SplashbackRadius, splashback_radius.go:30-58 (sequential semantics)
    if (lane == 0) {
        double rhoMin = vals[0];
        double dMin = a.dLim;
        int iMin = -1;
        for (int i = 1; i < bins - 1; i++) {
            const double v = vals[i];
            if (v < rhoMin) {
                rhoMin = v;
                const double d = ders[i];
                if (d < ders[i + 1] && d < ders[i - 1] && d < dMin) { dMin = d; iMin = i; }
            }
        }
        // rs[iMin], halo.go:360-370
        *out = iMin < 0 ? __longlong_as_double(0x7ff8000000000000ll)
                        : exp(h.lrMin + h.dlr * (static_cast<double>(iMin) + 0.5));
    }
}

// --------------------------------------------------------------- splines --

// math/interpolate/spline.go:203-237,266-294: natural cubic spline second
// derivatives.  Sequential (one thread).  tmp has n entries.
__device__ bool spline_y2(const double *xs, const double *ys, int n, double *y2, double *tmp) {
    y2[0] = 0; y2[n - 1] = 0;
    const int m = n - 2;
    if (m <= 0) return true;
    // TriDiagAt with as/bs/cs/rs formed on the fly (same expressions, same order)
    auto A = [&](int i) { return (xs[i + 1] - xs[i]) / 6; };
    auto Bq = [&](int i) { return (xs[i + 2] - xs[i]) / 3; };
    auto Cq = [&](int i) { return (xs[i + 2] - xs[i + 1]) / 6; };
    auto Rq = [&](int i) {
        return ((ys[i + 2] - ys[i + 1]) / (xs[i + 2] - xs[i + 1])) -
               ((ys[i + 1] - ys[i]) / (xs[i + 1] - xs[i]));
    };
    double *out = y2 + 1;
    double beta = Bq(0);
    if (beta == 0) return false;
    out[0] = Rq(0) / beta;
    for (int i = 1; i < m; i++) {
        tmp[i] = Cq(i - 1) / beta;
        beta = Bq(i) - A(i) * tmp[i];
        if (beta == 0) return false;
        out[i] = (Rq(i) - A(i) * out[i - 1]) / beta;
    }
    for (int i = m - 2; i >= 0; i--) out[i] -= tmp[i + 1] * out[i + 1];
    return true;
}

// Spline.Eval + bsearch, spline.go:80-97,176-200 (increasing tables only).
__device__ bool spline_eval(const double *xs, const double *ys, const double *y2, int n,
                            double x, double &y) {
    if (x <= xs[0] || x >= xs[n - 1]) {
        if (x == xs[0]) { y = ys[0]; return true; }
        if (x == xs[n - 1]) { y = ys[n - 1]; return true; }
        return false;
    }
    const double dxEst = (xs[n - 1] - xs[0]) / static_cast<double>(n - 1);
    int i;
    const long long guess = go_int((x - xs[0]) / dxEst);
    if (guess >= 0 && guess < n - 1 && xs[guess] <= x && xs[guess + 1] >= x) {
        i = static_cast<int>(guess);
    } else {
        int lo = 0, hi = n - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) / 2;
            if (x >= xs[mid]) lo = mid; else hi = mid;
        }
        if (lo == n - 1) return false;
        i = lo;
    }
    // calcCoeffs, spline.go:228-237
    const double h = xs[i + 1] - xs[i];
    const double ca = (-y2[i] / 6 + y2[i + 1] / 6) / h;
    const double cb = y2[i] / 2;
    const double cc = (ys[i + 1] - ys[i]) / h + h * (-y2[i] / 3 - y2[i + 1] / 6);
    const double cd = ys[i];
    const double dx = x - xs[i];
    y = ca * dx * dx * dx + cb * dx * dx + cc * dx + cd;
    return true;
}

// --------------------------------------------------------------- filter --

constexpr int kFilterThreads = 128;
constexpr int kMaxKde = (1 << (kMaxLevels + 1)) - 1;
constexpr int kMaxAnchor = (1 << kMaxLevels) + 10;
constexpr int kMaxMax = kKdeKnots / 2;

struct FilterSmem {
    double X[kKdeKnots], spRs[kKdeKnots];
    double th[kMaxAnchor], mx[kMaxAnchor], ay2[kMaxAnchor], atmp[kMaxAnchor];
    double conn[kMaxLevels + 1][1 << kMaxLevels];
    int nMax[kMaxKde];
    int status, kept, nValid, anchorN;
    double high, hCur;
};

// GetRFunc(level, Radial): kde.go:262-331.  Builds the (theta, r) spline of
// the anchors of `level` into sm.th/mx/ay2.  One thread.
__device__ bool build_rfunc(FilterSmem &sm, int level) {
    const int n = 1 << level;
    int buf = 5;
    if (buf > n) buf = n;
    auto finest = [&](int idx) {
        for (int i = 0; i <= level; i++) {
            const double r = sm.conn[level - i][idx >> i];
            if (!isnan(r)) return r;
        }
        return __longlong_as_double(0x7ff8000000000000ll);
    };
    auto theta = [&](int j) { return kTwoPi * (static_cast<double>(j) + 0.5) / static_cast<double>(n); };
    auto thAt = [&](int j) { return level == 0 ? kPi : theta(j); };
    int k = 0;
    for (int j = n - buf; j < n; j++, k++) { sm.th[k] = thAt(j) - kTwoPi; sm.mx[k] = finest(j); }
    for (int j = 0; j < n; j++, k++) { sm.th[k] = thAt(j); sm.mx[k] = finest(j); }
    for (int j = 0; j < buf; j++, k++) { sm.th[k] = thAt(j) + kTwoPi; sm.mx[k] = finest(j); }
    sm.anchorN = k;
    for (int i = 0; i < k - 1; i++) if (sm.th[i + 1] < sm.th[i]) return false;
    return spline_y2(sm.th, sm.mx, k, sm.ay2, sm.atmp);
}

// One CTA per (halo, ring): FilterPoints for that ring, penna.go:115-165.
__global__ void __launch_bounds__(kFilterThreads) filter_kernel(FilterArgs a) {
    extern __shared__ __align__(16) unsigned char raw[];
    FilterSmem &sm = *reinterpret_cast<FilterSmem *>(raw);
    const int nKde = (1 << (a.levels + 1)) - 1;
    double *Y = reinterpret_cast<double *>(raw + sizeof(FilterSmem));   // [nKde][100]
    double *Y2 = Y + nKde * kKdeKnots;                                 // [nKde][100]
    double *V = Y2 + nKde * kKdeKnots;                                 // [nKde][100] (tmp, then values)
    double *maxes = V + nKde * kKdeKnots;                              // [nKde][kMaxMax]
    double *pR = maxes + nKde * kMaxMax;                               // [spokes]
    int *pIdx = reinterpret_cast<int *>(pR + a.spokes);                // [spokes]
    short *pLo = reinterpret_cast<short *>(pIdx + a.spokes);           // [spokes]
    short *pHi = pLo + a.spokes;
    unsigned char *keep = reinterpret_cast<unsigned char *>(pHi + a.spokes);
    unsigned char *pNode = keep + a.spokes;                            // [levels][spokes]

    const int row = a.rowBegin + static_cast<int>(blockIdx.x);
    const int slot = row / a.rings, ring = row - slot * a.rings;
    const int haloIdx = a.batchHalos[slot];
    const HaloDev &h = a.halos[haloIdx];
    const int tid = threadIdx.x;
    const size_t ringBase = static_cast<size_t>(row) * a.spokes;
    const double *rsp = a.rspBySlot ? a.rsp + ringBase : a.rsp + (static_cast<size_t>(haloIdx) * a.rings + ring) * a.spokes;

    // valid points in LoS order (ring_buffer.go OkPolarCoords): staged through shared memory so
    // that the ordered compaction does not wait on one global load per LoS (n <= i: in place)
    for (int i = tid; i < a.spokes; i += kFilterThreads) pR[i] = rsp[i];
    __syncthreads();
    if (tid == 0) {
        int n = 0;
        double high = 0;
        for (int i = 0; i < a.spokes; i++) {
            const double r = pR[i];
            if (!isnan(r)) {
                pR[n] = r; pIdx[n] = i;
                if (n == 0 || r > high) high = r;
                n++;
            }
        }
        sm.nValid = n; sm.high = high; sm.status = 0; sm.kept = 0;
        if (n == 0) sm.status = SFK_HALO_NO_POINTS;
    }
    __syncthreads();
    const int nv = sm.nValid;
    for (int p = tid; p < nv; p += kFilterThreads)
        for (int l = 1; l <= a.levels; l++)
            pNode[(l - 1) * a.spokes + p] =
                static_cast<unsigned char>(go_int(a.ringPhi[pIdx[p]] / (kTwoPi / static_cast<double>(1 << l))));
    if (sm.status) {
        if (tid == 0) { atomicMax(&a.status[haloIdx], sm.status); a.fCount[slot * a.rings + ring] = 0; }
        return;
    }
    const double low = 0.0, high = sm.high;
    const double h0 = h.rMax / a.eta;
    double factor = 1.0;
    for (int attempt = 0; attempt < 10; attempt++) {
        const double hk = h0 * factor;
        // NewKDETree: knots and evaluation radii, kde.go:83-99, 14-22
        const double dxk = (high - low) / static_cast<double>(kKdeKnots - 1);
        const double drk = (high - low) / static_cast<double>(kKdeKnots);
        for (int i = tid; i < kKdeKnots; i += kFilterThreads) {
            sm.X[i] = i < kKdeKnots - 1 ? low + dxk * static_cast<double>(i) : high;
            sm.spRs[i] = i < kKdeKnots - 1 ? low + drk * static_cast<double>(i) : high;
        }
        const double maxDist = hk * 3;
        for (int p = tid; p < nv; p += kFilterThreads) {
            long long lo = go_int((pR[p] - maxDist - low) / dxk);
            long long hi = go_int((pR[p] + maxDist - low) / dxk) + 1;
            if (lo < 0) lo = 0;
            if (hi >= kKdeKnots) hi = kKdeKnots - 1;
            pLo[p] = static_cast<short>(lo);
            pHi[p] = static_cast<short>(hi);
        }
        __syncthreads();
        // GaussianKDE for every node of every level, kde.go:24-38.  One thread per
        // knot walks the points in order; exp() is evaluated once per (point, knot)
        // and added to the point's node at every level (binByTheta, kde.go:156-174),
        // so each KDE still sums its own points in the reference's order.
        for (int t = tid; t < nKde * kKdeKnots; t += kFilterThreads) Y[t] = 0.0;
        __syncthreads();
        // The points are in LoS order, so a point's wedge at the finest level is non-decreasing:
        // every finest-level KDE is the sum over one contiguous run of points and lives in a
        // register until the wedge changes (which depends on p only: no divergence).  A coarser
        // KDE is the sum of its two children -- the reference sums each KDE over its own points in
        // order (binByTheta, kde.go:156-174); regrouping moves a sum by an ulp, like the choice of
        // exp() does, and decides nothing beyond tie level.
        const double invHk = 1.0 / hk;
        const unsigned char *fine = pNode + (a.levels - 1) * a.spokes;
        for (int i = tid; i < kKdeKnots; i += kFilterThreads) {
            const double xi = sm.X[i];
            double acc = 0.0;
            int node = 0;
            double *yFine = Y + ((1 << a.levels) - 1) * kKdeKnots + i;
            for (int p = 0; p < nv; p++) {
                const int nd = fine[p];
                if (nd != node) {
                    yFine[node * kKdeKnots] = acc;
                    acc = 0.0;
                    node = nd;
                }
                if (i < pLo[p] || i > pHi[p]) continue;
                const double udx = (xi - pR[p]) * invHk;
                acc += exp(-udx * udx);
            }
            yFine[node * kKdeKnots] = acc;
            for (int l = a.levels - 1; l >= 0; l--) {
                double *par = Y + ((1 << l) - 1) * kKdeKnots + i;
                const double *ch = Y + ((2 << l) - 1) * kKdeKnots + i;
                for (int n = 0; n < (1 << l); n++) par[n * kKdeKnots] = ch[2 * n * kKdeKnots] + ch[(2 * n + 1) * kKdeKnots];
            }
        }
        __syncthreads();
        for (int k = tid; k < nKde; k += kFilterThreads)
            if (!spline_y2(sm.X, Y + k * kKdeKnots, kKdeKnots, Y2 + k * kKdeKnots, V + k * kKdeKnots))
                sm.status = SFK_HALO_SPLINE;
        __syncthreads();
        // localSplineMaxes, kde.go:193-206
        for (int t = tid; t < nKde * kKdeKnots; t += kFilterThreads) {
            const int kde = t / kKdeKnots, i = t - kde * kKdeKnots;
            double v = 0;
            if (!spline_eval(sm.X, Y + kde * kKdeKnots, Y2 + kde * kKdeKnots, kKdeKnots, sm.spRs[i], v))
                sm.status = SFK_HALO_SPLINE;
            V[t] = v;
        }
        __syncthreads();
        for (int k = tid; k < nKde; k += kFilterThreads) {
            const double *v = V + k * kKdeKnots;
            int m = 0;
            for (int i = 1; i < kKdeKnots - 1; i++)
                if (v[i] > v[i + 1] && v[i] > v[i - 1]) maxes[k * kMaxMax + m++] = sm.spRs[i];
            sm.nMax[k] = m;
        }
        __syncthreads();
        // connectMaxes, kde.go:211-258 (sequential tree walk)
        if (tid == 0 && sm.status == 0) {
            if (sm.nMax[0] == 0) {
                sm.status = SFK_HALO_NO_MAXIMUM;
            } else {
                sm.conn[0][0] = maxes[0];
                for (int split = 0; split < a.levels && sm.status == 0; split++) {
                    const int l = split + 1, nodes = 1 << l;
                    for (int node = 0; node < nodes; node++) {
                        const int kde = (1 << l) - 1 + node;
                        const double *nm = maxes + kde * kMaxMax;
                        const int cnt = sm.nMax[kde];
                        const double prev = sm.conn[split][node / 2];
                        const double nanv = __longlong_as_double(0x7ff8000000000000ll);
                        double connMax = nanv, cur = nanv;
                        if (!isnan(prev) && cnt > 0) {
                            int ci = -1;
                            double best = INFINITY;
                            for (int i = 0; i < cnt; i++) {
                                const double d = fabs(prev - nm[i]);
                                if (d < best) { ci = i; best = d; }
                            }
                            connMax = nm[ci];
                        }
                        if (!isnan(connMax) && (split == 0 || fabs(connMax - prev) < hk)) {
                            cur = connMax;
                        } else if (cnt > 0) {
                            if (!build_rfunc(sm, split)) { sm.status = SFK_HALO_SPLINE; break; }
                            const double thc = kTwoPi * (static_cast<double>(node) + 0.5) / static_cast<double>(nodes);
                            double spR;
                            if (!spline_eval(sm.th, sm.mx, sm.ay2, sm.anchorN, thc, spR)) {
                                sm.status = SFK_HALO_SPLINE;
                                break;
                            }
                            for (int i = 0; i < cnt; i++)
                                if (fabs(nm[i] - spR) < hk) cur = nm[i];
                        }
                        sm.conn[l][node] = cur;
                    }
                }
                // FilterNearby's curve: GetRFunc(levels, Radial), kde.go:355-369
                if (sm.status == 0 && !build_rfunc(sm, a.levels)) sm.status = SFK_HALO_SPLINE;
            }
        }
        __syncthreads();
        if (sm.status) break;
        for (int p = tid; p < nv; p += kFilterThreads) {
            double v;
            if (!spline_eval(sm.th, sm.mx, sm.ay2, sm.anchorN, a.ringPhi[pIdx[p]], v)) {
                sm.status = SFK_HALO_SPLINE;
                keep[p] = 0;
            } else {
                keep[p] = fabs(v - pR[p]) < hk / 2 ? 1 : 0;
            }
        }
        __syncthreads();
        if (tid == 0) {
            int m = 0;
            for (int p = 0; p < nv; p++)
                if (keep[p]) { a.fR[ringBase + m] = pR[p]; a.fIdx[ringBase + m] = pIdx[p]; m++; }
            sm.kept = m;
        }
        __syncthreads();
        if (sm.kept > 0 || sm.status) break;
        factor *= 1.1;
    }
    if (tid == 0) {
        if (sm.status) atomicMax(&a.status[haloIdx], sm.status);
        a.fCount[slot * a.rings + ring] = sm.status ? 0 : sm.kept;
    }
}

// ------------------------------------------------------------------ fit --

// gonum Dgetf2 + Dtrti2 + Dgetri (unblocked, row-major) == LAPACK
// DGETF2/DTRTI2/DGETRI: the inverse mat64.Dense.Inverse computes in pinv,
// penna.go:170-186.  One thread; a is n x n in shared memory.
__device__ bool invert_lu(double *a, int n, int *ipiv, double *work) {
    for (int j = 0; j < n; j++) {
        int jp = j;
        double amax = fabs(a[j * n + j]);
        for (int i = j + 1; i < n; i++) {
            const double v = fabs(a[i * n + j]);
            if (v > amax) { amax = v; jp = i; }
        }
        ipiv[j] = jp;
        if (a[jp * n + j] != 0) {
            if (jp != j)
                for (int k = 0; k < n; k++) {
                    const double t = a[j * n + k]; a[j * n + k] = a[jp * n + k]; a[jp * n + k] = t;
                }
            if (j < n - 1) {
                const double inv = 1 / a[j * n + j];
                for (int i = j + 1; i < n; i++) a[i * n + j] *= inv;
            }
        }
        if (j < n - 1)
            for (int i = j + 1; i < n; i++) {
                const double tmp = -1 * a[i * n + j];
                if (tmp == 0) continue;
                for (int k = j + 1; k < n; k++) a[i * n + k] += tmp * a[j * n + k];
            }
    }
    for (int i = 0; i < n; i++) if (a[i * n + i] == 0) return false;
    for (int j = 0; j < n; j++) {
        double ajj = 1 / a[j * n + j];
        a[j * n + j] = ajj;
        ajj *= -1;
        for (int i = 0; i < j; i++) {
            double tmp = 0;
            for (int k = i; k < j; k++) tmp += a[i * n + k] * a[k * n + j];
            a[i * n + j] = tmp;
        }
        for (int i = 0; i < j; i++) a[i * n + j] *= ajj;
    }
    for (int j = n - 1; j >= 0; j--) {
        for (int i = j + 1; i < n; i++) { work[i] = a[i * n + j]; a[i * n + j] = 0; }
        if (j < n - 1)
            for (int i = 0; i < n; i++) {
                double tmp = 0;
                for (int k = j + 1; k < n; k++) tmp += a[i * n + k] * work[k];
                a[i * n + j] += -1 * tmp;
            }
    }
    for (int j = n - 2; j >= 0; j--) {
        const int jp = ipiv[j];
        if (jp != j)
            for (int i = 0; i < n; i++) {
                const double t = a[i * n + j]; a[i * n + j] = a[i * n + jp]; a[i * n + jp] = t;
            }
    }
    return true;
}

constexpr int kFitThreads = 256;
constexpr int kFitTile = 256;   // points per shared-memory tile
constexpr int kFitChunks = 8;   // CTAs per halo
constexpr int kMaxP2 = 2 * kMaxOrder * kMaxOrder;

// Design-matrix row of one filtered point, penna.go:25-52: the point is lifted
// to the box frame through irot in float32 (PlaneToVolume, halo.go:471-475).
// math.Pow is called per entry in the reference; the distinct (base, exponent)
// pairs are evaluated once here, same values.
template <int P>
__device__ __forceinline__ double penna_row(const FitArgs &a, int slot, int ring, int k, double *row) {
    const size_t src = (static_cast<size_t>(slot) * a.rings + ring) * a.spokes + k;
    const double R = a.fR[src];
    const int li = a.fIdx[src];
    // fXs, fYs (penna.go:157-160)
    const double px = R * a.ringCos[li], py = R * a.ringSin[li];
    float fx, fy, fz;
    rotate_vec(a.irots + 9 * ring, static_cast<float>(px), static_cast<float>(py), 0.f, fx, fy, fz);
    const double x = fx, y = fy, z = fz;
    const double r = sqrt(x * x + y * y + z * z);
    const double costh = z / r;
    const double sinth = sqrt(1 - costh * costh);
    const double cosphi = x / r / sinth;
    const double sinphi = y / r / sinth;
    double sinPow[2 * P - 1], sphiPow[P];
#pragma unroll
    for (int e = 0; e < 2 * P - 1; e++) sinPow[e] = go_pow_int(sinth, e);
#pragma unroll
    for (int j = 0; j < P; j++) sphiPow[j] = go_pow_int(sinphi, j);
    int m = 0;
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
        const double ck = go_pow_int(costh, kk);
#pragma unroll
        for (int j = 0; j < P; j++) {
            double cphi = 1.0;
#pragma unroll
            for (int i = 0; i < P; i++) {
                row[m] = sinPow[i + j] * cphi * ck * sphiPow[j];
                m++;
                cphi *= cosphi;
            }
        }
    }
    return r;
}

// PennaVolumeFit, penna.go:88-108 -> PennaCoeffs :14-58, in four small kernels so that a
// halo's <= rings*spokes points are spread over kFitChunks CTAs instead of one:
//   fit_rows_kernel    design rows -> scratch; partial Gram matrices M M^T per chunk
//   fit_solve_kernel   sum of the partials (chunk order), gonum's LU inverse
//   fit_apply_kernel   partial rs^T (M^T inv) per chunk
//   fit_finish_kernel  sum of the partials (chunk order) -> coefficients
// Inside a chunk every accumulator adds its terms in ascending point order, like the
// reference's loops (and gonum's Dgemm); chunk partials are added in chunk order, which
// regroups the sum but keeps it deterministic.
__device__ __forceinline__ int fit_point_count(const FitArgs &a, int slot, int *ringOff) {
    // ringOff[r] = number of points in rings < r (shared memory, rings + 1 entries)
    if (threadIdx.x == 0) {
        int s = 0;
        for (int r = 0; r < a.rings; r++) { ringOff[r] = s; s += a.fCount[slot * a.rings + r]; }
        ringOff[a.rings] = s;
    }
    __syncthreads();
    return ringOff[a.rings];
}

__device__ __forceinline__ void fit_chunk_range(int N, int chunk, int &n0, int &n1) {
    const int tiles = (N + kFitTile - 1) / kFitTile;
    const int per = (tiles + kFitChunks - 1) / kFitChunks * kFitTile;
    n0 = min(N, chunk * per);
    n1 = min(N, n0 + per);
}

template <int P>
__global__ void __launch_bounds__(kFitThreads) fit_rows_kernel(FitArgs a) {
    constexpr int P2 = 2 * P * P;
    constexpr int NE = P2 * (P2 + 1) / 2;                       // upper triangle incl. diagonal
    constexpr int EPT = (NE + kFitThreads - 1) / kFitThreads;   // entries per thread
    extern __shared__ double fsm[];
    __shared__ int ringOff[kMaxRings + 1];
    double *tileM = fsm;   // [kFitTile][P2 + 1]: padded rows, conflict-free column walks
    const int chunk = blockIdx.x, slot = blockIdx.y, tid = threadIdx.x;
    const int haloIdx = a.batchHalos[slot];
    const int N = fit_point_count(a, slot, ringOff);
    if (chunk == 0 && tid == 0) a.nPoints[haloIdx] = N;
    int c0, c1;
    fit_chunk_range(N, chunk, c0, c1);
    // entry e of the upper triangle -> (gi, gj), gi <= gj
    int gi[EPT], gj[EPT];
    double acc[EPT];
#pragma unroll
    for (int q = 0; q < EPT; q++) {
        int e = tid + q * kFitThreads, i = 0;
        if (e >= NE) e = 0;
        while (e >= P2 - i) { e -= P2 - i; i++; }
        gi[q] = i; gj[q] = i + e; acc[q] = 0.0;
    }
    const size_t rowBase = static_cast<size_t>(slot) * a.rings * a.spokes;
    for (int n0 = c0; n0 < c1; n0 += kFitTile) {
        __syncthreads();
        const int n = n0 + tid;
        if (n < c1) {
            int lo = 0, hi = a.rings;   // ring with ringOff[ring] <= n < ringOff[ring + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) / 2;
                if (ringOff[mid] <= n) lo = mid; else hi = mid;
            }
            double row[P2];
            const double r = penna_row<P>(a, slot, lo, n - ringOff[lo], row);
            double *dst = a.rows + (rowBase + n) * (P2 + 1);
#pragma unroll
            for (int j = 0; j < P2; j++) { tileM[tid * (P2 + 1) + j] = row[j]; dst[j] = row[j]; }
            dst[P2] = r;
        }
        __syncthreads();
        const int cnt = min(kFitTile, c1 - n0);
#pragma unroll
        for (int q = 0; q < EPT; q++) {
            if (tid + q * kFitThreads < NE) {
                const double *pi = tileM + gi[q], *pj = tileM + gj[q];
                double s = acc[q];
#pragma unroll 8
                for (int k = 0; k < cnt; k++) s += pi[k * (P2 + 1)] * pj[k * (P2 + 1)];
                acc[q] = s;
            }
        }
    }
    double *out = a.gramPart + (static_cast<size_t>(slot) * kFitChunks + chunk) * (P2 * P2);
#pragma unroll
    for (int q = 0; q < EPT; q++)
        if (tid + q * kFitThreads < NE) {
            out[gi[q] * P2 + gj[q]] = acc[q];
            out[gj[q] * P2 + gi[q]] = acc[q];   // a*b == b*a: Dense.Mul's result is bitwise symmetric
        }
}

__global__ void __launch_bounds__(kMaxP2 * kMaxP2) fit_solve_kernel(FitArgs a) {
    __shared__ double G[kMaxP2 * kMaxP2];
    __shared__ double work[kMaxP2];
    __shared__ int ipiv[kMaxP2];
    __shared__ int ringOff[kMaxRings + 1];
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int haloIdx = a.batchHalos[slot];
    const int P2 = 2 * a.order * a.order;
    const int N = fit_point_count(a, slot, ringOff);
    if (a.status[haloIdx] != 0) return;   // a ring already failed
    if (N == 0) {
        if (tid == 0) a.status[haloIdx] = SFK_HALO_NO_POINTS;
        return;
    }
    if (tid < P2 * P2) {
        double s = 0.0;
        for (int c = 0; c < kFitChunks; c++)
            s += a.gramPart[(static_cast<size_t>(slot) * kFitChunks + c) * (P2 * P2) + tid];
        G[tid] = s;
    }
    __syncthreads();
    if (tid == 0 && !invert_lu(G, P2, ipiv, work)) a.status[haloIdx] = SFK_HALO_SINGULAR;
    __syncthreads();
    if (tid < P2 * P2) a.gramInv[static_cast<size_t>(slot) * (P2 * P2) + tid] = G[tid];
}

template <int P>
__global__ void __launch_bounds__(kFitThreads) fit_apply_kernel(FitArgs a) {
    constexpr int P2 = 2 * P * P;
    extern __shared__ double fsm[];
    __shared__ int ringOff[kMaxRings + 1];
    double *tileM = fsm;                              // [kFitTile][P2 + 1], last column = r
    double *inv = tileM + kFitTile * (P2 + 1);        // [P2][P2]
    const int chunk = blockIdx.x, slot = blockIdx.y, tid = threadIdx.x;
    const int haloIdx = a.batchHalos[slot];
    const int N = fit_point_count(a, slot, ringOff);
    double *out = a.csPart + (static_cast<size_t>(slot) * kFitChunks + chunk) * P2;
    if (a.status[haloIdx] != 0 || N == 0) return;
    for (int k = tid; k < P2 * P2; k += kFitThreads) inv[k] = a.gramInv[static_cast<size_t>(slot) * (P2 * P2) + k];
    int c0, c1;
    fit_chunk_range(N, chunk, c0, c1);
    const size_t rowBase = static_cast<size_t>(slot) * a.rings * a.spokes;
    // cs = rs^T (M^T inv) (mat.VecMult, mat.go:102): thread j sums over points in order
    double sum = 0.0;
    for (int n0 = c0; n0 < c1; n0 += kFitTile) {
        __syncthreads();
        const int n = n0 + tid;
        if (n < c1) {
            const double *src = a.rows + (rowBase + n) * (P2 + 1);
            // row of out2 = M^T inv (gonum Dense.Mul: axpy over l ascending)
            double o2[P2];
#pragma unroll
            for (int j = 0; j < P2; j++) o2[j] = 0.0;
#pragma unroll 1
            for (int l = 0; l < P2; l++) {
                const double t = src[l];
                if (t == 0) continue;
#pragma unroll
                for (int j = 0; j < P2; j++) o2[j] += t * inv[l * P2 + j];
            }
#pragma unroll
            for (int j = 0; j < P2; j++) tileM[tid * (P2 + 1) + j] = o2[j];
            tileM[tid * (P2 + 1) + P2] = src[P2];
        }
        __syncthreads();
        if (tid < P2) {
            const int cnt = min(kFitTile, c1 - n0);
#pragma unroll 8
            for (int k = 0; k < cnt; k++) sum += tileM[k * (P2 + 1) + P2] * tileM[k * (P2 + 1) + tid];
        }
    }
    if (tid < P2) out[tid] = sum;
}

__global__ void fit_finish_kernel(FitArgs a) {
    const int slot = blockIdx.x, tid = threadIdx.x;
    const int haloIdx = a.batchHalos[slot];
    const int P2 = 2 * a.order * a.order;
    if (tid >= P2 || a.status[haloIdx] != 0) return;
    double s = 0.0;
    for (int c = 0; c < kFitChunks; c++) s += a.csPart[(static_cast<size_t>(slot) * kFitChunks + c) * P2 + tid];
    a.coeffs[static_cast<size_t>(haloIdx) * P2 + tid] = s;
}

}  // namespace

// --------------------------------------------------------------- launch --

cudaError_t launch_cull(const CullArgs &a, cudaStream_t s) {
    const int64_t live = (a.n + a.sf3 - 1) / a.sf3;
    if (live <= 0 || a.nList <= 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((live + kCullThreads - 1) / kCullThreads);
    cull_kernel<<<grid, kCullThreads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_bucket(const Cand *cands, int64_t n, const int64_t *offset, uint32_t *cursor, Cand *sorted,
                          cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    bucket_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(cands, n, offset, cursor, sorted);
    return cudaGetLastError();
}

cudaError_t launch_hit_count(const Piece *pieces, int nPieces, const Cand *cands, const HaloDev *halos,
                             const float *norms, int rings, uint32_t *hitCount,
                             unsigned long long *rhoMaxBits, cudaStream_t s) {
    if (nPieces <= 0) return cudaSuccess;
    const size_t smem = sizeof(float) * 3 * rings + sizeof(uint32_t) * rings;
    hit_count_kernel<<<nPieces, kHitThreads, smem, s>>>(pieces, cands, halos, norms, rings, hitCount, rhoMaxBits);
    return cudaGetLastError();
}

cudaError_t launch_hits(const Piece *pieces, int nPieces, const Cand *cands,
                        const HaloDev *halos, const float *norms, const float *rots,
                        const double *ringCos, const double *ringSin,
                        int rings, int spokes, double dPhi, const int64_t *hitOffset,
                        uint32_t *hitCursor, const int32_t *haloSlot, HitRec *recs, bool prune, cudaStream_t s) {
    if (nPieces <= 0) return cudaSuccess;
    const size_t smem = sizeof(Cand) * kHitThreads + sizeof(float) * 3 * rings +
                        sizeof(uint32_t) * kHitThreads * kHitRingChunk;
    {   // per device, so set it on every launch (a second GPU in the same process must get it too)
        cudaError_t e = prune ? cudaFuncSetAttribute(hits_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)
                              : cudaFuncSetAttribute(hits_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    if (prune)
        hits_kernel<true><<<nPieces, kHitThreads, smem, s>>>(pieces, cands, halos, norms, rots, ringCos, ringSin, rings,
                                                             spokes, dPhi, hitOffset, hitCursor, haloSlot, recs);
    else
        hits_kernel<false><<<nPieces, kHitThreads, smem, s>>>(pieces, cands, halos, norms, rots, ringCos, ringSin, rings,
                                                              spokes, dPhi, hitOffset, hitCursor, haloSlot, recs);
    return cudaGetLastError();
}


// ---- PercentileProfile mode ---------------------------------------------------
// calcPercentile (cmd/shell.go:629-660): for every radial bin, the n-th largest
// GetRhos value over all rings*spokes lines of sight.  GetRhos is
// max(prefix * quantum, defaultRho) (halo.go:350-357), monotone in the int64
// prefix sum, so the selection runs on the exact integers and the clamp is
// applied to the selected one: same value as msort.NthLargest on the float64s.

// ProfileRing.Retrieve (profile_ring.go:124-130) in place: one warp per LoS.
__global__ void __launch_bounds__(256) prefix_kernel(unsigned long long *grid, int64_t nRows, int bins) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= nRows) return;
    long long *g = reinterpret_cast<long long *>(grid) + row * bins;
    long long carry = 0;
    for (int base = 0; base < bins; base += 32) {
        const int k = base + lane;
        long long v = k < bins ? g[k] : 0;
        for (int o = 1; o < 32; o <<= 1) {
            const long long up = __shfl_up_sync(0xffffffffu, v, o);
            if (lane >= o) v += up;
        }
        v += carry;
        if (k < bins) g[k] = v;
        carry = __shfl_sync(0xffffffffu, v, 31);
    }
}

// One CTA per (radial bin, halo): most-significant-byte-first radix select of
// the nth largest key, 8 passes over the bin's column of the prefix grid.
__global__ void __launch_bounds__(256) percentile_kernel(PercentileArgs a) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long sPrefix;
    __shared__ long long sNth;
    const int bin = blockIdx.x, slot = blockIdx.y;
    const int haloIdx = a.batchHalos[slot];
    const HaloDev &h = a.halos[haloIdx];
    const long long *col = reinterpret_cast<const long long *>(a.grid) +
                           static_cast<size_t>(slot) * a.nLos * a.bins + bin;
    if (threadIdx.x == 0) { sPrefix = 0; sNth = a.nth; }
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        hist[threadIdx.x] = 0;
        __syncthreads();
        const unsigned long long prefix = sPrefix;
        const unsigned long long mask = pass == 0 ? 0ull : ~0ull << (shift + 8);
        for (int64_t i = threadIdx.x; i < a.nLos; i += blockDim.x) {
            // order-preserving map int64 -> uint64
            const unsigned long long key = static_cast<unsigned long long>(col[i * a.bins]) ^ 0x8000000000000000ull;
            if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            long long n = sNth;
            int b = 255;
            for (; b > 0; b--) {
                if (static_cast<long long>(hist[b]) >= n) break;
                n -= hist[b];
            }
            sNth = n;
            sPrefix = prefix | (static_cast<unsigned long long>(b) << shift);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long acc = static_cast<long long>(sPrefix ^ 0x8000000000000000ull);
        double rho = static_cast<double>(acc) * h.quantum;
        if (rho < h.defaultRho) rho = h.defaultRho;   // GetRhos clamp, halo.go:353-355
        a.rows[static_cast<size_t>(haloIdx) * 2 * a.bins + a.bins + bin] = rho;
    }
}

cudaError_t launch_percentile(const PercentileArgs &a, int nb, cudaStream_t s) {
    if (nb <= 0) return cudaSuccess;
    const int64_t rows = static_cast<int64_t>(nb) * a.nLos;
    prefix_kernel<<<static_cast<unsigned>((rows * 32 + 255) / 256), 256, 0, s>>>(a.grid, rows, a.bins);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    percentile_kernel<<<dim3(a.bins, nb), 256, 0, s>>>(a);
    return cudaGetLastError();
}

// Sliding-moment form of the fast path; false when the shape is not served.
// (windows below 29 taps would need runs shorter than a lane to be recognised: tap loop)
bool los_moments_supported(const LosArgs &a) { return a.bins == 256 && a.window <= kLosFastMaxWindow && a.window >= 29; }

cudaError_t launch_los(const LosArgs &a, const SgPoly *sg, int nb, cudaStream_t s) {
    if (nb <= 0) return cudaSuccess;
    const int64_t warps = a.losCount;
    if (warps <= 0) return cudaSuccess;
    const unsigned grid = static_cast<unsigned>((warps + kLosWarps - 1) / kLosWarps);
    if (sg && los_moments_supported(a)) {
        const size_t fsmem = sizeof(double) * (4 * (a.window / 2 + 1) + 256 + kLosWarps * 8 * kLosFastRow);
        cudaError_t e = cudaMemsetAsync(a.flatCount, 0, sizeof(unsigned long long), s);
        if (e != cudaSuccess) return e;
        LosArgs m = a;
        m.flatLanes = std::max(1, std::min(2, (a.window / 2 - 6) / 8));
        los_kernel_fast<8, true><<<grid, kLosWarps * 32, fsmem, s>>>(m, nb, *sg);
        // the lines with exact-tie hazards again, literally; warps beyond the list's length exit at once
        LosArgs b = a;
        b.listMode = 1;
        const size_t tsmem = sizeof(double) * (2 * a.window + kLosWarps * 8 * kLosFastRow);
        los_kernel_fast<8, false><<<grid, kLosWarps * 32, tsmem, s>>>(b, nb, SgPoly{});
    } else if (a.bins == 256 && a.window <= kLosFastMaxWindow) {
        const size_t fsmem = sizeof(double) * (2 * a.window + kLosWarps * 8 * kLosFastRow);
        los_kernel_fast<8, false><<<grid, kLosWarps * 32, fsmem, s>>>(a, nb, SgPoly{});
    } else {
        const size_t smem = sizeof(double) * (2 * a.window + kLosWarps * 3 * a.bins);
        // per device, so set it on every launch (a second GPU in the same process must get it too)
        cudaError_t e = cudaFuncSetAttribute(los_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        los_kernel<<<grid, kLosWarps * 32, smem, s>>>(a, nb);
    }
    return cudaGetLastError();
}

cudaError_t launch_filter(const FilterArgs &a, int nb, cudaStream_t s) {
    if (nb <= 0) return cudaSuccess;
    const int nKde = (1 << (a.levels + 1)) - 1;
    const size_t smem = sizeof(FilterSmem) + sizeof(double) * (3 * nKde * kKdeKnots + nKde * kMaxMax + a.spokes) +
                        sizeof(int) * a.spokes + 2 * sizeof(short) * a.spokes + a.spokes * (1 + a.levels) + 16;
    {   // per device, so set it on every launch (a second GPU in the same process must get it too)
        cudaError_t e = cudaFuncSetAttribute(filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    if (a.rowCount <= 0) return cudaSuccess;
    filter_kernel<<<a.rowCount, kFilterThreads, smem, s>>>(a);
    return cudaGetLastError();
}

template <int P>
cudaError_t launch_fit_p(const FitArgs &a, int nb, cudaStream_t s) {
    constexpr int P2 = 2 * P * P;
    const size_t smemA = sizeof(double) * kFitTile * (P2 + 1);
    const size_t smemC = smemA + sizeof(double) * P2 * P2;
    {   // per device, so set it on every launch (a second GPU in the same process must get it too)
        cudaError_t e = cudaFuncSetAttribute(fit_rows_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(fit_apply_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
        if (e != cudaSuccess) return e;
    }
    fit_rows_kernel<P><<<dim3(kFitChunks, nb), kFitThreads, smemA, s>>>(a);
    fit_solve_kernel<<<nb, kMaxP2 * kMaxP2, 0, s>>>(a);
    fit_apply_kernel<P><<<dim3(kFitChunks, nb), kFitThreads, smemC, s>>>(a);
    fit_finish_kernel<<<nb, 32, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_fit(const FitArgs &a, int nb, cudaStream_t s) {
    if (nb <= 0) return cudaSuccess;
    switch (a.order) {
    case 1: return launch_fit_p<1>(a, nb, s);
    case 2: return launch_fit_p<2>(a, nb, s);
    case 3: return launch_fit_p<3>(a, nb, s);
    case 4: return launch_fit_p<4>(a, nb, s);
    default: return cudaErrorInvalidValue;
    }
}

size_t fit_scratch_doubles(int nb, int rings, int spokes, int order, size_t *rows, size_t *gramPart, size_t *gramInv,
                           size_t *csPart) {
    const size_t P2 = 2 * static_cast<size_t>(order) * order;
    *rows = static_cast<size_t>(nb) * rings * spokes * (P2 + 1);
    *gramPart = static_cast<size_t>(nb) * kFitChunks * P2 * P2;
    *gramInv = static_cast<size_t>(nb) * P2 * P2;
    *csPart = static_cast<size_t>(nb) * kFitChunks * P2;
    return *rows + *gramPart + *gramInv + *csPart;
}

}  // namespace sfk
