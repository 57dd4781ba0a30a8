// sfk_bin.cu -- the binning kernel: per-LoS chords of every (particle, ring)
// hit and ProfileRing.Insert (los/halo.go:241-277, profile_ring.go:92-120).
//
// One CTA per (LoS tile, ring, halo[, chunk of the hit list]); a tile is 64 lines of sight in
// the product instantiation (32 warps, one CTA per SM) and 32 otherwise (16 warps, two CTAs).
// The tile's TILE x bins accumulators are 64-bit fixed point, split into lo/hi
// 32-bit words in shared memory so that native 32-bit shared atomics can be
// used (64-bit shared atomics are CAS loops on sm_100).
//
// Every warp does the same job.  It draws 128-record chunks of the ring's hit
// list from a shared counter (single groups of 32 near the end of the list, so that the warps
// finish together), reads only the LoS ranges of those records
// (prefetched one chunk ahead), and copies the records that touch this tile --
// one in three or four -- into its private 64-slot ring in shared memory with
// cp.async.  Whenever 32 records are waiting, their ranges are clipped to the
// tile, a warp scan turns the clipped lengths into a list of (record, LoS)
// pairs, and the pairs are processed 32 at a time with one pair per lane (a
// lane finds its record from one REDUX.OR over the records' first-pair indices,
// a POPC and a running count), so every lane does useful work whatever the range widths.
// A CTA that owns its rows alone stores the finished tile into the grid
// (no memset, no read-modify-write); otherwise it adds it with RED.64.
// hits_kernel fills each list from both ends by the branch the records' chords take at the
// start edge of ProfileRing.Insert, so a warp mostly runs one of the two.
//
// Arithmetic.  Everything that DECIDES something is evaluated literally, in
// the reference's operation order with no FMA contraction (-fmad=false): the
// impact parameter b, b^2, the two radicands and their signs (NaN semantics of
// twoValIntrDist / oneValIntrDist), and all comparisons.  ln() and the
// division by dr of profile_ring.go:100 are replaced by a table of bin edges
// g[k] = exp(k*dr): bin = k with g[k] <= r/rMin < g[k+1], rem = ln(r/(rMin
// g[k]))/dr from a degree-6 log1p series (|t| < 0.0092, error < 1e-14).  This
// differs from floor((ln r - lowR)/dr) only when r lies within a few ulp of a
// bin edge, the same class of tie on which Go's and glibc's ln already differ.
#include <cstdint>
#include <cuda_runtime.h>

#include "sfk_kernels.cuh"
#include "sfk_device.cuh"

namespace sfk {

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// 16-byte asynchronous copy global -> shared (LDGSTS, L2 only: a CTA reads a record once): the data never visits registers
__device__ __forceinline__ void cp_async16(void *dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }


// ---- shared-memory accesses of the pair loop through 32-bit shared-window addresses.  The two
// bases (this round's records, the CTA's tables) are made opaque once per round, so they stay in
// two registers; as pointers the compiler re-derives them from the thread and CTA ids on every
// pass of the loop (7 instructions of ~170) under the kernel's 64-register cap.
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    asm volatile("" : "+r"(v));
    return v;
}
// v << n for n < 32, 0 otherwise
__device__ __forceinline__ unsigned shl_clamp(unsigned v, unsigned n) {
    unsigned r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(n));
    return r;
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t atoms_add(uint32_t addr, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void reds_add(uint32_t addr, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// Bin index and fractional position of radius r, rMin < r < rMax
// (profile_ring.go:100-101: Modf((ln r - lowR)/dr)).
__device__ __forceinline__ void bin_of(double r, double invRMin, float cK, double invDr60,
                                       uint32_t invGA, int &idx, double &rem) {
    const double u = r * invRMin;
    // float32 view of u (u in [1, rMax/rMin]): rebias the exponent, keep 20 mantissa bits.
    // The estimate is biased low by 1e-3 bins (its error is < 3e-4), so the true
    // bin is k or k + 1 and the series argument t is never negative.
    const float uf = __int_as_float((__double2hiint(u) - 0x38000000) << 3);
    float lg;   // uf is a normal number >= 1: the flush-to-zero form skips __log2f's denormal path
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(uf));
    // callers guarantee rMin < r < rMax, and the estimate never exceeds the true position:
    // 0 <= k <= bins - 1 without clamping
    int k = static_cast<int>(fmaf(lg, cK, -1.0e-3f));
    const double t = fma(u, lds_f64(invGA + 8u * k), -1.0);   // r / (rMin g[k]) - 1
    // 60 log1p(t)/t = 60 - 30 t + 20 t^2 - 15 t^3 + 12 t^4 - 10 t^5: every coefficient is an
    // immediate operand of its DFMA (1/3, 1/5, 1/6 would each cost a constant load per use under
    // the kernel's 64-register cap); invDr60 = 1/(60 dr)
    double L = fma(t, -10.0, 12.0);
    L = fma(L, t, -15.0);
    L = fma(L, t, 20.0);
    L = fma(L, t, -30.0);
    L = fma(L, t, 60.0);
    double x = (L * t) * invDr60;
    // if (x >= 1.0) { k += 1; x -= 1.0; } as two predicated instructions
    asm("{\n\t.reg .pred p;\n\tsetp.ge.f64 p, %0, 0d3FF0000000000000;\n\t@p add.f64 %0, %0, 0dBFF0000000000000;\n\t"
        "@p add.s32 %1, %1, 1;\n\t}" : "+d"(x), "+r"(k));
    idx = k;
    rem = x;
}

// 64-bit two's-complement add into split lo/hi words; row = shared address of the low word of
// (bin 0, this LoS), hiOff = byte distance from the low words to the high ones.
template <int TILE>
__device__ __forceinline__ void acc_add(uint32_t row, uint32_t hiOff, int idx, long long v) {
    const uint32_t vlo = static_cast<uint32_t>(v);
    const uint32_t vhi = static_cast<uint32_t>(static_cast<unsigned long long>(v) >> 32);
    const uint32_t at = row + idx * (TILE * 4);
    const uint32_t old = atoms_add(at, vlo);
    // hi += vhi + carry(old + vlo): add.cc / addc keep it at two integer instructions, and the
    // second atomic is issued unconditionally (vhi is almost never zero: values are ~2^45)
    uint32_t add;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(add) : "r"(old), "r"(vlo), "r"(vhi));
    reds_add(at + hiOff, add);
}

// Two adds of one edge, (idx, v) and (idx + 1, w): both returning low-word atomics are issued
// before either carry is formed, so their round trips overlap.  idx + 1 == bins (the reference
// drops that half, profile_ring.go:104-118) goes to a scratch row that is never flushed.
template <int TILE>
__device__ __forceinline__ void acc_add2(uint32_t row, uint32_t hiOff, int idx, int bins, long long v, long long w) {
    const int idxW = idx < bins - 1 ? idx + 1 : bins + 1;
    const uint32_t vlo = static_cast<uint32_t>(v), wlo = static_cast<uint32_t>(w);
    const uint32_t vhi = static_cast<uint32_t>(static_cast<unsigned long long>(v) >> 32);
    const uint32_t whi = static_cast<uint32_t>(static_cast<unsigned long long>(w) >> 32);
    const uint32_t atV = row + idx * (TILE * 4), atW = row + idxW * (TILE * 4);
    const uint32_t oldV = atoms_add(atV, vlo);
    const uint32_t oldW = atoms_add(atW, wlo);
    uint32_t addV, addW;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(addV) : "r"(oldV), "r"(vlo), "r"(vhi));
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %3, 0;\n\t}" : "=r"(addW) : "r"(oldW), "r"(wlo), "r"(whi));
    reds_add(atV + hiOff, addV);
    reds_add(atW + hiOff, addW);
}

constexpr int kChunkRecs = 128;   // raw records a warp draws at a time (4 x 32)
constexpr int kRingSlots = 64;    // per-warp ring of records that touch the tile

// One round: up to 32 records of the warp's ring (slots head .. head + cnt - 1): clip to the
// tile, expand to (record, LoS) pairs, evaluate the chords and add them to the tile.
template <bool TAPS, int TILE>
__device__ __forceinline__ void bin_round(const BinArgs &a, HitRec *wring, int head, int cnt, int lane, int t0,
                                          int t1, int bins, double rMin, double rMax, double invRMin,
                                          double invDr60, float cK, const double2 *sTrig, const double *sInvG,
                                          uint32_t *lo, uint32_t *hi, size_t tapRing, unsigned &nIns) {
    // rounds take 32 records at a time, so head is 0 or 32 and the round never wraps
    HitRec *grp = wring + head;
    // clip my record's two LoS ranges to the tile
    int len1 = 0, len2 = 0, st1 = 0, st2 = 0;
    if (lane < cnt) {
        const uint32_t r01 = grp[lane].r01, r23 = grp[lane].r23;
        st1 = max(static_cast<int>(r01 & 0xffffu), t0);
        len1 = max(min(static_cast<int>(r01 >> 16), t1) - st1, 0);
        st2 = max(static_cast<int>(r23 & 0xffffu), t0);
        len2 = max(min(static_cast<int>((r23 >> 16) & 0x7fffu), t1) - st2, 0);
        if (len1 == 0) st1 = t0;   // keep the packed start offsets within the tile
        if (len2 == 0) st2 = t0;
    }
    // cx^2 + cy^2 is formed once per record and takes the place of the LoS ranges (not needed
    // after the clipping above); its sign bit carries the centre-containing flag of r23
    if (lane < cnt) {
        const double cx = grp[lane].cx, cy = grp[lane].cy;
        const double pd2 = cx * cx + cy * cy;
        *reinterpret_cast<double *>(&grp[lane].r01) = (grp[lane].r23 & kHitCaseA) ? -pd2 : pd2;
    }
    __syncwarp();
    const unsigned nMine = static_cast<unsigned>(len1 + len2);
    unsigned incl = nMine;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned up = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up;
    }
    const unsigned nPairs = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned first = incl - nMine;   // index of my record's first pair
    // pair p of my record is LoS p + toLos1 of the tile while p < split, p + toLos2 from there on
    // (both biased by kLosBias to stay unsigned)
    constexpr unsigned kLosBias = 4096u;   // > 32 records x 64 LoS
    const unsigned metaA = (first + static_cast<unsigned>(len1)) |
                           (static_cast<unsigned>(st1 - t0) + kLosBias - first) << 16;
    const unsigned metaB = static_cast<unsigned>(st2 - t0) + kLosBias - static_cast<unsigned>(len1) - first;

    const uint32_t grpA = opaque_u32(smem_u32(grp));
    const uint32_t tabA = opaque_u32(smem_u32(sTrig));
    const uint32_t invGA = tabA + static_cast<uint32_t>(reinterpret_cast<const char *>(sInvG) - reinterpret_cast<const char *>(sTrig));
    const uint32_t loA = tabA + static_cast<uint32_t>(reinterpret_cast<const char *>(lo) - reinterpret_cast<const char *>(sTrig));
    const uint32_t hiOff = static_cast<uint32_t>(reinterpret_cast<const char *>(hi) - reinterpret_cast<const char *>(lo));
    constexpr unsigned kNoRecord = 0x80000000u;
    const unsigned firstKey = nMine ? first : kNoRecord;
    unsigned nBefore = 0;   // records whose first pair lies before this window
    for (unsigned p0 = 0; p0 < nPairs; p0 += 32) {
        // pair p -> record.  Every record of the round owns >= 1 pair, so the first-pair indices
        // are strictly increasing: one REDUX.OR marks the records that start inside this window
        // of 32 pairs, and pair p belongs to the last record starting at or before it.
        const unsigned p = p0 + lane;
        // 1 << (first - p0), or 0 where that is not in 0..31 (PTX shl clamps the amount to 32):
        // records that started earlier wrap around, lanes without a record carry kNoRecord
        const unsigned starts = __reduce_or_sync(0xffffffffu, shl_clamp(1u, firstKey - p0));
        const unsigned r = (nBefore + __popc(starts & (0xffffffffu >> (31 - lane))) - 1u) & 31u;
        nBefore += __popc(starts);
        const unsigned rmA = __shfl_sync(0xffffffffu, metaA, r);
        const unsigned rmB = __shfl_sync(0xffffffffu, metaB, r);
        if (p >= nPairs) continue;
        const int ll = static_cast<int>(p + (p < (rmA & 0xffffu) ? rmA >> 16 : rmB) - kLosBias);   // LoS within the tile
        const uint4 recA = lds128(grpA + r * 32u), recB = lds128(grpA + r * 32u + 16u);   // one HitRec
        const double2 cs = lds_f64x2(tabA + ll * 16u);
        const bool isA = (recA.w & kHitCaseA) != 0;
        const double rhoS = __hiloint2double(static_cast<int>(recB.y), static_cast<int>(recB.x));
        // literal geometry of insertToRing, los/halo.go:233-262
        const double cx = __uint_as_float(recA.x), cy = __uint_as_float(recA.y);
        const double projDist2 = fabs(__hiloint2double(static_cast<int>(recA.w), static_cast<int>(recA.z)));
        // max(rad2 - cz*cz, 0), formed once per record by hits_kernel
        const double projRad2 = __hiloint2double(static_cast<int>(recB.w), static_cast<int>(recB.z));
        const double b = cy * cs.x - cx * cs.y;
        const double b2 = b * b;
        const double m2 = projDist2 - b2;   // radicand of midDist / cMidDist
        const double d2 = projRad2 - b2;    // radicand of diff / radMidDist
        const bool bad = !(m2 >= 0) || !(d2 >= 0);   // a sqrt of the reference returns NaN
        // twoValIntrDist skips NaN (halo.go:262); oneValIntrDist's NaN goes into Insert
        if (bad && !isA) continue;
        const double mid = sqrt_pos0(m2), diff = sqrt_pos0(d2);
        const double rLo = mid - diff;
        // mid + diff, or diff - mid on the far side of a centre-containing circle
        // (diff - mid is diff + (-mid) exactly: the sign of mid is flipped as an integer)
        const bool far = isA && !(cx * cs.x + cy * cs.y > 0);
        const double rHi = diff + __hiloint2double(__double2hiint(mid) ^ (far ? 0x80000000 : 0), __double2loint(mid));
        const bool endNaN = bad;            // only reachable with isA
        // ProfileRing.Insert, profile_ring.go:92-120, with ln monotone:
        // end <= lowR <=> rHi <= rMin, start >= highR <=> rLo >= rMax, ...
        if (!endNaN && rHi <= rMin) continue;
        if (!isA && rLo >= rMax) continue;
        nIns++;
        const uint32_t row = loA + ll * 4u;
        const size_t tapRow = tapRing + t0 + ll;
        if (TAPS) atomicAdd(&a.tapInsert[tapRow], 1u);
        const long long Q = __double2ll_rn(rhoS);
        const bool startIn = !isA && rLo > rMin;
        const bool endIn = !endNaN && rHi < rMax;
        if (startIn) {
            int idx; double rem;
            bin_of(rLo, invRMin, cK, invDr60, invGA, idx, rem);
            const long long q1 = __double2ll_rn(rhoS * rem);
            acc_add2<TILE>(row, hiOff, idx, bins, Q - q1, q1);
            if (TAPS && idx < bins) atomicAdd(&a.tapStart[tapRow * bins + idx], 1u);
        } else {
            acc_add<TILE>(row, hiOff, 0, Q);
            if (TAPS) atomicAdd(&a.tapStart[tapRow * bins], 1u);
        }
        if (endIn) {
            int idx; double rem;
            bin_of(rHi, invRMin, cK, invDr60, invGA, idx, rem);
            const long long q1 = __double2ll_rn(rhoS * rem);
            acc_add2<TILE>(row, hiOff, idx, bins, q1 - Q, -q1);
            if (TAPS && idx < bins) atomicAdd(&a.tapEnd[tapRow * bins + idx], 1u);
        }
    }
}

// BINS: RadialBins as a compile-time constant (256, the default), or 0 for any other value; with
// it the high words sit at a constant offset from the low ones and the row stride folds into
// the atomics' addresses.
// TILE: lines of sight per CTA.  32 with 16 warps and two CTAs per SM, or -- the product
// instantiation -- 64 with 32 warps and one CTA per SM: a CTA scans every record of its ring for
// the ones that touch its tile, so twice the tile is half the scanning and fewer, longer clipped
// ranges per record (-2.5 %); the 132 KB of accumulators then leave room for one CTA only.
template <bool TAPS, int BINS, int TILE>
__global__ void __launch_bounds__(TILE * 16, 64 / TILE) bin_kernel(BinArgs a) {
    constexpr int kTileLos = TILE;              // shadows the narrow width of sfk_internal.h
    constexpr int kBinConsumers = TILE / 2;     // warps
    constexpr int kBinThreads = kBinConsumers * 32;
    extern __shared__ __align__(128) unsigned char smemRaw[];
    const int bins = BINS ? BINS : a.bins;
    // + 1: row for a start/end edge that lands exactly on rMax; + 1: acc_add2's scratch row
    const int stride = bins + 2;
    // byte offsets from the one dynamic array: the compiler must keep seeing shared addresses
    HitRec *rings = reinterpret_cast<HitRec *>(smemRaw);                          // [warps][kRingSlots]
    double2 *sTrig = reinterpret_cast<double2 *>(rings + kBinConsumers * kRingSlots);   // [TILE] (cos, sin)
    double *sInvG = reinterpret_cast<double *>(sTrig + kTileLos);                 // [bins + 1]
    unsigned int *nextChunk = reinterpret_cast<unsigned int *>(sInvG + bins + 1);     // 8-byte slot
    // accumulators are bin-major, [stride][TILE]: a warp's lanes are different LoS of (mostly) one
    // record, so the bank is the LoS (mod 32) and the atomics are conflict-free whatever the bins are
    uint32_t *lo = nextChunk + 2;
    uint32_t *hi = lo + kTileLos * stride;

    const int nTiles = (a.spokes + kTileLos - 1) / kTileLos;
    const int tile = blockIdx.x % nTiles;
    const int chunk = blockIdx.x / nTiles;
    const int ring = blockIdx.y;
    const int slot = blockIdx.z;
    const int haloIdx = a.batchHalos[slot];
    const HaloDev &h = a.halos[haloIdx];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t0 = tile * kTileLos;
    const int t1 = min(t0 + kTileLos, a.spokes);

    for (int k = threadIdx.x; k < 2 * kTileLos * stride; k += kBinThreads) lo[k] = 0;
    for (int k = threadIdx.x; k <= bins; k += kBinThreads) sInvG[k] = a.binInvG[k];
    if (threadIdx.x < kTileLos) {
        const int l = min(t0 + static_cast<int>(threadIdx.x), a.spokes - 1);
        sTrig[threadIdx.x] = make_double2(a.ringCos[l], a.ringSin[l]);
    }
    if (threadIdx.x == 0) *nextChunk = 0;
    __syncthreads();

    // this CTA's share of the ring's records
    const size_t hr = static_cast<size_t>(slot) * a.rings + ring;
    // hits_kernel fills a list from both ends (circles whose chords all start in bin 0 at the
    // front): records 0 .. nFront - 1 are contiguous, the others follow after a gap
    const int64_t nFront = a.hitCount[2 * hr];
    const int64_t total = nFront + a.hitCount[2 * hr + 1];
    const int gap = static_cast<int>(a.hitOffset[hr + 1] - a.hitOffset[hr] - total);
    const int64_t rec0 = (total * chunk / a.chunks) & ~int64_t(31);
    const int64_t rec1 = chunk + 1 == a.chunks ? total : (total * (chunk + 1) / a.chunks) & ~int64_t(31);
    const HitRec *src = a.recs + a.hitOffset[hr] + rec0;
    const int nRecs = static_cast<int>(rec1 - rec0);
    // first record of this CTA's share that sits behind the gap
    const int backFrom = static_cast<int>(max(int64_t(0), min(int64_t(nRecs), nFront - rec0)));
    unsigned nIns = 0;

    {
        const double rMin = h.rMin, rMax = h.rMax, invRMin = h.invRMin;
        const double invDr60 = a.invDr60;
        const float cK = a.cK;
        const size_t tapRing = (static_cast<size_t>(haloIdx) * a.rings + ring) * a.spokes;
        HitRec *wring = rings + warp * kRingSlots;
        int head = 0, count = 0;   // records waiting in the ring: slots head .. head + count - 1 (mod 64)

        // A draw is the first record of a chunk; bit 31: a single group of 32 records instead of
        // kChunkRecs.  Guided self-scheduling: whole chunks while plenty is left, single groups for
        // the last kTailRecs records of the list, so that the warps of the CTA finish within one
        // group of each other (with one CTA per SM nothing else fills the SM while the last warp
        // works off a whole chunk).
        constexpr unsigned kSmall = 0x80000000u;
        constexpr unsigned kTailRecs = kBinThreads * 4;   // one whole chunk per warp
        // (r01, r23) of record start + sub * 32 + lane of the drawn chunk; (0, 0) past its end
        auto load_ranges = [&](unsigned c, uint2 (&rr)[kChunkRecs / 32]) {
            const int start = static_cast<int>(c & ~kSmall), len = (c & kSmall) ? 32 : kChunkRecs;
#pragma unroll
            for (int sub = 0; sub < kChunkRecs / 32; sub++) {
                const int idx = start + sub * 32 + lane;
                rr[sub] = make_uint2(0u, 0u);
                if (sub * 32 < len && idx < nRecs) rr[sub] = *reinterpret_cast<const uint2 *>(&src[idx + (idx >= backFrom ? gap : 0)].r01);
            }
        };
        auto draw = [&]() {
            unsigned c = 0;
            if (lane == 0) {
                const unsigned seen = *reinterpret_cast<volatile unsigned *>(nextChunk);
                const bool small = seen + kTailRecs >= static_cast<unsigned>(nRecs);
                c = atomicAdd(nextChunk, small ? 32u : static_cast<unsigned>(kChunkRecs)) | (small ? kSmall : 0u);
            }
            return __shfl_sync(0xffffffffu, c, 0);
        };
        unsigned cur = draw();
        uint2 rr[kChunkRecs / 32];
        load_ranges(cur, rr);
        for (;;) {
            const bool flush = (cur & ~kSmall) >= static_cast<unsigned>(nRecs);   // nothing left to draw: work off what is waiting
            if (flush && count == 0) break;
            if (!flush) {
                // draw the next chunk and start its range loads before working on this one
                const unsigned nxt = draw();
                uint2 nr[kChunkRecs / 32];
                load_ranges(nxt, nr);
#pragma unroll
                for (int sub = 0; sub < kChunkRecs / 32; sub++) {
                    // a kept record has at least one LoS in the tile (bin_round relies on it)
                    const bool ov = min(static_cast<int>(rr[sub].x >> 16), t1) > max(static_cast<int>(rr[sub].x & 0xffffu), t0) ||
                                    min(static_cast<int>((rr[sub].y >> 16) & 0x7fffu), t1) > max(static_cast<int>(rr[sub].y & 0xffffu), t0);
                    const unsigned m = __ballot_sync(0xffffffffu, ov);
                    if (ov) {
                        const int at = (head + count + __popc(m & ((1u << lane) - 1u))) & (kRingSlots - 1);
                        const int idx = static_cast<int>(cur & ~kSmall) + sub * 32 + lane;
                        const HitRec *from = src + idx + (idx >= backFrom ? gap : 0);
                        cp_async16(&wring[at], from);
                        cp_async16(reinterpret_cast<char *>(&wring[at]) + 16, reinterpret_cast<const char *>(from) + 16);
                    }
                    count += __popc(m);
                    if (count >= 32) {
                        cp_async_wait_all();
                        __syncwarp();
                        bin_round<TAPS, TILE>(a, wring, head, 32, lane, t0, t1, bins, rMin, rMax, invRMin, invDr60, cK,
                                        sTrig, sInvG, lo, hi, tapRing, nIns);
                        __syncwarp();   // the slots may be overwritten from here on
                        head = (head + 32) & (kRingSlots - 1);
                        count -= 32;
                    }
                }
                cur = nxt;
#pragma unroll
                for (int sub = 0; sub < kChunkRecs / 32; sub++) rr[sub] = nr[sub];
            } else {
                cp_async_wait_all();
                __syncwarp();
                bin_round<TAPS, TILE>(a, wring, head, count, lane, t0, t1, bins, rMin, rMax, invRMin, invDr60, cK,
                                sTrig, sInvG, lo, hi, tapRing, nIns);
                count = 0;
            }
        }
        // the "particle-ray intersections" metric
        unsigned tot = nIns;
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
        if (lane == 0 && tot) atomicAdd(&a.nIntersections[haloIdx], static_cast<unsigned long long>(tot));
    }
    __syncthreads();

    // flush the tile into the [LoS][bin] rows of the grid
    unsigned long long *g = a.grid + (static_cast<size_t>(slot) * a.rings + ring) * a.spokes * bins;
    if (a.chunks == 1 && (bins & 3) == 0) {
        // This CTA is the only writer of its 32 rows: plain stores of every cell, zeros included,
        // so the grid needs no memset and no read-modify-write.  Lane = LoS (conflict-free reads
        // of the bin-major tile), four consecutive bins = one full 32-byte sector per lane.
        for (int tl = lane; tl < kTileLos; tl += 32) {
            const int gl = t0 + tl;
            for (int b0 = warp * 4; b0 < bins; b0 += kBinConsumers * 4) {
                unsigned long long v[4];
#pragma unroll
                for (int c = 0; c < 4; c++)
                    v[c] = (static_cast<unsigned long long>(hi[(b0 + c) * kTileLos + tl]) << 32) | lo[(b0 + c) * kTileLos + tl];
                if (gl < a.spokes) {
                    ulonglong2 *dst = reinterpret_cast<ulonglong2 *>(g + static_cast<size_t>(gl) * bins + b0);
                    dst[0] = make_ulonglong2(v[0], v[1]);
                    dst[1] = make_ulonglong2(v[2], v[3]);
                }
            }
            // index == bins lands in the next LoS's first bin (derivs is one slice), which may belong
            // to another CTA: such edges (start or end within an ulp of rMax) go to a side list that
            // bin_overflow_kernel adds once every tile is stored
            if (warp == 0 && gl + 1 < a.spokes) {
                const unsigned long long v = (static_cast<unsigned long long>(hi[bins * kTileLos + tl]) << 32) | lo[bins * kTileLos + tl];
                if (v != 0) {
                    const unsigned at = atomicAdd(a.ovfCount, 1u);
                    if (at < a.ovfCap) {
                        a.ovfList[2 * at] = static_cast<unsigned long long>((g - a.grid) + static_cast<size_t>(gl + 1) * bins);
                        a.ovfList[2 * at + 1] = v;
                    }
                }
            }
        }
        return;
    }
    for (int k = threadIdx.x; k < kTileLos * stride; k += kBinThreads) {
        const int l = k / stride, bin = k - l * stride;
        const int gl = tile * kTileLos + l;
        if (gl >= a.spokes) continue;
        const int at = bin * kTileLos + l;   // bank-conflicting read, once per CTA
        const unsigned long long v = (static_cast<unsigned long long>(hi[at]) << 32) | lo[at];
        if (v == 0) continue;
        if (bin < bins) {
            atomicAdd(&g[static_cast<size_t>(gl) * bins + bin], v);
        } else if (bin == bins && gl + 1 < a.spokes) {
            // index == bins lands in the next LoS's first bin (derivs is one slice)
            atomicAdd(&g[static_cast<size_t>(gl + 1) * bins], v);
        }
    }
}

// Adds the side list of the plain-store flush (edges that land in the next LoS's first bin) and
// re-arms it; ovfCount[1] counts entries that did not fit (sfk_finish_snapshot fails on it).
__global__ void bin_overflow_kernel(unsigned long long *grid, const unsigned long long *list, unsigned *count, unsigned cap) {
    const unsigned n = count[0];
    for (unsigned e = threadIdx.x; e < min(n, cap); e += blockDim.x) atomicAdd(&grid[list[2 * e]], list[2 * e + 1]);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (n > cap) count[1] += n - cap;
        count[0] = 0;
    }
}

}  // namespace

bool bin_stores_whole_grid(const BinArgs &a) { return a.chunks == 1 && (a.bins & 3) == 0; }

// LoS per CTA of the instantiation launch_bin picks
int bin_tile_los(int bins, bool taps) { return !taps && bins == 256 ? 64 : kTileLos; }

size_t bin_smem_bytes(int bins, bool taps) {
    const int tile = bin_tile_los(bins, taps), warps = tile / 2;
    return sizeof(HitRec) * warps * kRingSlots + sizeof(double2) * tile + sizeof(double) * (bins + 1) +
           2 * sizeof(unsigned int) + sizeof(uint32_t) * 2 * tile * (bins + 2);
}

cudaError_t launch_bin(const BinArgs &a, int nb, bool taps, cudaStream_t s) {
    if (nb <= 0) return cudaSuccess;
    const size_t smem = bin_smem_bytes(a.bins, taps);
    {   // per device, so set it on every launch
        cudaError_t e = cudaFuncSetAttribute(bin_kernel<true, 0, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bin_kernel<false, 0, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(bin_kernel<false, 256, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
    }
    const int tile = bin_tile_los(a.bins, taps);
    const int nTiles = (a.spokes + tile - 1) / tile;
    dim3 grid(nTiles * a.chunks, a.rings, nb);
    if (taps) bin_kernel<true, 0, 32><<<grid, 512, smem, s>>>(a);
    else if (a.bins == 256) bin_kernel<false, 256, 64><<<grid, 1024, smem, s>>>(a);
    else bin_kernel<false, 0, 32><<<grid, 512, smem, s>>>(a);
    if (bin_stores_whole_grid(a)) bin_overflow_kernel<<<1, 256, 0, s>>>(a.grid, a.ovfList, a.ovfCount, a.ovfCap);
    return cudaGetLastError();
}

}  // namespace sfk
