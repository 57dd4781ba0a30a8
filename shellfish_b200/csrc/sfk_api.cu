// sfk_api.cu -- C ABI (include/shellfish_b200.h) and host orchestration of
// the sm_100a shell-finder kernels.  Mirrors the control flow of
// cmd/shell.go: createHalos -> sphereLoop (per block) -> haloAnalysis.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "sfk_comm.h"
#include "sfk_kernels.cuh"
#include "sfk_sgmath.cuh"
#include "sfk_shellmath.cuh"

using namespace sfk;

namespace {

// Go constant 4*math.Pi/3 folded exactly (cmd/shell.go:485,549).
const double kFourPiOver3 = 4.18879020478639098461685784437267051226289253250014;

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;  // elements
    ~DevBuf() { release(); }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    // Grows without preserving contents.
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T));
        if (e == cudaSuccess) cap = n; else p = nullptr;
        return e;
    }
    cudaError_t upload(const std::vector<T> &v, cudaStream_t s) {
        cudaError_t e = ensure(v.size());
        if (e != cudaSuccess || v.empty()) return e;
        return cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    }
};

struct StageTimer {
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[6];
    void begin(int stage, cudaStream_t s) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, s);
        ev[stage].push_back({a, b});
    }
    void end(int stage, cudaStream_t s) { cudaEventRecord(ev[stage].back().second, s); }
    double collect(int stage) {
        double ms = 0;
        for (auto &pr : ev[stage]) {
            float t = 0;
            if (cudaEventElapsedTime(&t, pr.first, pr.second) == cudaSuccess) ms += t;
            cudaEventDestroy(pr.first); cudaEventDestroy(pr.second);
        }
        ev[stage].clear();
        return ms;
    }
};
enum { ST_CULL = 0, ST_HITS, ST_BIN, ST_LOS, ST_FILTER, ST_FIT };

}  // namespace

struct sfk_ctx {
    sfk_params p{};
    Tables tab;
    std::string err;
    cudaStream_t stream = nullptr;
    int rings = 0, spokes = 0, bins = 0, P2 = 0;
    bool taps = false;

    DevBuf<float> dNorms, dRots, dIrots;
    DevBuf<double> dCos, dSin, dPhi, dTaps0, dTaps1, dBinInvG, dLogTab;
    bool tapsReady = false;  // Sav-Gol kernels cached for the process, smooth.go:84-96
    SgPoly sg{};             // the same taps as polynomials of the offset (sliding-moment smoothing)
    bool sgReady = false;

    // snapshot state
    bool inSnapshot = false, finished = false;
    sfk_cosmo cosmo0{};
    float minMass = 0;
    std::vector<HaloDev> halos;
    std::vector<double> ox, oy, oz;   // float64 origins for SheetIntersect
    DevBuf<HaloDev> dHalos;
    // candidates: appended in arrival order by cull_kernel (dCandsRaw), bucketed by halo at
    // finish (dCands).  candUpper bounds the device-side count without a round trip.
    DevBuf<Cand> dCandsRaw, dCands;
    DevBuf<unsigned long long> dCandTotal;
    DevBuf<uint32_t> dHaloCandCount, dCandCursor;
    DevBuf<int64_t> dCandOffset;
    int64_t candUpper = 0, candCount = 0;
    bool pushed = false, cullOpen = false;
    std::vector<int64_t> candPerHalo;
    std::vector<int64_t> interPerHalo;   // ProfileRing.Insert calls per halo (after finish)

    // block staging: two buffers, so that the host-to-device copy of a block (copyStream)
    // overlaps the cull of the previous one (stream)
    DevBuf<float> dXyz[2], dMass[2];
    cudaStream_t copyStream = nullptr;
    cudaEvent_t evCopied[2] = {nullptr, nullptr}, evCulled[2] = {nullptr, nullptr};
    bool culledValid[2] = {false, false};
    int64_t pushIdx = 0;
    DevBuf<int32_t> dHaloList;

    // results (per halo of the snapshot)
    DevBuf<double> dRsp, dCoeffs;
    DevBuf<int32_t> dStatus;
    DevBuf<unsigned long long> dNInter, dRhoMax;
    DevBuf<long long> dNPoints;
    DevBuf<uint32_t> dHitCount;
    DevBuf<double> dProfiles;
    DevBuf<double> dFitScratch;
    // `stats` M_sp pass
    DevBuf<float> dMassCenters, dMassLow2, dMassHigh2;
    DevBuf<double> dMassCoeffs, dMassSums;
    DevBuf<int> dMassList;       // [nHalo + 1 + 6]: halo list, its length, bounding-box keys
    int64_t massHalos = 0;
    int massOrder = 0;
    bool massOpen = false;
    bool massDevSynced = false;   // work on other streams was waited for at the first device push of this pass
    // ShellParticleFile pass
    bool bandOn = false;
    int bandInside = 0;
    DevBuf<float> dBandLow2, dBandHigh2;
    DevBuf<double> dBandDelta;
    DevBuf<long long> dBandList;
    DevBuf<unsigned long long> dBandCount;
    std::vector<long long> hBand;
    // `stats` shell properties
    DevBuf<double> dPropCoeffs, dPropSums, dGlX, dGlW;
    int glTheta = 0;
    DevBuf<uint32_t> dTapInsert, dTapStart, dTapEnd;
    DevBuf<unsigned long long> dGridTap;   // SFK_FLAG_KEEP_GRID: [nHalo][rings][spokes][bins]
    bool keepGrid = false;

    // batch scratch
    DevBuf<Piece> dPieces;
    DevBuf<HitRec> dRecs;
    std::vector<int32_t> blockList;   // halos touching the block being pushed
    DevBuf<int64_t> dHitOffset;
    DevBuf<uint32_t> dHitCursor, dBatchHitCount;
    DevBuf<int32_t> dBatchHalos, dHaloSlot;
    DevBuf<unsigned long long> dGrid, dOvfList;
    DevBuf<unsigned> dOvfCount;
    DevBuf<double> dFR;
    DevBuf<int32_t> dFIdx, dFCount;

    // state shared by the stages of sfk_finish_snapshot / the split-halo calls
    std::vector<std::vector<Piece>> perHalo;
    std::vector<uint32_t> hHitCount;
    std::vector<int64_t> hitsPerHalo;
    std::vector<int32_t> order;
    DevBuf<uint32_t> dHitCountGlobal;
    int splitPhase = 0;
    // NCCL communicator of this context (sfk_comm_init) and the row-indexed R_sp scratch of the
    // ring-partitioned analysis
    ncclComm_t comm = nullptr;
    int commRank = 0, commSize = 1;
    DevBuf<double> dRspRows;
    DevBuf<long long> dFlatList;
    DevBuf<unsigned long long> dFlatCount;
    bool flatTotalArmed = false;

    sfk_stats stats{};
    StageTimer timer;
    std::vector<double> hCoeffs;
    std::vector<int32_t> hStatus;
};

namespace {

int fail(sfk_ctx *c, int code, const std::string &msg) {
    if (c) c->err = msg;
    return code;
}

#define CU(call)                                                                       \
    do {                                                                               \
        cudaError_t e_ = (call);                                                       \
        if (e_ != cudaSuccess)                                                         \
            return fail(ctx, SFK_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// No C++ exception may cross the C ABI: every entry point that takes a context is a
// function-try-block ending in this handler.
#define SFK_CATCH(ctx)                                                                        \
    catch (const std::bad_alloc &) { return fail(ctx, SFK_E_NOMEM, "out of host memory"); }   \
    catch (const std::exception &e) { return fail(ctx, SFK_E_INVALID, std::string("internal error: ") + e.what()); }

// ShellConfig.validate, cmd/shell.go:143-187
int validate(sfk_ctx *ctx, const sfk_params &p) {
    char buf[160];
    auto bad_i = [&](const char *name, long long v) {
        snprintf(buf, sizeof buf, "The variable '%s' was set to %lld.", name, v);
        return fail(ctx, SFK_E_INVALID, buf);
    };
    auto bad_f = [&](const char *name, double v) {
        snprintf(buf, sizeof buf, "The variable '%s' was set to %g.", name, v);
        return fail(ctx, SFK_E_INVALID, buf);
    };
    if (p.subsample_factor <= 0) return bad_i("SubsampleFactor", p.subsample_factor);
    if (p.radial_bins <= 0) return bad_i("RadialBins", p.radial_bins);
    if (p.spokes <= 0) return bad_i("Spokes", p.spokes);
    if (p.rings <= 0) return bad_i("Rings", p.rings);
    if (p.r_max_mult <= 0) return bad_f("RMaxMult", p.r_max_mult);
    if (p.r_min_mult <= 0) return bad_f("RMinMult", p.r_min_mult);
    if (p.r_kernel_mult <= 0) return bad_f("RKernelMult", p.r_kernel_mult);
    if (p.eta <= 0) return bad_f("Eta", p.eta);
    if (p.order <= 0) return bad_i("Order", p.order);
    if (p.levels <= 0) return bad_i("Levels", p.levels);
    if (p.smoothing_window <= 0) return bad_i("SmoothingWindow", p.smoothing_window);
    if (p.r_min_mult >= p.r_max_mult) {
        snprintf(buf, sizeof buf,
                 "The variable '%s' was set to %g, but the variable '%s' was set to %g.",
                 "RMinMult", p.r_min_mult, "RMaxMult", p.r_max_mult);
        return fail(ctx, SFK_E_INVALID, buf);
    }
    // limits of this implementation
    if (p.percentile_profile && !(p.percentile >= 0 && p.percentile <= 100))   // msort.Percentile panics, median.go:11-13
        return fail(ctx, SFK_E_INVALID, "percentile must be in the range [0, 1]");
    if (p.order > kMaxOrder) return fail(ctx, SFK_E_UNSUPPORTED, "Order above 4 is not supported.");
    if (p.levels > kMaxLevels) return fail(ctx, SFK_E_UNSUPPORTED, "Levels above 6 is not supported.");
    if (p.rings > kMaxRings) return fail(ctx, SFK_E_UNSUPPORTED, "Rings above 1024 is not supported.");
    if (p.spokes > 16384) return fail(ctx, SFK_E_UNSUPPORTED, "Spokes above 16384 is not supported.");
    // NewSavGolKernel's panics (filter.go:218-262) are only reachable where the reference builds
    // the kernels: Smooth returns ok = false before that when RadialBins <= SmoothingWindow
    // (smooth.go:51-53), and PercentileProfile never smooths
    if (!p.percentile_profile && p.radial_bins > p.smoothing_window) {
        if (p.smoothing_window % 2 != 1) return fail(ctx, SFK_E_INVALID, "Kernel width must be odd.");
        if (p.smoothing_window <= 4) return fail(ctx, SFK_E_INVALID, "Kernel width cannot be smaller than pOrder.");
    }
    if (bin_smem_bytes(static_cast<int>(p.radial_bins), (p.flags & SFK_FLAG_TAPS) != 0) > 227 * 1024)
        return fail(ctx, SFK_E_UNSUPPORTED, "RadialBins too large for one shared-memory tile.");
    return SFK_OK;
}

// los/halo.go:442-467
double wrap_dist(double x1, double x2, double width) {
    double dist = x1 - x2;
    if (dist > width / 2) return dist - width;
    if (dist < width / -2) return dist + width;
    return dist;
}
bool in_range(double x, double r, double low, double width, double tw) {
    return (x + r > low && x - r < low + width) ||
           (wrap_dist(x, low, tw) > -r && wrap_dist(x, low + width, tw) < r);
}
bool sheet_intersect(const sfk_ctx *c, size_t i, const float origin[3], const float width[3], double tw) {
    const double r = c->halos[i].rMax;
    return in_range(c->ox[i], r, origin[0], width[0], tw) &&
           in_range(c->oy[i], r, origin[1], width[1], tw) &&
           in_range(c->oz[i], r, origin[2], width[2], tw);
}

// Makes room for `extra` more candidates.  The device-side count is only known to the host
// as an upper bound (every pushed block could have kept all its particles for every listed
// halo); when that bound no longer fits, the exact count is fetched (one stream sync) and the
// list grows geometrically if it must.
int reserve_cands(sfk_ctx *ctx, int64_t extra) {
    if (static_cast<size_t>(ctx->candUpper + extra) <= ctx->dCandsRaw.cap) { ctx->candUpper += extra; return SFK_OK; }
    unsigned long long exact = 0;
    CU(cudaMemcpyAsync(&exact, ctx->dCandTotal.p, sizeof exact, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->candUpper = static_cast<int64_t>(exact);
    const size_t need = static_cast<size_t>(ctx->candUpper + extra);
    if (need > ctx->dCandsRaw.cap) {
        const size_t cap = std::max<size_t>({need, ctx->dCandsRaw.cap * 3 / 2, size_t(1) << 22});
        Cand *np = nullptr;
        CU(cudaMalloc(&np, cap * sizeof(Cand)));
        if (exact > 0)
            CU(cudaMemcpyAsync(np, ctx->dCandsRaw.p, exact * sizeof(Cand), cudaMemcpyDeviceToDevice, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->dCandsRaw.release();
        ctx->dCandsRaw.p = np;
        ctx->dCandsRaw.cap = cap;
    }
    ctx->candUpper += extra;
    return SFK_OK;
}

// binIntersections for one header, cmd/shell.go:599-611 (halo order kept)
void block_halo_list(const sfk_ctx *ctx, const float origin[3], const float width[3], double tw,
                     std::vector<int32_t> &list, int64_t &nOwned) {
    list.clear();
    nOwned = 0;
    for (size_t i = 0; i < ctx->halos.size(); i++)
        if (ctx->halos[i].valid && sheet_intersect(ctx, i, origin, width, tw)) {
            list.push_back(static_cast<int32_t>(i));
            nOwned += ctx->halos[i].owned ? 1 : 0;
        }
}

int push_block_impl(sfk_ctx *ctx, const float *dxyz, const float *dmass, int64_t n,
                    const sfk_cosmo *cosmo, double tw, const std::vector<int32_t> &list, int64_t nOwned) {
    if (list.empty() || n <= 0) return SFK_OK;
    const int nList = static_cast<int>(list.size());
    const int64_t sf = ctx->p.subsample_factor;
    const int64_t sf3 = sf * sf * sf;
    int rc = reserve_cands(ctx, (n + sf3 - 1) / sf3 * nOwned);
    if (rc) return rc;
    // pageable source: the copy is staged before the call returns, `list` may die
    CU(ctx->dHaloList.upload(list, ctx->stream));
    CullArgs a{};
    a.xyz = dxyz; a.mass = dmass; a.uniformMass = ctx->minMass; a.n = n; a.sf3 = sf3;
    a.haloList = ctx->dHaloList.p; a.nList = nList; a.halos = ctx->dHalos.p;
    a.tw = static_cast<float>(tw);
    // chanLoadSphereVec uses the block's own header, cmd/shell.go:487-488
    a.rhoM = rho_average(cosmo->h100 * 100, cosmo->omega_m, cosmo->omega_l, cosmo->z);
    a.tma = sf3 == 1 && reinterpret_cast<uintptr_t>(dxyz) % 16 == 0 && reinterpret_cast<uintptr_t>(dmass) % 16 == 0;   // null is aligned
    a.cands = ctx->dCandsRaw.p; a.cap = static_cast<int64_t>(ctx->dCandsRaw.cap);
    a.total = ctx->dCandTotal.p; a.haloCount = ctx->dHaloCandCount.p;
    if (!ctx->cullOpen) { ctx->timer.begin(ST_CULL, ctx->stream); ctx->cullOpen = true; }
    CU(launch_cull(a, ctx->stream));
    ctx->stats.launches++;
    ctx->pushed = true;
    return SFK_OK;
}

// ---- stages of haloAnalysis' GPU replacement ---------------------------------

// Result buffers, Savitzky-Golay taps, work pieces and the raw hit-count pass.
int stage_prepare(sfk_ctx *ctx) {
    cudaStream_t st = ctx->stream;
    const int64_t nHalo = static_cast<int64_t>(ctx->halos.size());
    const int R = ctx->rings, n = ctx->spokes, B = ctx->bins, P2 = ctx->P2;
    const size_t rn = static_cast<size_t>(R) * n, rnb = rn * B;
    ctx->hCoeffs.assign(static_cast<size_t>(nHalo) * P2, 0.0);
    ctx->hStatus.assign(nHalo, SFK_HALO_OK);
    ctx->perHalo.assign(nHalo, {});
    ctx->order.clear();
    if (nHalo == 0) return SFK_OK;

    CU(ctx->dRsp.ensure(nHalo * rn));
    CU(ctx->dCoeffs.ensure(static_cast<size_t>(nHalo) * P2));
    CU(ctx->dStatus.ensure(nHalo));
    CU(ctx->dNInter.ensure(nHalo));
    CU(ctx->dRhoMax.ensure(nHalo));
    CU(ctx->dNPoints.ensure(nHalo));
    CU(ctx->dHitCount.ensure(static_cast<size_t>(nHalo) * R));
    CU(cudaMemsetAsync(ctx->dRsp.p, 0xff, nHalo * rn * sizeof(double), st));  // NaN
    CU(cudaMemsetAsync(ctx->dCoeffs.p, 0, static_cast<size_t>(nHalo) * P2 * sizeof(double), st));
    CU(cudaMemsetAsync(ctx->dStatus.p, 0, nHalo * sizeof(int32_t), st));
    CU(cudaMemsetAsync(ctx->dNInter.p, 0, nHalo * sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(ctx->dRhoMax.p, 0, nHalo * sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(ctx->dNPoints.p, 0, nHalo * sizeof(long long), st));
    CU(cudaMemsetAsync(ctx->dHitCount.p, 0, static_cast<size_t>(nHalo) * R * sizeof(uint32_t), st));
    if (ctx->taps) {
        CU(ctx->dProfiles.ensure(nHalo * rnb));
        CU(ctx->dTapInsert.ensure(nHalo * rn));
        CU(ctx->dTapStart.ensure(nHalo * rnb));
        CU(ctx->dTapEnd.ensure(nHalo * rnb));
        CU(cudaMemsetAsync(ctx->dProfiles.p, 0, nHalo * rnb * sizeof(double), st));
        CU(cudaMemsetAsync(ctx->dTapInsert.p, 0, nHalo * rn * sizeof(uint32_t), st));
        CU(cudaMemsetAsync(ctx->dTapStart.p, 0, nHalo * rnb * sizeof(uint32_t), st));
        CU(cudaMemsetAsync(ctx->dTapEnd.p, 0, nHalo * rnb * sizeof(uint32_t), st));
    }
    if (ctx->keepGrid) {
        CU(ctx->dGridTap.ensure(nHalo * rnb));
        CU(cudaMemsetAsync(ctx->dGridTap.p, 0, nHalo * rnb * sizeof(unsigned long long), st));
    }

    // Savitzky-Golay taps: built once per process from the first smoothed
    // profile's dx = ln r1 - ln r0 (smooth.go:63,84-96)
    if (!ctx->tapsReady && B > ctx->p.smoothing_window) {
        for (const HaloDev &h : ctx->halos) {
            if (!h.valid) continue;
            const double r0 = std::exp(h.lrMin + h.dlr * (0.0 + 0.5));
            const double r1 = std::exp(h.lrMin + h.dlr * (1.0 + 0.5));
            const double dx = std::log(r1) - std::log(r0);
            std::vector<double> k0, k1;
            if (!savgol_taps(4, static_cast<int>(ctx->p.smoothing_window), 0, 0, k0) ||
                !savgol_taps(4, static_cast<int>(ctx->p.smoothing_window), 1, dx, k1))
                return fail(ctx, SFK_E_INVALID, "Savitzky-Golay kernel construction failed");
            CU(ctx->dTaps0.upload(k0, st));
            CU(ctx->dTaps1.upload(k1, st));
            CU(cudaStreamSynchronize(st));
            ctx->tapsReady = true;
            ctx->sgReady = savgol_poly(k0, k1, ctx->sg);
            break;
        }
    }
    if (!ctx->tapsReady) {
        std::vector<double> z(std::max<int64_t>(ctx->p.smoothing_window, 1), 0.0);
        CU(ctx->dTaps0.upload(z, st));
        CU(ctx->dTaps1.upload(z, st));
        CU(cudaStreamSynchronize(st));
    }

    // bucket the candidates by halo (counting sort on the device), then cut every halo's run
    // into pieces of <= 256 candidates
    {
        std::vector<uint32_t> cnt(nHalo);
        unsigned long long total = 0;
        CU(cudaMemcpyAsync(cnt.data(), ctx->dHaloCandCount.p, nHalo * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
        CU(cudaMemcpyAsync(&total, ctx->dCandTotal.p, sizeof total, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        if (total > ctx->dCandsRaw.cap) return fail(ctx, SFK_E_NOMEM, "candidate list overflow");
        ctx->candCount = static_cast<int64_t>(total);
        std::vector<int64_t> offset(nHalo);
        int64_t run = 0;
        for (int64_t i = 0; i < nHalo; i++) { offset[i] = run; run += cnt[i]; ctx->candPerHalo[i] = cnt[i]; }
        CU(ctx->dCands.ensure(std::max<int64_t>(run, 1)));
        CU(ctx->dCandOffset.upload(offset, st));
        CU(ctx->dCandCursor.ensure(nHalo));
        CU(cudaMemsetAsync(ctx->dCandCursor.p, 0, nHalo * sizeof(uint32_t), st));
        CU(launch_bucket(ctx->dCandsRaw.p, ctx->candCount, ctx->dCandOffset.p, ctx->dCandCursor.p, ctx->dCands.p, st));
        ctx->stats.launches++;
        if (ctx->cullOpen) { ctx->timer.end(ST_CULL, st); ctx->cullOpen = false; }
        CU(cudaStreamSynchronize(st));   // `offset` is a local vector
        for (int64_t i = 0; i < nHalo; i++)
            for (int64_t off = 0; off < static_cast<int64_t>(cnt[i]); off += 256)
                ctx->perHalo[i].push_back({offset[i] + off, static_cast<int32_t>(std::min<int64_t>(256, cnt[i] - off)),
                                           static_cast<int32_t>(i)});
    }
    std::vector<Piece> all;
    for (auto &v : ctx->perHalo) all.insert(all.end(), v.begin(), v.end());

    // raw hits per (halo, ring) and largest rho per halo
    ctx->timer.begin(ST_HITS, st);
    CU(ctx->dPieces.upload(all, st));
    CU(launch_hit_count(ctx->dPieces.p, static_cast<int>(all.size()), ctx->dCands.p, ctx->dHalos.p,
                        ctx->dNorms.p, R, ctx->dHitCount.p, ctx->dRhoMax.p, st));
    ctx->stats.launches++;
    ctx->timer.end(ST_HITS, st);
    CU(cudaStreamSynchronize(st));   // `all` is a local vector
    return SFK_OK;
}

// Fixed-point scale per halo and the processing order.  Record offsets use this
// GPU's own raw counts; the scale uses the counts of the whole halo (all GPUs
// in split mode) so that every rank quantises identically:
// |sum| <= 2 * hits_ring_max * rho_max < 2^62.
int stage_scale(sfk_ctx *ctx, bool split) {
    cudaStream_t st = ctx->stream;
    const int64_t nHalo = static_cast<int64_t>(ctx->halos.size());
    const int R = ctx->rings;
    ctx->hHitCount.assign(static_cast<size_t>(nHalo) * R, 0);
    std::vector<uint32_t> global(static_cast<size_t>(nHalo) * R);
    std::vector<unsigned long long> rhoMaxBits(nHalo);
    CU(cudaMemcpyAsync(ctx->hHitCount.data(), ctx->dHitCount.p, ctx->hHitCount.size() * sizeof(uint32_t),
                       cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(global.data(), split ? ctx->dHitCountGlobal.p : ctx->dHitCount.p,
                       global.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(rhoMaxBits.data(), ctx->dRhoMax.p, nHalo * sizeof(unsigned long long),
                       cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    ctx->hitsPerHalo.assign(nHalo, 0);
    for (int64_t i = 0; i < nHalo; i++) {
        uint32_t hmax = 0;
        for (int r = 0; r < R; r++) {
            hmax = std::max(hmax, global[i * R + r]);
            ctx->hitsPerHalo[i] += ctx->hHitCount[i * R + r];
        }
        double rhoMax;
        std::memcpy(&rhoMax, &rhoMaxBits[i], sizeof rhoMax);
        int e = 0;
        if (hmax > 0 && rhoMax > 0 && std::isfinite(rhoMax)) {
            const double bound = 2.0 * static_cast<double>(hmax) * rhoMax;
            e = 61 - std::ilogb(bound);
            e = std::max(-900, std::min(900, e));
        }
        ctx->halos[i].scale = std::ldexp(1.0, e);
        ctx->halos[i].quantum = std::ldexp(1.0, -e);
        ctx->stats.n_ring_hits += ctx->hitsPerHalo[i];
    }
    CU(ctx->dHalos.upload(ctx->halos, st));
    CU(ctx->dHaloSlot.ensure(nHalo));
    for (int64_t i = 0; i < nHalo; i++) {
        if (ctx->halos[i].valid && ctx->halos[i].owned) ctx->order.push_back(static_cast<int32_t>(i));
        else if (ctx->halos[i].valid) ctx->hStatus[i] = SFK_HALO_NOT_OWNED;
        else ctx->hStatus[i] = SFK_HALO_SKIPPED;
    }
    return SFK_OK;
}

// Halo.Insert for one batch of halos: ring-frame records, then the binning kernel.
int stage_bin(sfk_ctx *ctx, const std::vector<int32_t> &batch) {
    cudaStream_t st = ctx->stream;
    const int64_t nHalo = static_cast<int64_t>(ctx->halos.size());
    const int R = ctx->rings, n = ctx->spokes, B = ctx->bins;
    const size_t rn = static_cast<size_t>(R) * n, rnb = rn * B;
    const int nb = static_cast<int>(batch.size());
    std::vector<int32_t> haloSlot(nHalo, -1);
    std::vector<int64_t> hitOffset(static_cast<size_t>(nb) * R + 1);   // + 1: end of the last list
    std::vector<Piece> pieces;
    int64_t run = 0;
    for (int s = 0; s < nb; s++) {
        const int32_t hIdx = batch[s];
        haloSlot[hIdx] = s;
        for (int r = 0; r < R; r++) {
            hitOffset[static_cast<size_t>(s) * R + r] = run;
            run += ctx->hHitCount[static_cast<size_t>(hIdx) * R + r];
        }
        pieces.insert(pieces.end(), ctx->perHalo[hIdx].begin(), ctx->perHalo[hIdx].end());
    }
    hitOffset[static_cast<size_t>(nb) * R] = run;
    CU(ctx->dRecs.ensure(std::max<int64_t>(run, 1)));
    CU(ctx->dGrid.ensure(static_cast<size_t>(nb) * rnb));
    CU(ctx->dHitCursor.ensure(static_cast<size_t>(nb) * R * 2));   // records written from the front / from the back
    CU(ctx->dFR.ensure(static_cast<size_t>(nb) * rn));
    CU(ctx->dFIdx.ensure(static_cast<size_t>(nb) * rn));
    CU(ctx->dFCount.ensure(static_cast<size_t>(nb) * R));
    CU(ctx->dPieces.upload(pieces, st));
    CU(ctx->dBatchHalos.upload(batch, st));
    CU(ctx->dHaloSlot.upload(haloSlot, st));
    CU(ctx->dHitOffset.upload(hitOffset, st));
    CU(cudaMemsetAsync(ctx->dHitCursor.p, 0, static_cast<size_t>(nb) * R * 2 * sizeof(uint32_t), st));

    ctx->timer.begin(ST_HITS, st);
    CU(launch_hits(ctx->dPieces.p, static_cast<int>(pieces.size()), ctx->dCands.p, ctx->dHalos.p,
                   ctx->dNorms.p, ctx->dRots.p, ctx->dCos.p, ctx->dSin.p, R, n, ctx->tab.dPhi,
                   ctx->dHitOffset.p, ctx->dHitCursor.p, ctx->dHaloSlot.p, ctx->dRecs.p,
                   !(ctx->p.flags & SFK_FLAG_NO_PRUNE), st));
    ctx->stats.launches++;
    ctx->timer.end(ST_HITS, st);

    BinArgs ba{};
    ba.recs = ctx->dRecs.p; ba.hitOffset = ctx->dHitOffset.p; ba.hitCount = ctx->dHitCursor.p;
    ba.halos = ctx->dHalos.p; ba.batchHalos = ctx->dBatchHalos.p;
    ba.ringCos = ctx->dCos.p; ba.ringSin = ctx->dSin.p;
    ba.grid = ctx->dGrid.p; ba.nIntersections = ctx->dNInter.p;
    ba.tapInsert = ctx->dTapInsert.p; ba.tapStart = ctx->dTapStart.p; ba.tapEnd = ctx->dTapEnd.p;
    ba.rings = R; ba.spokes = n; ba.bins = B;
    ba.binInvG = ctx->dBinInvG.p;
    ba.invDr60 = 1.0 / (60.0 * ctx->tab.dr0);
    ba.cK = static_cast<float>(0.6931471805599453 / ctx->tab.dr0);
    // split long hit lists when there are too few CTAs to fill 148 SMs (x 2 CTAs of a 32-LoS tile,
    // x 1 of a 64-LoS tile) for a few waves
    const int tile = bin_tile_los(B, ctx->taps);
    const int64_t ctas = static_cast<int64_t>((n + tile - 1) / tile) * R * nb;
    const int64_t want = 148 * (64 / tile) * 4;
    ba.chunks = static_cast<int>(std::max<int64_t>(1, std::min<int64_t>(64, (want + ctas - 1) / ctas)));
    constexpr unsigned kOvfCap = 1u << 16;
    if (!ctx->dOvfCount.p) {
        CU(ctx->dOvfList.ensure(2 * kOvfCap));
        CU(ctx->dOvfCount.ensure(2));
        CU(cudaMemsetAsync(ctx->dOvfCount.p, 0, 2 * sizeof(unsigned), st));
    }
    ba.ovfList = ctx->dOvfList.p; ba.ovfCount = ctx->dOvfCount.p; ba.ovfCap = kOvfCap;
    // partial tiles are added with RED.64 and need a zeroed grid; whole tiles are stored
    if (!bin_stores_whole_grid(ba))
        CU(cudaMemsetAsync(ctx->dGrid.p, 0, static_cast<size_t>(nb) * rnb * sizeof(unsigned long long), st));
    ctx->timer.begin(ST_BIN, st);
    CU(launch_bin(ba, nb, ctx->taps, st));
    ctx->stats.launches++; ctx->stats.bin_launches++;
    ctx->timer.end(ST_BIN, st);
    if (ctx->keepGrid)
        for (int s = 0; s < nb; s++)
            CU(cudaMemcpyAsync(ctx->dGridTap.p + static_cast<size_t>(batch[s]) * rnb, ctx->dGrid.p + static_cast<size_t>(s) * rnb,
                               rnb * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    // host vectors of this batch are read by async copies: drain before they die
    CU(cudaStreamSynchronize(st));
    return SFK_OK;
}

int stage_fit(sfk_ctx *ctx, int nb);

// haloAnalysis for the batch whose grid is resident: RingBuffer.Splashback,
// FilterPoints, PennaVolumeFit (cmd/shell.go:613-627).  With rowCount >= 0 only the (slot,
// ring) rows [rowBegin, rowBegin + rowCount) are analysed, R_sp goes to the row-indexed
// scratch and the fit is left to the caller (ring-partitioned split-halo analysis).
int stage_analyze(sfk_ctx *ctx, int nb, int rowBegin = 0, int rowCount = -1) {
    cudaStream_t st = ctx->stream;
    const int R = ctx->rings, n = ctx->spokes, B = ctx->bins;
    const bool part = rowCount >= 0;
    if (!part) { rowBegin = 0; rowCount = nb * R; }
    if (ctx->p.percentile_profile) {
        // calcPercentile replaces calcCoeffs, cmd/shell.go:515-517
        PercentileArgs pa{};
        pa.grid = ctx->dGrid.p; pa.halos = ctx->dHalos.p; pa.batchHalos = ctx->dBatchHalos.p;
        pa.rows = ctx->dCoeffs.p; pa.nLos = static_cast<int64_t>(R) * n; pa.bins = B;
        // msort.Percentile, median.go:15-21
        long long nth = static_cast<long long>(ctx->p.percentile / 100 * static_cast<double>(pa.nLos));
        if (nth == 0) nth++;
        pa.nth = nth;
        ctx->timer.begin(ST_LOS, st);
        CU(launch_percentile(pa, nb, st));
        ctx->stats.launches += 2;
        ctx->timer.end(ST_LOS, st);
        return SFK_OK;
    }
    LosArgs la{};
    la.grid = ctx->dGrid.p; la.halos = ctx->dHalos.p; la.batchHalos = ctx->dBatchHalos.p;
    la.valTaps = ctx->dTaps0.p; la.derivTaps = ctx->dTaps1.p; la.logTab = ctx->dLogTab.p;
    la.rsp = part ? ctx->dRspRows.p : ctx->dRsp.p; la.profiles = ctx->taps ? ctx->dProfiles.p : nullptr;
    la.rings = R; la.spokes = n; la.bins = B; la.window = static_cast<int>(ctx->p.smoothing_window);
    la.dLim = ctx->p.los_slope_cutoff;
    la.losBegin = static_cast<int64_t>(rowBegin) * n; la.losCount = static_cast<int64_t>(rowCount) * n;
    la.rspBySlot = part ? 1 : 0;
    CU(ctx->dFlatList.ensure(static_cast<size_t>(std::max<int64_t>(la.losCount, 1))));
    CU(ctx->dFlatCount.ensure(2));
    la.flatList = ctx->dFlatList.p; la.flatCount = ctx->dFlatCount.p; la.flatTotal = ctx->dFlatCount.p + 1; la.listMode = 0;
    if (!ctx->flatTotalArmed) {
        CU(cudaMemsetAsync(ctx->dFlatCount.p + 1, 0, sizeof(unsigned long long), st));
        ctx->flatTotalArmed = true;
    }
    // Sliding-moment smoothing only when every ln(rho) is finite and normal (rho_bg > 0; with
    // Gadget-2's MinMass() == 0 the reference's -Inf / NaN propagation must stay sample-exact).
    bool moments = !(ctx->p.flags & SFK_FLAG_LOS_TAPLOOP) && ctx->sgReady && los_moments_supported(la);
    for (const HaloDev &h : ctx->halos)
        if (h.valid && !(h.defaultRho >= 1e-290 && h.defaultRho <= 1e290 && h.quantum >= 1e-290)) moments = false;
    ctx->timer.begin(ST_LOS, st);
    CU(launch_los(la, moments ? &ctx->sg : nullptr, nb, st));
    ctx->stats.launches += moments ? 2 : 1;
    ctx->timer.end(ST_LOS, st);

    FilterArgs fa{};
    fa.rsp = la.rsp; fa.halos = ctx->dHalos.p; fa.batchHalos = ctx->dBatchHalos.p;
    fa.rowBegin = rowBegin; fa.rowCount = rowCount; fa.rspBySlot = la.rspBySlot;
    fa.fR = ctx->dFR.p; fa.fIdx = ctx->dFIdx.p; fa.fCount = ctx->dFCount.p; fa.status = ctx->dStatus.p;
    fa.rings = R; fa.spokes = n; fa.levels = static_cast<int>(ctx->p.levels);
    fa.eta = ctx->p.eta; fa.ringPhi = ctx->dPhi.p;
    ctx->timer.begin(ST_FILTER, st);
    CU(launch_filter(fa, nb, st));
    ctx->stats.launches++;
    ctx->timer.end(ST_FILTER, st);
    return part ? SFK_OK : stage_fit(ctx, nb);
}

// PennaVolumeFit for the nb halos of the batch from their filtered points.
int stage_fit(sfk_ctx *ctx, int nb) {
    cudaStream_t st = ctx->stream;
    const int R = ctx->rings, n = ctx->spokes;
    FitArgs ft{};
    ft.fR = ctx->dFR.p; ft.fIdx = ctx->dFIdx.p; ft.fCount = ctx->dFCount.p;
    ft.batchHalos = ctx->dBatchHalos.p; ft.irots = ctx->dIrots.p;
    ft.ringCos = ctx->dCos.p; ft.ringSin = ctx->dSin.p;
    ft.coeffs = ctx->dCoeffs.p; ft.status = ctx->dStatus.p; ft.nPoints = ctx->dNPoints.p;
    ft.rings = R; ft.spokes = n; ft.order = static_cast<int>(ctx->p.order);
    size_t nRows, nGram, nInv, nCs;
    CU(ctx->dFitScratch.ensure(fit_scratch_doubles(nb, R, n, ft.order, &nRows, &nGram, &nInv, &nCs)));
    ft.rows = ctx->dFitScratch.p; ft.gramPart = ft.rows + nRows; ft.gramInv = ft.gramPart + nGram;
    ft.csPart = ft.gramInv + nInv;
    ctx->timer.begin(ST_FIT, st);
    CU(launch_fit(ft, nb, st));
    ctx->stats.launches += 4;
    ctx->timer.end(ST_FIT, st);
    return SFK_OK;
}

int stage_collect(sfk_ctx *ctx, double *coeffs, int32_t *status) {
    cudaStream_t st = ctx->stream;
    const int64_t nHalo = static_cast<int64_t>(ctx->halos.size());
    const int P2 = ctx->P2;
    std::vector<int32_t> dstat(nHalo);
    std::vector<unsigned long long> nInter(nHalo);
    std::vector<long long> nPts(nHalo);
    unsigned long long flatTotal = 0;
    if (ctx->flatTotalArmed)
        CU(cudaMemcpyAsync(&flatTotal, ctx->dFlatCount.p + 1, sizeof flatTotal, cudaMemcpyDeviceToHost, st));
    unsigned ovf[2] = {0, 0};
    if (ctx->dOvfCount.p) CU(cudaMemcpyAsync(ovf, ctx->dOvfCount.p, sizeof ovf, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ctx->hCoeffs.data(), ctx->dCoeffs.p, ctx->hCoeffs.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(dstat.data(), ctx->dStatus.p, nHalo * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(nInter.data(), ctx->dNInter.p, nHalo * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(nPts.data(), ctx->dNPoints.p, nHalo * sizeof(long long), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (ovf[1] != 0) {
        CU(cudaMemsetAsync(ctx->dOvfCount.p, 0, 2 * sizeof(unsigned), st));
        return fail(ctx, SFK_E_UNSUPPORTED, "bin_kernel: side list of edges on rMax exhausted");
    }
    ctx->interPerHalo.assign(nInter.begin(), nInter.end());
    for (int64_t i = 0; i < nHalo; i++) {
        if (ctx->hStatus[i] == SFK_HALO_OK) ctx->hStatus[i] = dstat[i];
        if (ctx->hStatus[i] != SFK_HALO_OK)
            for (int k = 0; k < P2; k++) ctx->hCoeffs[i * P2 + k] = 0.0;
        else if (ctx->p.percentile_profile)   // Halo.GetRs, halo.go:360-370
            for (int k = 0; k < ctx->bins; k++)
                ctx->hCoeffs[i * P2 + k] = std::exp(ctx->halos[i].lrMin + ctx->halos[i].dlr * (static_cast<double>(k) + 0.5));
        ctx->stats.n_intersections += static_cast<int64_t>(nInter[i]);
        ctx->stats.n_points += nPts[i];
    }
    ctx->stats.n_candidates = ctx->candCount;
    ctx->stats.n_los_literal = static_cast<int64_t>(flatTotal);
    ctx->flatTotalArmed = false;
    ctx->stats.ms_cull = ctx->timer.collect(ST_CULL);
    ctx->stats.ms_hits = ctx->timer.collect(ST_HITS);
    ctx->stats.ms_bin = ctx->timer.collect(ST_BIN);
    ctx->stats.ms_los = ctx->timer.collect(ST_LOS);
    ctx->stats.ms_filter = ctx->timer.collect(ST_FILTER);
    ctx->stats.ms_fit = ctx->timer.collect(ST_FIT);
    if (coeffs) std::memcpy(coeffs, ctx->hCoeffs.data(), ctx->hCoeffs.size() * sizeof(double));
    if (status) std::memcpy(status, ctx->hStatus.data(), nHalo * sizeof(int32_t));
    ctx->finished = true;
    ctx->splitPhase = 0;
    return SFK_OK;
}

}  // namespace

extern "C" {


int sfk_device_count(void) {
    int n = 0;
    return cudaGetDeviceCount(&n) == cudaSuccess ? n : 0;
}

void sfk_default_params(sfk_params *p) {
    std::memset(p, 0, sizeof *p);
    p->subsample_factor = 1; p->radial_bins = 256; p->spokes = 256; p->rings = 100;
    p->r_max_mult = 3; p->r_min_mult = 0.3; p->r_kernel_mult = 0.2; p->eta = 10;
    p->order = 3; p->levels = 3; p->smoothing_window = 121; p->los_slope_cutoff = 0.0;
    p->background_rho_mult = 0.5; p->percentile_profile = 0; p->percentile = 50.0;
    p->seed = 0; p->device = 0; p->flags = 0;
}

static int create_impl(sfk_ctx *ctx, const sfk_params *params);

int sfk_create(sfk_ctx **out, const sfk_params *params) {
    if (!out || !params) return SFK_E_INVALID;
    *out = nullptr;
    sfk_ctx *ctx = new (std::nothrow) sfk_ctx();
    if (!ctx) return SFK_E_NOMEM;
    *out = ctx;  // returned even on failure so the caller can read the error
    return create_impl(ctx, params);
}

static int create_impl(sfk_ctx *ctx, const sfk_params *params) try {
    ctx->p = *params;
    int rc = validate(ctx, *params);
    if (rc) return rc;
    int nDev = 0;
    if (cudaGetDeviceCount(&nDev) != cudaSuccess || nDev <= 0)
        return fail(ctx, SFK_E_NODEVICE, "no CUDA device: this library has no CPU path");
    if (params->device < 0 || params->device >= nDev) return fail(ctx, SFK_E_INVALID, "bad device ordinal");
    CU(cudaSetDevice(params->device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, params->device));
    if (prop.major != 10)
        return fail(ctx, SFK_E_NODEVICE, "kernels are built for sm_100a only (found sm_" +
                                             std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    CU(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CU(cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
        CU(cudaEventCreateWithFlags(&ctx->evCopied[k], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->evCulled[k], cudaEventDisableTiming));
    }
    ctx->rings = static_cast<int>(params->rings);
    ctx->spokes = static_cast<int>(params->spokes);
    ctx->bins = static_cast<int>(params->radial_bins);
    // row length of the result table: Penna coefficients, or radii + densities
    // (calcPercentile returns 2*RadialBins values, cmd/shell.go:632)
    ctx->P2 = params->percentile_profile ? static_cast<int>(2 * params->radial_bins)
                                         : static_cast<int>(2 * params->order * params->order);
    ctx->taps = (params->flags & SFK_FLAG_TAPS) != 0;
    ctx->keepGrid = (params->flags & SFK_FLAG_KEEP_GRID) != 0;
    build_tables(*params, ctx->tab);
    CU(ctx->dNorms.upload(ctx->tab.norms, ctx->stream));
    CU(ctx->dRots.upload(ctx->tab.rots, ctx->stream));
    CU(ctx->dIrots.upload(ctx->tab.irots, ctx->stream));
    CU(ctx->dCos.upload(ctx->tab.ringCos, ctx->stream));
    CU(ctx->dSin.upload(ctx->tab.ringSin, ctx->stream));
    CU(ctx->dPhi.upload(ctx->tab.ringPhi, ctx->stream));
    CU(ctx->dBinInvG.upload(ctx->tab.binInvG, ctx->stream));
    std::vector<double> logTab;
    build_log_table(logTab);
    CU(ctx->dLogTab.upload(logTab, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SFK_OK;
} SFK_CATCH(ctx)

void sfk_destroy(sfk_ctx *ctx) {
    if (!ctx) return;
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm) {
        if (const NcclApi *api = nccl_api(nullptr)) api->CommDestroy(ctx->comm);
        ctx->comm = nullptr;
    }
    if (ctx->copyStream) { cudaStreamSynchronize(ctx->copyStream); cudaStreamDestroy(ctx->copyStream); }
    for (int k = 0; k < 2; k++) {
        if (ctx->evCopied[k]) cudaEventDestroy(ctx->evCopied[k]);
        if (ctx->evCulled[k]) cudaEventDestroy(ctx->evCulled[k]);
    }
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *sfk_last_error(const sfk_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int sfk_begin_snapshot(sfk_ctx *ctx, const sfk_cosmo *cosmo, float min_mass) try {
    if (!ctx || !cosmo) return SFK_E_INVALID;
    CU(cudaSetDevice(ctx->p.device));
    ctx->cosmo0 = *cosmo;
    ctx->minMass = min_mass;
    ctx->halos.clear(); ctx->ox.clear(); ctx->oy.clear(); ctx->oz.clear();
    ctx->candPerHalo.clear();
    ctx->candCount = 0; ctx->candUpper = 0; ctx->pushed = false;
    if (ctx->cullOpen) { ctx->timer.end(ST_CULL, ctx->stream); ctx->cullOpen = false; ctx->timer.collect(ST_CULL); }
    CU(ctx->dCandTotal.ensure(1));
    CU(cudaMemsetAsync(ctx->dCandTotal.p, 0, sizeof(unsigned long long), ctx->stream));
    ctx->inSnapshot = true; ctx->finished = false; ctx->splitPhase = 0;
    ctx->stats = sfk_stats{};
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_add_halos(sfk_ctx *ctx, int64_t n, const double *x, const double *y, const double *z,
                  const double *r200m) try {
    if (!ctx || !ctx->inSnapshot || n < 0) return fail(ctx, SFK_E_INVALID, "sfk_add_halos: no snapshot open");
    if (!ctx->halos.empty()) return fail(ctx, SFK_E_INVALID, "sfk_add_halos: halos already set");
    const sfk_params &p = ctx->p;
    const int64_t sf = p.subsample_factor;
    const double rhoM = rho_average(ctx->cosmo0.h100 * 100, ctx->cosmo0.omega_m, ctx->cosmo0.omega_l, ctx->cosmo0.z);
    ctx->halos.resize(n);
    ctx->ox.assign(x, x + n); ctx->oy.assign(y, y + n); ctx->oz.assign(z, z + n);
    for (int64_t i = 0; i < n; i++) {
        HaloDev h{};
        const double r = r200m[i];
        h.valid = r > 0 ? 1 : 0;   // cmd/shell.go:535-538
        h.owned = 1;
        if (h.valid) {
            // createHalos, cmd/shell.go:540-556
            const double rMax = r * p.r_max_mult, rMin = r * p.r_min_mult;
            const double rad0 = r * p.r_kernel_mult;
            const double vol0 = kFourPiOver3 * rad0 * rad0 * rad0;
            const double rho = ((static_cast<double>(ctx->minMass) * static_cast<double>(sf * sf * sf)) / vol0) / rhoM;
            h.defaultRho = rho * p.background_rho_mult;
            h.rMax = rMax; h.rMin = rMin; h.invRMin = 1.0 / rMin;
            h.x0 =static_cast<float>(x[i]); h.y0 = static_cast<float>(y[i]); h.z0 = static_cast<float>(z[i]);
            // loadSphereVecs, cmd/shell.go:442 and Halo.Intersect, halo.go:148-152
            h.rad = rMax * p.r_kernel_mult / p.r_max_mult;
            h.rad2 = h.rad * h.rad;
            h.sphVol = kFourPiOver3 * h.rad * h.rad * h.rad;
            double lo = rMin - h.rad, hi = rMax + h.rad;
            if (lo < 0) lo = 0;
            h.lo2 = static_cast<float>(lo * lo); h.hi2 = static_cast<float>(hi * hi);
            // ProfileRing.Init, profile_ring.go:75-82; GetRs, halo.go:365-366
            h.lowR = std::log(rMin); h.highR = std::log(rMax);
            h.dr = (h.highR - h.lowR) / static_cast<double>(ctx->bins);
            h.lrMin = std::log(rMin);
            h.dlr = (std::log(rMax) - std::log(rMin)) / static_cast<double>(ctx->bins);
            h.scale = 1.0; h.quantum = 1.0;
        }
        ctx->halos[i] = h;
    }
    ctx->candPerHalo.assign(n, 0);
    CU(ctx->dHalos.upload(ctx->halos, ctx->stream));
    CU(ctx->dHaloCandCount.ensure(std::max<int64_t>(n, 1)));
    CU(cudaMemsetAsync(ctx->dHaloCandCount.p, 0, std::max<int64_t>(n, 1) * sizeof(uint32_t), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->stats.n_halos = n;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_set_owned(sfk_ctx *ctx, int64_t n, const uint8_t *owned) try {
    if (!ctx || !ctx->inSnapshot || !owned || n != static_cast<int64_t>(ctx->halos.size()))
        return fail(ctx, SFK_E_INVALID, "sfk_set_owned: call after sfk_add_halos with one flag per halo");
    if (ctx->pushed) return fail(ctx, SFK_E_INVALID, "sfk_set_owned: blocks were already pushed");
    for (int64_t i = 0; i < n; i++) ctx->halos[i].owned = owned[i] ? 1 : 0;
    CU(ctx->dHalos.upload(ctx->halos, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_block_halo_mask(sfk_ctx *ctx, const float origin[3], const float width[3],
                        double total_width, uint8_t *hit, int64_t *n_hit) try {
    if (!ctx || !ctx->inSnapshot) return fail(ctx, SFK_E_INVALID, "sfk_block_halo_mask: no snapshot open");
    int64_t cnt = 0;
    for (size_t i = 0; i < ctx->halos.size(); i++) {
        const bool h = ctx->halos[i].valid && sheet_intersect(ctx, i, origin, width, total_width);
        if (hit) hit[i] = h;
        cnt += h;
    }
    if (n_hit) *n_hit = cnt;
    return SFK_OK;
} SFK_CATCH(ctx)

// Host pushes: the copy runs on its own stream into one of two staging buffers, so it overlaps
// the cull of the previous block.  wait: return once the caller's buffers have been consumed
// (cgo may not retain Go pointers; io.VectorBuffer reuses its slices) -- not once the block has
// been culled; otherwise the copy is only queued (sfk_push_block_async).
static int push_block_host(sfk_ctx *ctx, const char *who, const float *xyz, const float *mass, int64_t n,
                           const sfk_cosmo *cosmo, double total_width, const float origin[3],
                           const float width[3], bool wait) {
    if (!ctx || !ctx->inSnapshot || ctx->finished || !cosmo || n < 0 || (n > 0 && !xyz))
        return fail(ctx, SFK_E_INVALID, std::string(who) + ": bad arguments or no snapshot open");
    CU(cudaSetDevice(ctx->p.device));
    std::vector<int32_t> &list = ctx->blockList;
    int64_t nOwned = 0;
    block_halo_list(ctx, origin, width, total_width, list, nOwned);
    if (list.empty() || n == 0) return SFK_OK;   // sphereLoop skips the read, cmd/shell.go:389-391
    const int slot = static_cast<int>(ctx->pushIdx++ & 1);
    if (ctx->culledValid[slot]) CU(cudaStreamWaitEvent(ctx->copyStream, ctx->evCulled[slot], 0));
    if (static_cast<size_t>(3 * n) > ctx->dXyz[slot].cap || (mass && static_cast<size_t>(n) > ctx->dMass[slot].cap)) {
        if (ctx->culledValid[slot]) CU(cudaEventSynchronize(ctx->evCulled[slot]));   // the old buffer is still being read
        CU(ctx->dXyz[slot].ensure(3 * n));
        if (mass) CU(ctx->dMass[slot].ensure(n));
    }
    CU(cudaMemcpyAsync(ctx->dXyz[slot].p, xyz, 3 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->copyStream));
    if (mass) CU(cudaMemcpyAsync(ctx->dMass[slot].p, mass, n * sizeof(float), cudaMemcpyHostToDevice, ctx->copyStream));
    CU(cudaEventRecord(ctx->evCopied[slot], ctx->copyStream));
    if (wait) CU(cudaStreamSynchronize(ctx->copyStream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->evCopied[slot], 0));
    int rc = push_block_impl(ctx, ctx->dXyz[slot].p, mass ? ctx->dMass[slot].p : nullptr, n, cosmo, total_width, list, nOwned);
    if (rc) return rc;
    CU(cudaEventRecord(ctx->evCulled[slot], ctx->stream));
    ctx->culledValid[slot] = true;
    return SFK_OK;
}

int sfk_push_block(sfk_ctx *ctx, const float *xyz, const float *mass, int64_t n,
                   const sfk_cosmo *cosmo, double total_width, const float origin[3],
                   const float width[3]) try {
    return push_block_host(ctx, "sfk_push_block", xyz, mass, n, cosmo, total_width, origin, width, true);
} SFK_CATCH(ctx)

int sfk_push_block_async(sfk_ctx *ctx, const float *xyz, const float *mass, int64_t n,
                         const sfk_cosmo *cosmo, double total_width, const float origin[3],
                         const float width[3]) try {
    return push_block_host(ctx, "sfk_push_block_async", xyz, mass, n, cosmo, total_width, origin, width, false);
} SFK_CATCH(ctx)

int sfk_push_sync(sfk_ctx *ctx) try {
    if (!ctx) return SFK_E_INVALID;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaStreamSynchronize(ctx->copyStream));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_push_block_device(sfk_ctx *ctx, const float *d_xyz, const float *d_mass, int64_t n,
                          const sfk_cosmo *cosmo, double total_width, const float origin[3],
                          const float width[3]) try {
    if (!ctx || !ctx->inSnapshot || ctx->finished || !cosmo || n < 0 || (n > 0 && !d_xyz))
        return fail(ctx, SFK_E_INVALID, "sfk_push_block_device: bad arguments or no snapshot open");
    CU(cudaSetDevice(ctx->p.device));
    // The caller's buffers were filled on other streams: everything queued on the device before
    // the FIRST device push of a snapshot is waited for once; later buffers must be complete when
    // they are handed over.  The cull of the block is queued, not finished, on return.
    if (!ctx->pushed) CU(cudaDeviceSynchronize());
    int64_t nOwned = 0;
    block_halo_list(ctx, origin, width, total_width, ctx->blockList, nOwned);
    return push_block_impl(ctx, d_xyz, d_mass, n, cosmo, total_width, ctx->blockList, nOwned);
} SFK_CATCH(ctx)

int64_t sfk_row_length(const sfk_ctx *ctx) { return ctx ? ctx->P2 : 0; }

int sfk_finish_snapshot(sfk_ctx *ctx, double *coeffs, int32_t *status) try {
    if (!ctx || !ctx->inSnapshot) return fail(ctx, SFK_E_INVALID, "sfk_finish_snapshot: no snapshot open");
    if (ctx->splitPhase != 0) return fail(ctx, SFK_E_INVALID, "sfk_finish_snapshot: a split-halo pass is open");
    CU(cudaSetDevice(ctx->p.device));
    int rc = stage_prepare(ctx);
    if (rc) return rc;
    if (ctx->halos.empty()) { ctx->finished = true; return SFK_OK; }
    rc = stage_scale(ctx, false);
    if (rc) return rc;
    // batches of halos in input order, bounded by hit-record and grid memory
    // 256 halos in flight: 13 GB of bin grid + up to 24 GB of hit records out of 180 GB; fewer,
    // larger batches mean fewer kernel tails and fewer latency-bound fit launches per snapshot
    const int64_t maxBatch = 256;
    const int64_t hitBudget = (int64_t(24) << 30) / static_cast<int64_t>(sizeof(HitRec));
    size_t pos = 0;
    while (pos < ctx->order.size()) {
        std::vector<int32_t> batch;
        int64_t hits = 0;
        while (pos < ctx->order.size() && static_cast<int64_t>(batch.size()) < maxBatch) {
            const int64_t hh = ctx->hitsPerHalo[ctx->order[pos]];
            if (!batch.empty() && hits + hh > hitBudget) break;
            batch.push_back(ctx->order[pos++]);
            hits += hh;
        }
        rc = stage_bin(ctx, batch);
        if (rc) return rc;
        rc = stage_analyze(ctx, static_cast<int>(batch.size()));
        if (rc) return rc;
    }
    return stage_collect(ctx, coeffs, status);
} SFK_CATCH(ctx)

// ---- split-halo mode: one halo's particles spread over several GPUs -----------

int sfk_split_count(sfk_ctx *ctx) try {
    if (!ctx || !ctx->inSnapshot || ctx->splitPhase != 0)
        return fail(ctx, SFK_E_INVALID, "sfk_split_count: no snapshot open or split pass already started");
    CU(cudaSetDevice(ctx->p.device));
    int rc = stage_prepare(ctx);
    if (rc) return rc;
    const size_t nc = ctx->halos.size() * static_cast<size_t>(ctx->rings);
    CU(ctx->dHitCountGlobal.ensure(nc));
    if (nc) CU(cudaMemcpyAsync(ctx->dHitCountGlobal.p, ctx->dHitCount.p, nc * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->splitPhase = 1;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_split_bin(sfk_ctx *ctx) try {
    if (!ctx || ctx->splitPhase != 1) return fail(ctx, SFK_E_INVALID, "sfk_split_bin: call sfk_split_count first");
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaDeviceSynchronize());   // the caller's collectives ran on other streams
    int rc = stage_scale(ctx, true);
    if (rc) return rc;
    const size_t rnb = static_cast<size_t>(ctx->rings) * ctx->spokes * ctx->bins;
    if (ctx->order.size() * rnb * sizeof(unsigned long long) > (size_t(96) << 30))
        return fail(ctx, SFK_E_NOMEM, "sfk_split_bin: the bin grids of all halos must fit in HBM at once");
    rc = stage_bin(ctx, ctx->order);
    if (rc) return rc;
    ctx->splitPhase = 2;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_split_buffer(sfk_ctx *ctx, int32_t which, void **d_ptr, int64_t *count) try {
    if (!ctx || !d_ptr || !count) return SFK_E_INVALID;
    const int64_t nHalo = static_cast<int64_t>(ctx->halos.size());
    switch (which) {
    case SFK_SPLIT_HIT_COUNTS:
        if (ctx->splitPhase < 1) break;
        *d_ptr = ctx->dHitCountGlobal.p; *count = nHalo * ctx->rings; return SFK_OK;
    case SFK_SPLIT_RHO_MAX:
        if (ctx->splitPhase < 1) break;
        *d_ptr = ctx->dRhoMax.p; *count = nHalo; return SFK_OK;
    case SFK_SPLIT_GRID:
        if (ctx->splitPhase < 2) break;
        *d_ptr = ctx->dGrid.p;
        *count = static_cast<int64_t>(ctx->order.size()) * ctx->rings * ctx->spokes * ctx->bins;
        return SFK_OK;
    default: break;
    }
    return fail(ctx, SFK_E_INVALID, "sfk_split_buffer: buffer not available in this phase");
} SFK_CATCH(ctx)

int sfk_split_finish(sfk_ctx *ctx, double *coeffs, int32_t *status) try {
    if (!ctx || ctx->splitPhase != 2) return fail(ctx, SFK_E_INVALID, "sfk_split_finish: call sfk_split_bin first");
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaDeviceSynchronize());   // the caller's all-reduce of the grid
    int rc = stage_analyze(ctx, static_cast<int>(ctx->order.size()));
    if (rc) return rc;
    return stage_collect(ctx, coeffs, status);
} SFK_CATCH(ctx)

// ---- split-halo mode with the exchange inside the library (NCCL) ------------------------

#define NC(call)                                                                             \
    do {                                                                                     \
        ncclResult_t r_ = (call);                                                            \
        if (r_ != ncclSuccess)                                                               \
            return fail(ctx, SFK_E_CUDA, std::string(#call) + ": " + api->GetErrorString(r_)); \
    } while (0)

int sfk_comm_unique_id(void *id128) {
    std::string err;
    const NcclApi *api = nccl_api(&err);
    if (!api || !id128) return SFK_E_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == SFK_COMM_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    if (api->GetUniqueId(&id) != ncclSuccess) return SFK_E_CUDA;
    std::memcpy(id128, &id, sizeof id);
    return SFK_OK;
}

int sfk_comm_init(sfk_ctx *ctx, const void *id128, int32_t rank, int32_t nranks) try {
    if (!ctx || !id128 || nranks < 1 || rank < 0 || rank >= nranks)
        return fail(ctx, SFK_E_INVALID, "sfk_comm_init: bad arguments");
    if (ctx->comm) return fail(ctx, SFK_E_INVALID, "sfk_comm_init: communicator already set");
    std::string err;
    const NcclApi *api = nccl_api(&err);
    if (!api) return fail(ctx, SFK_E_UNSUPPORTED, "sfk_comm_init: " + err);
    CU(cudaSetDevice(ctx->p.device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    NC(api->CommInitRank(&ctx->comm, nranks, id, rank));
    ctx->commRank = rank; ctx->commSize = nranks;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_split_finish_comm(sfk_ctx *ctx, double *coeffs, int32_t *status) try {
    if (!ctx || !ctx->inSnapshot || ctx->splitPhase != 0)
        return fail(ctx, SFK_E_INVALID, "sfk_split_finish_comm: no snapshot open or a split pass is already open");
    if (!ctx->comm) return fail(ctx, SFK_E_INVALID, "sfk_split_finish_comm: call sfk_comm_init first");
    if (ctx->p.percentile_profile)
        return fail(ctx, SFK_E_UNSUPPORTED, "sfk_split_finish_comm: PercentileProfile is not supported in split-halo mode");
    const NcclApi *api = nccl_api(nullptr);
    cudaStream_t st = ctx->stream;
    const int G = ctx->commSize, g = ctx->commRank;
    CU(cudaSetDevice(ctx->p.device));
    // raw hit counts of this rank's particles; the halo-wide counts and rho_max fix the
    // fixed-point scale, identically on every rank
    int rc = stage_prepare(ctx);
    if (rc) return rc;
    const int64_t nHalo = static_cast<int64_t>(ctx->halos.size());
    if (nHalo == 0) { ctx->finished = true; return SFK_OK; }
    const int R = ctx->rings, n = ctx->spokes, B = ctx->bins;
    const size_t nc = static_cast<size_t>(nHalo) * R;
    CU(ctx->dHitCountGlobal.ensure(nc));
    CU(cudaMemcpyAsync(ctx->dHitCountGlobal.p, ctx->dHitCount.p, nc * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    NC(api->GroupStart());
    NC(api->AllReduce(ctx->dHitCountGlobal.p, ctx->dHitCountGlobal.p, nc, ncclUint32, ncclSum, ctx->comm, st));
    NC(api->AllReduce(ctx->dRhoMax.p, ctx->dRhoMax.p, nHalo, ncclUint64, ncclMax, ctx->comm, st));
    NC(api->GroupEnd());
    rc = stage_scale(ctx, true);
    if (rc) return rc;
    const int nb = static_cast<int>(ctx->order.size());
    if (nb == 0) return stage_collect(ctx, coeffs, status);
    // (halo, ring) rows, padded so that every rank owns the same number
    const int rows = nb * R;
    const int rowsPer = (rows + G - 1) / G, rowsPad = rowsPer * G;
    const size_t rowCells = static_cast<size_t>(n) * B;
    if (rowsPad * rowCells * sizeof(unsigned long long) > (size_t(96) << 30))
        return fail(ctx, SFK_E_NOMEM, "sfk_split_finish_comm: the bin grids of all halos must fit in HBM at once");
    CU(ctx->dGrid.ensure(rowsPad * rowCells));
    CU(ctx->dRspRows.ensure(static_cast<size_t>(rowsPad) * n));
    CU(ctx->dFR.ensure(static_cast<size_t>(rowsPad) * n));
    CU(ctx->dFIdx.ensure(static_cast<size_t>(rowsPad) * n));
    CU(ctx->dFCount.ensure(rowsPad));
    rc = stage_bin(ctx, ctx->order);   // this rank's private grid (Halo.Split's copy)
    if (rc) return rc;
    // Halo.Join: sum the private grids; rank g receives rows [g rowsPer, (g + 1) rowsPer) in place
    NC(api->ReduceScatter(ctx->dGrid.p, ctx->dGrid.p + static_cast<size_t>(g) * rowsPer * rowCells,
                          rowsPer * rowCells, ncclInt64, ncclSum, ctx->comm, st));
    const int rowBegin = g * rowsPer, rowCount = std::max(0, std::min(rows, rowBegin + rowsPer) - rowBegin);
    CU(cudaMemsetAsync(ctx->dFCount.p + rowBegin, 0, rowsPer * sizeof(int32_t), st));
    rc = stage_analyze(ctx, nb, rowBegin, rowCount);   // RingBuffer.Splashback + FilterPoints of my rings
    if (rc) return rc;
    // every rank gets every ring's R_sp and filtered points, then fits all halos itself: the
    // fit sums in the single-GPU order, so the coefficients are bitwise those of one GPU
    NC(api->GroupStart());
    NC(api->AllGather(ctx->dRspRows.p + static_cast<size_t>(rowBegin) * n, ctx->dRspRows.p, static_cast<size_t>(rowsPer) * n,
                      ncclFloat64, ctx->comm, st));
    NC(api->AllGather(ctx->dFR.p + static_cast<size_t>(rowBegin) * n, ctx->dFR.p, static_cast<size_t>(rowsPer) * n, ncclFloat64,
                      ctx->comm, st));
    NC(api->AllGather(ctx->dFIdx.p + static_cast<size_t>(rowBegin) * n, ctx->dFIdx.p, static_cast<size_t>(rowsPer) * n, ncclInt32,
                      ctx->comm, st));
    NC(api->AllGather(ctx->dFCount.p + rowBegin, ctx->dFCount.p, rowsPer, ncclInt32, ctx->comm, st));
    NC(api->AllReduce(ctx->dStatus.p, ctx->dStatus.p, nHalo, ncclInt32, ncclMax, ctx->comm, st));
    NC(api->AllReduce(ctx->dNInter.p, ctx->dNInter.p, nHalo, ncclUint64, ncclSum, ctx->comm, st));
    NC(api->GroupEnd());
    for (int s = 0; s < nb; s++)
        CU(cudaMemcpyAsync(ctx->dRsp.p + static_cast<size_t>(ctx->order[s]) * R * n, ctx->dRspRows.p + static_cast<size_t>(s) * R * n,
                           static_cast<size_t>(R) * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    rc = stage_fit(ctx, nb);
    if (rc) return rc;
    return stage_collect(ctx, coeffs, status);
} SFK_CATCH(ctx)

// ---- `stats` mode: splashback mass --------------------------------------------------

int sfk_sphere_block_hit(const float center[3], float radius, const float origin[3], const float width[3],
                         double total_width) {
    // sphereSheetIntersect, cmd/stats.go:685-692
    return in_range(center[0], radius, origin[0], width[0], total_width) &&
           in_range(center[1], radius, origin[1], width[1], total_width) &&
           in_range(center[2], radius, origin[2], width[2], total_width);
}

int sfk_mass_begin(sfk_ctx *ctx, int64_t n_halo, int32_t order, const float *centers, const double *coeffs,
                   const double *r_low, const double *r_high) try {
    if (!ctx || n_halo < 0 || order <= 0 || (n_halo > 0 && (!centers || !coeffs || !r_low || !r_high)))
        return fail(ctx, SFK_E_INVALID, "sfk_mass_begin: bad arguments");
    if (order > kMaxOrder) return fail(ctx, SFK_E_UNSUPPORTED, "Order above 4 is not supported.");
    CU(cudaSetDevice(ctx->p.device));
    const size_t P2 = 2 * static_cast<size_t>(order) * order;
    std::vector<float> c(centers, centers + 3 * n_halo), lo(n_halo), hi(n_halo);
    for (int64_t i = 0; i < n_halo; i++) {
        // low2, high2 := float32(rLow*rLow), float32(rHigh*rHigh), cmd/stats.go:522
        lo[i] = static_cast<float>(r_low[i] * r_low[i]);
        hi[i] = static_cast<float>(r_high[i] * r_high[i]);
    }
    std::vector<double> cs(coeffs, coeffs + P2 * n_halo);
    CU(ctx->dMassCenters.upload(c, ctx->stream));
    CU(ctx->dMassLow2.upload(lo, ctx->stream));
    CU(ctx->dMassHigh2.upload(hi, ctx->stream));
    CU(ctx->dMassCoeffs.upload(cs, ctx->stream));
    CU(ctx->dMassSums.ensure(n_halo));
    CU(ctx->dMassList.ensure(n_halo + 7));
    if (n_halo) CU(cudaMemsetAsync(ctx->dMassSums.p, 0, n_halo * sizeof(double), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));   // the host vectors die here
    ctx->massHalos = n_halo; ctx->massOrder = order; ctx->massOpen = true; ctx->massDevSynced = false;
    ctx->bandOn = false;
    return SFK_OK;
} SFK_CATCH(ctx)

static int band_pass(sfk_ctx *ctx, const float *dxyz, int64_t n, double total_width);

static int mass_push_impl(sfk_ctx *ctx, const float *dxyz, const float *dmass, int64_t n, double total_width) try {
    MassArgs a{};
    a.xyz = dxyz; a.mass = dmass; a.n = n;
    a.tw2 = static_cast<float>(total_width) / 2;
    a.centers = ctx->dMassCenters.p; a.low2 = ctx->dMassLow2.p; a.high2 = ctx->dMassHigh2.p;
    a.coeffs = ctx->dMassCoeffs.p; a.masses = ctx->dMassSums.p;
    a.list = ctx->dMassList.p; a.listCount = ctx->dMassList.p + ctx->massHalos;
    a.bboxKeys = reinterpret_cast<unsigned *>(ctx->dMassList.p + ctx->massHalos + 1);
    a.nHalo = static_cast<int>(ctx->massHalos); a.order = ctx->massOrder; a.P2 = 2 * ctx->massOrder * ctx->massOrder;
    CU(launch_mass(a, ctx->stream));
    ctx->stats.launches += 3;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_mass_push_block(sfk_ctx *ctx, const float *xyz, const float *mass, int64_t n, double total_width) try {
    if (!ctx || !ctx->massOpen || n < 0 || (n > 0 && (!xyz || !mass)))
        return fail(ctx, SFK_E_INVALID, "sfk_mass_push_block: call sfk_mass_begin first");
    ctx->hBand.clear();
    if (n == 0 || ctx->massHalos == 0) return SFK_OK;
    CU(cudaSetDevice(ctx->p.device));
    CU(cudaStreamSynchronize(ctx->stream));   // a shell-pass cull may still read the staging buffer
    CU(ctx->dXyz[0].ensure(3 * n));
    CU(ctx->dMass[0].ensure(n));
    CU(cudaMemcpyAsync(ctx->dXyz[0].p, xyz, 3 * n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->dMass[0].p, mass, n * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    int rc = mass_push_impl(ctx, ctx->dXyz[0].p, ctx->dMass[0].p, n, total_width);
    if (rc) return rc;
    if (ctx->bandOn && (rc = band_pass(ctx, ctx->dXyz[0].p, n, total_width)) != 0) return rc;
    CU(cudaStreamSynchronize(ctx->stream));   // the caller may reuse its buffers (io.VectorBuffer does)
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_mass_push_block_device(sfk_ctx *ctx, const float *d_xyz, const float *d_mass, int64_t n,
                               double total_width) try {
    if (!ctx || !ctx->massOpen || n < 0 || (n > 0 && (!d_xyz || !d_mass)))
        return fail(ctx, SFK_E_INVALID, "sfk_mass_push_block_device: call sfk_mass_begin first");
    ctx->hBand.clear();
    if (n == 0 || ctx->massHalos == 0) return SFK_OK;
    CU(cudaSetDevice(ctx->p.device));
    // as sfk_push_block_device: work queued on other streams is waited for once per pass; later
    // buffers must be complete on the device when they are handed over
    if (!ctx->massDevSynced) { CU(cudaDeviceSynchronize()); ctx->massDevSynced = true; }
    int rc = mass_push_impl(ctx, d_xyz, d_mass, n, total_width);
    if (rc) return rc;
    if (ctx->bandOn) return band_pass(ctx, d_xyz, n, total_width);
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_mass_set_band(sfk_ctx *ctx, double shell_width, const float *radii, const double *r_low,
                      const double *r_high) try {
    if (!ctx || !ctx->massOpen || shell_width == 0 || (ctx->massHalos > 0 && (!radii || !r_low || !r_high)))
        return fail(ctx, SFK_E_INVALID, "sfk_mass_set_band: call sfk_mass_begin first; ShellWidth must not be 0");
    CU(cudaSetDevice(ctx->p.device));
    const int64_t n = ctx->massHalos;
    std::vector<float> lo(n), hi(n);
    std::vector<double> delta(n);
    for (int64_t i = 0; i < n; i++) {
        // appendShellParticlesChan, cmd/stats.go:554-560 (a negative rLow - delta is squared too)
        double d = static_cast<double>(radii[i]) * shell_width;
        if (shell_width < 0) d = 0;
        const double rl = r_low[i] - d, rh = r_high[i] + d;
        lo[i] = static_cast<float>(rl * rl); hi[i] = static_cast<float>(rh * rh);
        delta[i] = d;
    }
    CU(ctx->dBandLow2.upload(lo, ctx->stream));
    CU(ctx->dBandHigh2.upload(hi, ctx->stream));
    CU(ctx->dBandDelta.upload(delta, ctx->stream));
    CU(ctx->dBandCount.ensure(1));
    CU(ctx->dBandList.ensure(size_t(1) << 20));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->bandInside = shell_width < 0 ? 1 : 0;
    ctx->bandOn = true;
    ctx->hBand.clear();
    return SFK_OK;
} SFK_CATCH(ctx)

// The band pass over the block that sits in dXyz[0] (sfk_mass_push_block just put it there).
static int band_pass(sfk_ctx *ctx, const float *dxyz, int64_t n, double total_width) {
    for (;;) {
        CU(cudaMemsetAsync(ctx->dBandCount.p, 0, sizeof(unsigned long long), ctx->stream));
        MassArgs a{};
        a.xyz = dxyz; a.mass = nullptr; a.n = n;
        a.tw2 = static_cast<float>(total_width) / 2;
        a.centers = ctx->dMassCenters.p; a.low2 = ctx->dBandLow2.p; a.high2 = ctx->dBandHigh2.p;
        a.coeffs = ctx->dMassCoeffs.p; a.masses = nullptr;
        a.list = ctx->dMassList.p; a.listCount = ctx->dMassList.p + ctx->massHalos;
        a.bboxKeys = reinterpret_cast<unsigned *>(ctx->dMassList.p + ctx->massHalos + 1);
        a.nHalo = static_cast<int>(ctx->massHalos); a.order = ctx->massOrder; a.P2 = 2 * ctx->massOrder * ctx->massOrder;
        a.bandDelta = ctx->dBandDelta.p; a.bandInside = ctx->bandInside;
        a.bandList = ctx->dBandList.p; a.bandCount = ctx->dBandCount.p; a.bandCap = ctx->dBandList.cap;
        CU(launch_band(a, ctx->stream));
        ctx->stats.launches += 3;
        unsigned long long cnt = 0;
        CU(cudaMemcpyAsync(&cnt, ctx->dBandCount.p, sizeof cnt, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (cnt > ctx->dBandList.cap) {   // grow and run the block again
            CU(ctx->dBandList.ensure(static_cast<size_t>(cnt) * 5 / 4));
            continue;
        }
        ctx->hBand.resize(cnt);
        if (cnt) CU(cudaMemcpy(ctx->hBand.data(), ctx->dBandList.p, cnt * sizeof(long long), cudaMemcpyDeviceToHost));
        std::sort(ctx->hBand.begin(), ctx->hBand.end());   // by halo, then by particle index: the reference's order at Threads = 1
        return SFK_OK;
    }
}

int sfk_mass_take_band(sfk_ctx *ctx, int64_t *n, const int64_t **pairs) try {
    if (!ctx || !ctx->massOpen || !ctx->bandOn || !n || !pairs)
        return fail(ctx, SFK_E_INVALID, "sfk_mass_take_band: call sfk_mass_set_band and push a block first");
    static_assert(sizeof(long long) == sizeof(int64_t), "int64");
    *n = static_cast<int64_t>(ctx->hBand.size());
    *pairs = reinterpret_cast<const int64_t *>(ctx->hBand.data());
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_mass_finish(sfk_ctx *ctx, double *masses) try {
    if (!ctx || !ctx->massOpen || (ctx->massHalos > 0 && !masses))
        return fail(ctx, SFK_E_INVALID, "sfk_mass_finish: call sfk_mass_begin first");
    CU(cudaSetDevice(ctx->p.device));
    if (ctx->massHalos)
        CU(cudaMemcpyAsync(masses, ctx->dMassSums.p, ctx->massHalos * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->massOpen = false;
    return SFK_OK;
} SFK_CATCH(ctx)

// ---- `stats` mode: shell property integrals --------------------------------------------

static_assert(SS_COUNT == SFK_SHELL_SUMS, "sum layout");

// Gauss-Legendre nodes / weights of the solid-angle rule, built once per context.
static int ensure_rule(sfk_ctx *ctx) {
    if (ctx->glTheta == SFK_SHELL_RULE_THETA) return SFK_OK;
    std::vector<double> x, w;
    gauss_legendre(SFK_SHELL_RULE_THETA, x, w);
    CU(ctx->dGlX.upload(x, ctx->stream));
    CU(ctx->dGlW.upload(w, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->glTheta = SFK_SHELL_RULE_THETA;
    return SFK_OK;
}

int sfk_shell_props(sfk_ctx *ctx, int64_t n_halo, int32_t order, const double *coeffs, double *props) try {
    if (!ctx || n_halo < 0 || order <= 0 || (n_halo > 0 && (!coeffs || !props)))
        return fail(ctx, SFK_E_INVALID, "sfk_shell_props: bad arguments");
    if (order > kMaxOrder) return fail(ctx, SFK_E_UNSUPPORTED, "Order above 4 is not supported.");
    if (n_halo == 0) return SFK_OK;
    if (n_halo > 0x7fffffff) return fail(ctx, SFK_E_INVALID, "sfk_shell_props: too many halos");
    CU(cudaSetDevice(ctx->p.device));
    const int nTheta = SFK_SHELL_RULE_THETA, nPhi = SFK_SHELL_RULE_PHI;
    int rcRule = ensure_rule(ctx);
    if (rcRule) return rcRule;
    const size_t P2 = 2 * static_cast<size_t>(order) * order;
    CU(ctx->dPropCoeffs.ensure(P2 * n_halo));
    CU(ctx->dPropSums.ensure(static_cast<size_t>(SS_COUNT) * n_halo));
    CU(cudaMemcpyAsync(ctx->dPropCoeffs.p, coeffs, P2 * n_halo * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ShellPropArgs a{};
    a.coeffs = ctx->dPropCoeffs.p; a.glX = ctx->dGlX.p; a.glW = ctx->dGlW.p; a.sums = ctx->dPropSums.p;
    a.nHalo = static_cast<int>(n_halo); a.order = order; a.nTheta = nTheta; a.nPhi = nPhi;
    CU(launch_shell_props(a, ctx->stream));
    ctx->stats.launches += 1;
    std::vector<double> sums(static_cast<size_t>(SS_COUNT) * n_halo);
    CU(cudaMemcpyAsync(sums.data(), ctx->dPropSums.p, sums.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (int64_t h = 0; h < n_halo; h++)
        shell_props_finalize(sums.data() + h * SS_COUNT, props + h * SFK_SHELL_PROPS);
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_angular_fraction(sfk_ctx *ctx, int64_t n_halo, int32_t order, const double *coeffs, const double *r_min,
                         const double *r_max, int32_t bins, double *rs, double *fs) try {
    if (!ctx || n_halo < 0 || order <= 0 || bins < 0 ||
        (n_halo > 0 && bins > 0 && (!coeffs || !r_min || !r_max || !rs || !fs)))
        return fail(ctx, SFK_E_INVALID, "sfk_angular_fraction: bad arguments");
    if (order > kMaxOrder) return fail(ctx, SFK_E_UNSUPPORTED, "Order above 4 is not supported.");
    if (bins > 4096) return fail(ctx, SFK_E_UNSUPPORTED, "Bins above 4096 is not supported.");
    if (n_halo == 0 || bins == 0) return SFK_OK;
    if (n_halo > 0x7fffffff) return fail(ctx, SFK_E_INVALID, "sfk_angular_fraction: too many halos");
    CU(cudaSetDevice(ctx->p.device));
    const size_t P2 = 2 * static_cast<size_t>(order) * order;
    // one staging buffer: coefficients, rMin, rMax, then the fractions
    CU(ctx->dPropCoeffs.ensure((P2 + 2) * n_halo));
    CU(ctx->dPropSums.ensure(std::max<size_t>(static_cast<size_t>(bins), SS_COUNT) * n_halo));
    double *dMin = ctx->dPropCoeffs.p + P2 * n_halo, *dMax = dMin + n_halo;
    CU(cudaMemcpyAsync(ctx->dPropCoeffs.p, coeffs, P2 * n_halo * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dMin, r_min, n_halo * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dMax, r_max, n_halo * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    ShellFractionArgs a{};
    a.coeffs = ctx->dPropCoeffs.p; a.rMin = dMin; a.rMax = dMax;
    a.fs = ctx->dPropSums.p;
    a.nHalo = static_cast<int>(n_halo); a.order = order; a.bins = bins;
    a.nTheta = SFK_FRACTION_STRATA_THETA; a.nPhi = SFK_FRACTION_STRATA_PHI;
    CU(launch_shell_fraction(a, ctx->stream));
    ctx->stats.launches += 1;
    CU(cudaMemcpyAsync(fs, ctx->dPropSums.p, static_cast<size_t>(bins) * n_halo * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    // bin centres, shell.go:304-309
    for (int64_t h = 0; h < n_halo; h++) {
        const double lrMin = std::log(r_min[h]), lrMax = std::log(r_max[h]);
        const double dlr = (lrMax - lrMin) / static_cast<double>(bins);
        for (int i = 0; i < bins; i++) rs[h * bins + i] = std::exp(lrMin + (static_cast<double>(i) + 0.5) * dlr);
    }
    CU(cudaStreamSynchronize(ctx->stream));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_shell_sums(sfk_ctx *ctx, int64_t n_halo, double *sums) try {
    if (!ctx || n_halo < 0 || (n_halo > 0 && !sums) || static_cast<size_t>(n_halo) * SS_COUNT > ctx->dPropSums.cap)
        return fail(ctx, SFK_E_INVALID, "sfk_shell_sums: call sfk_shell_props with at least n_halo halos first");
    CU(cudaSetDevice(ctx->p.device));
    if (n_halo)
        CU(cudaMemcpyAsync(sums, ctx->dPropSums.p, static_cast<size_t>(n_halo) * SS_COUNT * sizeof(double),
                           cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_axes_from_moments(const double moments[9], double out[6]) {
    if (!moments || !out) return SFK_E_INVALID;
    axes_from_moments(moments, out);
    return SFK_OK;
}

int sfk_get_stats(const sfk_ctx *ctx, sfk_stats *out) {
    if (!ctx || !out) return SFK_E_INVALID;
    *out = ctx->stats;
    return SFK_OK;
}

void *sfk_stream(const sfk_ctx *ctx) { return ctx ? static_cast<void *>(ctx->stream) : nullptr; }

static int check_halo(sfk_ctx *ctx, int64_t halo, bool needTaps) {
    if (!ctx || !ctx->finished) return fail(ctx, SFK_E_INVALID, "taps are valid after sfk_finish_snapshot");
    if (halo < 0 || halo >= static_cast<int64_t>(ctx->halos.size())) return fail(ctx, SFK_E_INVALID, "halo index out of range");
    if (needTaps && !ctx->taps) return fail(ctx, SFK_E_INVALID, "context was created without SFK_FLAG_TAPS");
    return SFK_OK;
}

int sfk_get_rsp(sfk_ctx *ctx, int64_t halo, double *r) try {
    int rc = check_halo(ctx, halo, false);
    if (rc) return rc;
    const size_t rn = static_cast<size_t>(ctx->rings) * ctx->spokes;
    CU(cudaMemcpy(r, ctx->dRsp.p + halo * rn, rn * sizeof(double), cudaMemcpyDeviceToHost));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_get_candidate_count(sfk_ctx *ctx, int64_t halo, int64_t *n) try {
    int rc = check_halo(ctx, halo, false);
    if (rc) return rc;
    *n = ctx->candPerHalo[halo];
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_get_intersection_count(sfk_ctx *ctx, int64_t halo, int64_t *n) try {
    int rc = check_halo(ctx, halo, false);
    if (rc) return rc;
    *n = halo < static_cast<int64_t>(ctx->interPerHalo.size()) ? ctx->interPerHalo[halo] : 0;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_get_profiles(sfk_ctx *ctx, int64_t halo, double *rho) try {
    int rc = check_halo(ctx, halo, true);
    if (rc) return rc;
    const size_t rnb = static_cast<size_t>(ctx->rings) * ctx->spokes * ctx->bins;
    CU(cudaMemcpy(rho, ctx->dProfiles.p + halo * rnb, rnb * sizeof(double), cudaMemcpyDeviceToHost));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_get_counts(sfk_ctx *ctx, int64_t halo, uint32_t *n_insert, uint32_t *start_edges,
                   uint32_t *end_edges) try {
    int rc = check_halo(ctx, halo, true);
    if (rc) return rc;
    const size_t rn = static_cast<size_t>(ctx->rings) * ctx->spokes, rnb = rn * ctx->bins;
    if (n_insert) CU(cudaMemcpy(n_insert, ctx->dTapInsert.p + halo * rn, rn * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (start_edges) CU(cudaMemcpy(start_edges, ctx->dTapStart.p + halo * rnb, rnb * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    if (end_edges) CU(cudaMemcpy(end_edges, ctx->dTapEnd.p + halo * rnb, rnb * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_get_grid(sfk_ctx *ctx, int64_t halo, int64_t *grid, double *quantum) try {
    int rc = check_halo(ctx, halo, false);
    if (rc) return rc;
    if (!ctx->keepGrid) return fail(ctx, SFK_E_INVALID, "context was created without SFK_FLAG_KEEP_GRID");
    const size_t rnb = static_cast<size_t>(ctx->rings) * ctx->spokes * ctx->bins;
    if (grid) CU(cudaMemcpy(grid, ctx->dGridTap.p + halo * rnb, rnb * sizeof(int64_t), cudaMemcpyDeviceToHost));
    if (quantum) *quantum = ctx->halos[halo].quantum;
    return SFK_OK;
} SFK_CATCH(ctx)

int sfk_get_ring_tables(const sfk_ctx *ctx, float *norms, float *rot, float *irot) {
    if (!ctx) return SFK_E_INVALID;
    if (norms) std::memcpy(norms, ctx->tab.norms.data(), ctx->tab.norms.size() * sizeof(float));
    if (rot) std::memcpy(rot, ctx->tab.rots.data(), ctx->tab.rots.size() * sizeof(float));
    if (irot) std::memcpy(irot, ctx->tab.irots.data(), ctx->tab.irots.size() * sizeof(float));
    return SFK_OK;
}

}  // extern "C"
