// sfk_mass.cu -- the M_sp pass of `shellfish stats`: massContained /
// massContainedChan (cmd/stats.go:451-543) and Shell.Contains
// (los/analyze/shell.go:349-354) for every halo over one particle block.
//
// The reference tests every particle of a block against every halo of the
// snapshot (cmd/stats.go:322-327), and so does this kernel: a thread owns a few
// particles and walks the halo table staged in shared memory.  The distance test
// is the reference's float32 chain including its periodic `wrap`, which moves a
// coordinate by HALF the box (cmd/stats.go:429-436) -- particles half a box away
// along an axis are therefore counted too, and must be to match M_sp.  Only the
// few particles between rLow and rHigh evaluate the Penna-Dines surface.
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

#include "sfk_device.cuh"
#include "sfk_kernels.cuh"

namespace sfk {

namespace {

constexpr int kMassThreads = 256;
constexpr int kMassPerThread = 4;
constexpr int kMassHaloChunk = 256;

// PennaFunc(cs, order, order, 2)(phi, theta) > r, penna.go:62-82 + shell.go:349-354
__device__ bool shell_contains(const double *cs, int order, double x, double y, double z) {
    const double r = sqrt(x * x + y * y + z * z);
    const double phi = atan2(y, x);
    const double theta = acos(z / r);
    double sinPhi, cosPhi, sinTh, cosTh;
    sincos(phi, &sinPhi, &cosPhi);
    sincos(theta, &sinTh, &cosTh);
    int idx = 0;
    double sum = 0.0;
    for (int k = 0; k < 2; k++) {
        const double cosK = go_pow_int(cosTh, k);
        for (int j = 0; j < order; j++) {
            const double sinJ = go_pow_int(sinPhi, j);
            for (int i = 0; i < order; i++) {
                const double cosI = go_pow_int(cosPhi, i);
                const double sinIJ = go_pow_int(sinTh, i + j);
                sum += cs[idx] * sinIJ * cosK * sinJ * cosI;
                idx++;
            }
        }
    }
    return sum > r;
}

// appendShellParticlesChan, cmd/stats.go:545-597: is the particle within `delta` of the
// surface (delta >= 0), or anywhere inside it (inside != 0, ShellWidth < 0)?
__device__ bool shell_band(const double *cs, int order, float xf, float yf, float zf, float r2, double delta, int inside) {
    const double x = xf, y = yf, z = zf;
    const double r = sqrt(static_cast<double>(r2));
    const double phi = atan2(y, x);
    const double theta = acos(z / r);
    double sinPhi, cosPhi, sinTh, cosTh;
    sincos(phi, &sinPhi, &cosPhi);
    sincos(theta, &sinTh, &cosTh);
    int idx = 0;
    double rs = 0.0;
    for (int k = 0; k < 2; k++) {
        const double cosK = go_pow_int(cosTh, k);
        for (int j = 0; j < order; j++) {
            const double sinJ = go_pow_int(sinPhi, j);
            for (int i = 0; i < order; i++) {
                const double cosI = go_pow_int(cosPhi, i);
                const double sinIJ = go_pow_int(sinTh, i + j);
                rs += cs[idx] * sinIJ * cosK * sinJ * cosI;
                idx++;
            }
        }
    }
    if (inside) return rs > r;
    return rs + delta > r && rs - delta < r;
}

__device__ __forceinline__ float wrap_half(float x, float tw2) {   // cmd/stats.go:429-436
    if (x > tw2) return x - tw2;
    if (x < -tw2) return x + tw2;
    return x;
}

// ---- per-block halo list -------------------------------------------------------------
// A block only has to be tested against the halos its particles can reach.  After
// `wrap_half` a coordinate difference d = x - c lands within [-r, r] iff d lies in
// [-r, r], (tw2, tw2 + r] or [-tw2 - r, -tw2): the halo itself and the two "ghost"
// slabs half a box away that the reference's wrap folds onto it.  bbox_kernel finds
// the block's coordinate range, list_kernel keeps (in halo order) the halos for which
// all three axes can hit, with a margin far above float32 rounding.  Skipped pairs
// fail r2 < high2 in the reference, so the sums are unchanged.

__device__ __forceinline__ unsigned f2key(float f) {   // order-preserving float -> uint
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

__global__ void __launch_bounds__(256) bbox_kernel(const float *xyz, int64_t n, unsigned *keys /*[6]: min xyz, max xyz*/) {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const float v = xyz[3 * i + d];
            lo[d] = fminf(lo[d], v); hi[d] = fmaxf(hi[d], v);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; d++) {
        for (int o = 16; o > 0; o >>= 1) {
            lo[d] = fminf(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
            hi[d] = fmaxf(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&keys[d], f2key(lo[d]));
            atomicMax(&keys[3 + d], f2key(hi[d]));
        }
    }
}

__device__ __forceinline__ bool axis_can_hit(double bmin, double bmax, double c, double r, double tw2, double m) {
    const double lo = bmin - c, hi = bmax - c;   // range of x - c over the block
    return (lo <= r + m && hi >= -r - m) || (lo <= tw2 + r + m && hi >= tw2 - m) ||
           (lo <= -tw2 + m && hi >= -tw2 - r - m);
}

__global__ void __launch_bounds__(1024) list_kernel(MassArgs a) {
    __shared__ int warpCnt[32];
    __shared__ int base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double bmin[3], bmax[3];
    for (int d = 0; d < 3; d++) { bmin[d] = key2f(a.bboxKeys[d]); bmax[d] = key2f(a.bboxKeys[3 + d]); }
    const double tw2 = a.tw2, m = 1e-4 * (2.0 * tw2) + 1e-6;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (int h0 = 0; h0 < a.nHalo; h0 += 1024) {
        const int h = h0 + threadIdx.x;
        bool keep = false;
        if (h < a.nHalo) {
            const double r = sqrt(static_cast<double>(a.high2[h])) * (1.0 + 1e-6);
            keep = true;
            for (int d = 0; d < 3; d++)
                keep = keep && axis_can_hit(bmin[d], bmax[d], a.centers[3 * h + d], r, tw2, m);
        }
        const unsigned bal = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warpCnt[warp] = __popc(bal);
        __syncthreads();
        int before = base;
        for (int w = 0; w < warp; w++) before += warpCnt[w];
        if (keep) a.list[before + __popc(bal & ((1u << lane) - 1u))] = h;
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int w = 0; w < 32; w++) t += warpCnt[w]; base += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *a.listCount = base;
}

__global__ void __launch_bounds__(kMassThreads) mass_kernel(MassArgs a) {
    __shared__ float4 sC[kMassHaloChunk];    // centre (float32) and low2
    __shared__ float sHigh2[kMassHaloChunk];
    __shared__ int sIdx[kMassHaloChunk];
    const int lane = threadIdx.x & 31;
    const int64_t base = (static_cast<int64_t>(blockIdx.x) * kMassThreads + threadIdx.x) * kMassPerThread;
    float px[kMassPerThread], py[kMassPerThread], pz[kMassPerThread], pm[kMassPerThread];
#pragma unroll
    for (int q = 0; q < kMassPerThread; q++) {
        const int64_t i = base + q;
        const bool ok = i < a.n;
        px[q] = ok ? a.xyz[3 * i] : 0.f;
        py[q] = ok ? a.xyz[3 * i + 1] : 0.f;
        pz[q] = ok ? a.xyz[3 * i + 2] : 0.f;
        pm[q] = ok ? a.mass[i] : -1.f;   // negative: no particle
    }
    const float tw2 = a.tw2;
    const int nList = *a.listCount;
    for (int h0 = 0; h0 < nList; h0 += kMassHaloChunk) {
        __syncthreads();
        for (int k = threadIdx.x; k < kMassHaloChunk && h0 + k < nList; k += kMassThreads) {
            const int h = a.list[h0 + k];
            sIdx[k] = h;
            sC[k] = make_float4(a.centers[3 * h], a.centers[3 * h + 1], a.centers[3 * h + 2], a.low2[h]);
            sHigh2[k] = a.high2[h];
        }
        __syncthreads();
        const int cnt = min(kMassHaloChunk, nList - h0);
        for (int k = 0; k < cnt; k++) {
            const float4 c = sC[k];
            const float high2 = sHigh2[k];
            double sum = 0.0;
            bool any = false;
#pragma unroll
            for (int q = 0; q < kMassPerThread; q++) {
                const float x = wrap_half(px[q] - c.x, tw2);
                const float y = wrap_half(py[q] - c.y, tw2);
                const float z = wrap_half(pz[q] - c.z, tw2);
                const float r2 = x * x + y * y + z * z;
                bool in = r2 < c.w;
                if (!in && r2 < high2 && pm[q] >= 0.f)
                    in = shell_contains(a.coeffs + static_cast<size_t>(sIdx[k]) * a.P2, a.order, x, y, z);
                if (in && pm[q] >= 0.f) { sum += static_cast<double>(pm[q]); any = true; }
            }
            if (__any_sync(0xffffffffu, any)) {
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) atomicAdd(&a.masses[sIdx[k]], sum);
            }
        }
    }
}

// The ShellParticleFile pass: the (halo, particle) pairs of one block that appendShellParticlesChan
// keeps, appended as halo << 32 | particle index.  a.low2 / a.high2 hold the widened band here.
__global__ void __launch_bounds__(kMassThreads) band_kernel(MassArgs a) {
    __shared__ float4 sC[kMassHaloChunk];
    __shared__ float sHigh2[kMassHaloChunk];
    __shared__ int sIdx[kMassHaloChunk];
    const int64_t i = static_cast<int64_t>(blockIdx.x) * kMassThreads + threadIdx.x;
    const bool live = i < a.n;
    const float px = live ? a.xyz[3 * i] : 0.f, py = live ? a.xyz[3 * i + 1] : 0.f, pz = live ? a.xyz[3 * i + 2] : 0.f;
    const float tw2 = a.tw2;
    const int nList = *a.listCount;
    for (int h0 = 0; h0 < nList; h0 += kMassHaloChunk) {
        __syncthreads();
        for (int k = threadIdx.x; k < kMassHaloChunk && h0 + k < nList; k += kMassThreads) {
            const int h = a.list[h0 + k];
            sIdx[k] = h;
            sC[k] = make_float4(a.centers[3 * h], a.centers[3 * h + 1], a.centers[3 * h + 2], a.low2[h]);
            sHigh2[k] = a.high2[h];
        }
        __syncthreads();
        const int cnt = min(kMassHaloChunk, nList - h0);
        for (int k = 0; k < cnt; k++) {
            const float4 c = sC[k];
            const float x = wrap_half(px - c.x, tw2), y = wrap_half(py - c.y, tw2), z = wrap_half(pz - c.z, tw2);
            const float r2 = x * x + y * y + z * z;
            if (!live || !(r2 < sHigh2[k]) || (!a.bandInside && !(r2 > c.w))) continue;
            const int h = sIdx[k];
            if (!shell_band(a.coeffs + static_cast<size_t>(h) * a.P2, a.order, x, y, z, r2, a.bandDelta[h], a.bandInside)) continue;
            const unsigned long long at = atomicAdd(a.bandCount, 1ull);
            if (at < a.bandCap) a.bandList[at] = static_cast<long long>(h) << 32 | static_cast<long long>(i);
        }
    }
}

}  // namespace

cudaError_t launch_band(const MassArgs &a, cudaStream_t s) {
    if (a.n <= 0 || a.nHalo <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.bboxKeys, 0xff, 3 * sizeof(unsigned), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(a.bboxKeys + 3, 0, 3 * sizeof(unsigned), s);
    if (e != cudaSuccess) return e;
    bbox_kernel<<<static_cast<unsigned>(std::min<int64_t>(148 * 8, (a.n + 255) / 256)), 256, 0, s>>>(a.xyz, a.n, a.bboxKeys);
    list_kernel<<<1, 1024, 0, s>>>(a);
    band_kernel<<<static_cast<unsigned>((a.n + kMassThreads - 1) / kMassThreads), kMassThreads, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mass(const MassArgs &a, cudaStream_t s) {
    if (a.n <= 0 || a.nHalo <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(a.bboxKeys, 0xff, 3 * sizeof(unsigned), s);       // min keys
    if (e == cudaSuccess) e = cudaMemsetAsync(a.bboxKeys + 3, 0, 3 * sizeof(unsigned), s);   // max keys
    if (e != cudaSuccess) return e;
    bbox_kernel<<<static_cast<unsigned>(std::min<int64_t>(148 * 8, (a.n + 255) / 256)), 256, 0, s>>>(a.xyz, a.n, a.bboxKeys);
    list_kernel<<<1, 1024, 0, s>>>(a);
    const int64_t perCta = static_cast<int64_t>(kMassThreads) * kMassPerThread;
    mass_kernel<<<static_cast<unsigned>((a.n + perCta - 1) / perCta), kMassThreads, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sfk
