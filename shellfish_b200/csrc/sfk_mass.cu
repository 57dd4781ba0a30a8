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
#include <cstdint>
#include <cuda_runtime.h>

#include "sfk_kernels.cuh"

namespace sfk {

namespace {

constexpr int kMassThreads = 256;
constexpr int kMassPerThread = 4;
constexpr int kMassHaloChunk = 256;

// Go's math.Pow for small non-negative integer exponents (src/math/pow.go):
// repeated squaring on frexp mantissas.
__device__ double go_pow_small(double x, int yi) {
    if (yi == 0 || x == 1) return 1;
    if (yi == 1) return x;
    if (isnan(x)) return x;
    if (x == 0) return 0;
    if (isinf(x)) return pow(x, static_cast<double>(yi));
    double a1 = 1.0;
    int ae = 0, xe;
    double x1 = frexp(x, &xe);
    for (int i = yi; i != 0; i >>= 1) {
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    return ldexp(a1, ae);
}

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
        const double cosK = go_pow_small(cosTh, k);
        for (int j = 0; j < order; j++) {
            const double sinJ = go_pow_small(sinPhi, j);
            for (int i = 0; i < order; i++) {
                const double cosI = go_pow_small(cosPhi, i);
                const double sinIJ = go_pow_small(sinTh, i + j);
                sum += cs[idx] * sinIJ * cosK * sinJ * cosI;
                idx++;
            }
        }
    }
    return sum > r;
}

__device__ __forceinline__ float wrap_half(float x, float tw2) {   // cmd/stats.go:429-436
    if (x > tw2) return x - tw2;
    if (x < -tw2) return x + tw2;
    return x;
}

__global__ void __launch_bounds__(kMassThreads) mass_kernel(MassArgs a) {
    __shared__ float4 sC[kMassHaloChunk];    // centre (float32) and low2
    __shared__ float sHigh2[kMassHaloChunk];
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
    for (int h0 = 0; h0 < a.nHalo; h0 += kMassHaloChunk) {
        __syncthreads();
        for (int k = threadIdx.x; k < kMassHaloChunk && h0 + k < a.nHalo; k += kMassThreads) {
            sC[k] = make_float4(a.centers[3 * (h0 + k)], a.centers[3 * (h0 + k) + 1], a.centers[3 * (h0 + k) + 2],
                                a.low2[h0 + k]);
            sHigh2[k] = a.high2[h0 + k];
        }
        __syncthreads();
        const int cnt = min(kMassHaloChunk, a.nHalo - h0);
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
                    in = shell_contains(a.coeffs + static_cast<size_t>(h0 + k) * a.P2, a.order, x, y, z);
                if (in && pm[q] >= 0.f) { sum += static_cast<double>(pm[q]); any = true; }
            }
            if (__any_sync(0xffffffffu, any)) {
                for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
                if (lane == 0) atomicAdd(&a.masses[h0 + k], sum);
            }
        }
    }
}

}  // namespace

cudaError_t launch_mass(const MassArgs &a, cudaStream_t s) {
    if (a.n <= 0 || a.nHalo <= 0) return cudaSuccess;
    const int64_t perCta = static_cast<int64_t>(kMassThreads) * kMassPerThread;
    mass_kernel<<<static_cast<unsigned>((a.n + perCta - 1) / perCta), kMassThreads, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sfk
