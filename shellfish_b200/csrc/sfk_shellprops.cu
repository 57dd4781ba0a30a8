// sfk_shellprops.cu -- shell property integrals of `shellfish stats`:
// Shell.Volume, SurfaceArea, Axes and RadialRange (los/analyze/shell.go:58-283)
// for every halo of the catalog (cmd/stats.go:274-289).
//
// The reference draws MonteCarloSamples = 50 000 solid angles per quantity from
// Go's process-global math/rand (four independent passes per halo, ~10^6 surface
// evaluations).  Nothing pins that stream (PARITY UNPINNED), so the same
// integrals are taken here over one deterministic product rule -- Gauss-Legendre
// in cos(theta) x equally spaced phi, the measure randomAngle samples -- whose
// error is far below the Monte-Carlo noise; results are compared with the
// literal Monte-Carlo restatement kept with the tests, statistically.
//
// One CTA per halo; a thread walks nodes t, t + 256, ... with the per-node
// arithmetic of sfk_shellmath.cuh, the 13 partial sums are combined by a fixed
// shuffle / shared-memory tree (deterministic).  Compute bound (FP64), bytes are
// negligible: 2*order^2 coefficients in, 13 doubles out per halo.
#include <cuda_runtime.h>

#include "sfk_kernels.cuh"
#include "sfk_shellmath.cuh"

namespace sfk {

namespace {

constexpr int kPropThreads = 256;

__global__ void __launch_bounds__(kPropThreads) shell_props_kernel(ShellPropArgs a) {
    __shared__ double sCs[2 * kShellMaxOrder * kShellMaxOrder];
    __shared__ double sPart[kPropThreads / 32][SS_COUNT];
    const int h = blockIdx.x;
    const int P2 = 2 * a.order * a.order;
    for (int i = threadIdx.x; i < P2; i += kPropThreads) sCs[i] = a.coeffs[static_cast<size_t>(h) * P2 + i];
    __syncthreads();

    double acc[SS_COUNT];
#pragma unroll
    for (int i = 0; i < SS_COUNT; i++) acc[i] = 0.0;
    acc[SS_RMIN] = INFINITY; acc[SS_RMAX] = -INFINITY;
    const int nNodes = a.nTheta * a.nPhi;
    for (int q = threadIdx.x; q < nNodes; q += kPropThreads) {
        double phi, theta, w;
        shell_rule_node(q, a.nPhi, a.glX, a.glW, phi, theta, w);
        shell_node(sCs, a.order, phi, theta, w, acc);
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < SS_COUNT; i++) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) {
            const double u = __shfl_xor_sync(0xffffffffu, v, o);
            v = i == SS_RMIN ? fmin(v, u) : i == SS_RMAX ? fmax(v, u) : v + u;
        }
        if (lane == 0) sPart[warp][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < SS_COUNT) {
        const int i = threadIdx.x;
        double v = sPart[0][i];
        for (int w = 1; w < kPropThreads / 32; w++) {
            const double u = sPart[w][i];
            v = i == SS_RMIN ? fmin(v, u) : i == SS_RMAX ? fmax(v, u) : v + u;
        }
        a.sums[static_cast<size_t>(h) * SS_COUNT + i] = v;
    }
}

// Shell.AngularFractionProfile, shell.go:295-335, for `prof`'s angular-fraction type
// (cmd/prof.go:628-661): histogram of ln R(phi, theta) over the rule's nodes in `bins`
// log-radial bins of [rMin, rMax], reverse cumulative sum, normalised by its first entry.
// The directions are nTheta x nPhi equal-area strata (shell_stratum), so the histogram
// holds integer counts: independent of the atomic order.
__global__ void __launch_bounds__(kPropThreads) shell_fraction_kernel(ShellFractionArgs a) {
    extern __shared__ unsigned long long sHist[];
    __shared__ double sCs[2 * kShellMaxOrder * kShellMaxOrder];
    const int h = blockIdx.x;
    const int P2 = 2 * a.order * a.order;
    for (int i = threadIdx.x; i < P2; i += kPropThreads) sCs[i] = a.coeffs[static_cast<size_t>(h) * P2 + i];
    for (int i = threadIdx.x; i < a.bins; i += kPropThreads) sHist[i] = 0ull;
    __syncthreads();
    const double lrMin = log(a.rMin[h]), lrMax = log(a.rMax[h]);
    const double dlr = (lrMax - lrMin) / static_cast<double>(a.bins);
    const int nNodes = a.nTheta * a.nPhi;
    for (int q = threadIdx.x; q < nNodes; q += kPropThreads) {
        double phi, theta;
        shell_stratum(q, a.nTheta, a.nPhi, phi, theta);
        const int b = shell_fraction_bin(penna_radius(sCs, a.order, phi, theta), lrMin, dlr, a.bins);
        if (b >= 0) atomicAdd(&sHist[b], 1ull);
    }
    __syncthreads();
    if (threadIdx.x == 0)   // reverse cumulative sum, shell.go:325-328
        for (int i = a.bins - 2; i >= 0; i--) sHist[i] += sHist[i + 1];
    __syncthreads();
    const double n0 = static_cast<double>(sHist[0]);
    for (int i = threadIdx.x; i < a.bins; i += kPropThreads)
        a.fs[static_cast<size_t>(h) * a.bins + i] = static_cast<double>(sHist[i]) / n0;   // :330-332
}

}  // namespace

cudaError_t launch_shell_fraction(const ShellFractionArgs &a, cudaStream_t s) {
    if (a.nHalo <= 0 || a.bins <= 0) return cudaSuccess;
    shell_fraction_kernel<<<static_cast<unsigned>(a.nHalo), kPropThreads, a.bins * sizeof(unsigned long long), s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_shell_props(const ShellPropArgs &a, cudaStream_t s) {
    if (a.nHalo <= 0) return cudaSuccess;
    shell_props_kernel<<<static_cast<unsigned>(a.nHalo), kPropThreads, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sfk
