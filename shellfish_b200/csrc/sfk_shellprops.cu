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

}  // namespace

cudaError_t launch_shell_props(const ShellPropArgs &a, cudaStream_t s) {
    if (a.nHalo <= 0) return cudaSuccess;
    shell_props_kernel<<<static_cast<unsigned>(a.nHalo), kPropThreads, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace sfk
