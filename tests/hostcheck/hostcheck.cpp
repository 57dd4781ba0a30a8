// hostcheck.cpp -- TEST INFRASTRUCTURE.  Compiles the host/device-shared
// arithmetic of shellfish_b200/csrc/sfk_shellmath.cuh (the per-node code of
// shell_props_kernel) and the closing arithmetic of sfk_axes.cpp with g++, so the
// CPU test suite can check them against the oracle without a GPU.  Never loaded by
// the product path: libsfk.so runs this arithmetic only inside its CUDA kernel.
#include <cmath>
#include <vector>

#include "../../shellfish_b200/csrc/sfk_internal.h"
#include "../../shellfish_b200/csrc/sfk_sgmath.cuh"
#include "../../shellfish_b200/csrc/sfk_shellmath.cuh"

extern "C" {

double hc_penna_radius(const double *cs, int order, double phi, double theta) {
    return sfk::penna_radius(cs, order, phi, theta);
}

double hc_cos_norm(const double *cs, int order, double phi, double theta) {
    return sfk::shell_cos_norm(cs, order, phi, theta);
}

// The sums shell_props_kernel produces for one halo, nodes visited in index order.
void hc_shell_sums(const double *cs, int order, int nTheta, int nPhi, double *sums) {
    std::vector<double> x, w;
    sfk::gauss_legendre(nTheta, x, w);
    for (int i = 0; i < sfk::SS_COUNT; i++) sums[i] = 0.0;
    sums[sfk::SS_RMIN] = INFINITY; sums[sfk::SS_RMAX] = -INFINITY;
    for (int q = 0; q < nTheta * nPhi; q++) {
        double phi, theta, wt;
        sfk::shell_rule_node(q, nPhi, x.data(), w.data(), phi, theta, wt);
        sfk::shell_node(cs, order, phi, theta, wt, sums);
    }
}

void hc_shell_props(const double *cs, int order, int nTheta, int nPhi, double *props) {
    double sums[sfk::SS_COUNT];
    hc_shell_sums(cs, order, nTheta, nPhi, sums);
    sfk::shell_props_finalize(sums, props);
}

// shell_fraction_kernel's histogram for one halo.
void hc_angular_fraction(const double *cs, int order, int nTheta, int nPhi, double rMin, double rMax, int bins,
                         double *fs) {
    std::vector<unsigned long long> hist(bins, 0ull);
    const double lrMin = std::log(rMin), lrMax = std::log(rMax);
    const double dlr = (lrMax - lrMin) / static_cast<double>(bins);
    for (int q = 0; q < nTheta * nPhi; q++) {
        double phi, theta;
        sfk::shell_stratum(q, nTheta, nPhi, phi, theta);
        const int b = sfk::shell_fraction_bin(sfk::penna_radius(cs, order, phi, theta), lrMin, dlr, bins);
        if (b >= 0) hist[b] += 1;
    }
    for (int i = bins - 2; i >= 0; i--) hist[i] += hist[i + 1];
    for (int i = 0; i < bins; i++) fs[i] = static_cast<double>(hist[i]) / static_cast<double>(hist[0]);
}

// The sliding-moment Savitzky-Golay path of los_kernel_fast<PER> for one line of sight:
// ys[bins] = ln(rho); lane L evaluates outputs L*per .. L*per + per - 1 (one direct anchor,
// per - 1 shifts) on the Extension-padded row, exactly as the kernel's lanes do.
// per < 0: |per| outputs per lane with the paired anchor (sg_accumulate_pair).
int hc_sg_moments(const double *ys, int bins, const double *taps0, const double *taps1, int window, int perArg,
                  double *s0, double *s1) {
    const bool paired = perArg < 0;
    const int per = paired ? -perArg : perArg;
    std::vector<double> t0(taps0, taps0 + window), t1(taps1, taps1 + window);
    sfk::SgPoly poly;
    if (!sfk::savgol_poly(t0, t1, poly)) return -1;
    const int center = window / 2, padL = center + 1;
    std::vector<double> pos(bins + 2 * window + 2 * per + 2);
    for (size_t p = 0; p < pos.size(); p++) {
        long i = static_cast<long>(p) - padL;
        pos[p] = ys[i < 0 ? 0 : (i >= bins ? bins - 1 : i)];
    }
    std::vector<double> upow(4 * static_cast<size_t>(window));
    for (int j = 0; j < window; j++) {
        const double u = static_cast<double>(j - center) / static_cast<double>(center);
        upow[4 * j] = u; upow[4 * j + 1] = u * u; upow[4 * j + 2] = u * u * u; upow[4 * j + 3] = (u * u) * (u * u);
    }
    for (int i0 = 0; i0 < bins; i0 += per) {
        double M[5] = {0, 0, 0, 0, 0};
        if (paired) {   // paired anchor: taps center +- m share one update
            M[0] = pos[i0 + 1 + center];
            for (int m = 1; m <= center; m++)
                sfk::sg_accumulate_pair(M, pos[i0 + 1 + center + m], pos[i0 + 1 + center - m], &upow[4 * (center + m)]);
        } else {
            for (int j = 0; j < window; j++) sfk::sg_accumulate(M, pos[i0 + 1 + j], &upow[4 * j]);
        }
        for (int c = 0; c < per && i0 + c < bins; c++) {
            sfk::sg_outputs(M, poly, s0[i0 + c], s1[i0 + c]);
            sfk::sg_advance(M, poly, pos[i0 + c + 1], pos[i0 + c + 1 + window]);
        }
    }
    return 0;
}

void hc_gauss_legendre(int n, double *x, double *w) {
    std::vector<double> xv, wv;
    sfk::gauss_legendre(n, xv, wv);
    for (int i = 0; i < n; i++) { x[i] = xv[i]; w[i] = wv[i]; }
}

}
