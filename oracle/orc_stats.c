/* orc_stats.c -- oracle (TEST INFRASTRUCTURE): the M_sp pass of `shellfish
 * stats`, cmd/stats.go:429-443 (wrap), :451-478 (massContained), :513-543
 * (massContainedChan), :685-692 (sphereSheetIntersect) and Shell.Contains,
 * los/analyze/shell.go:349-354.
 * PARITY UNPINNED: the reference holds no known answer for this pass, and the
 * rLow / rHigh it feeds in come from Monte-Carlo sampling on Go's global RNG
 * (shell.go:21-24, RadialRange); here they are inputs. */
#include "shellfish_oracle.h"

#include <math.h>
#include <stdlib.h>

/* shell.go:349-354 */
int orc_shell_contains(const double *cs, int order, double x, double y, double z) {
    double r = sqrt(x * x + y * y + z * z);
    double phi = atan2(y, x);
    double theta = acos(z / r);
    return orc_penna_func(cs, order, order, 2, phi, theta) > r;
}

/* cmd/stats.go:429-436: subtracts HALF the box width (reference quirk, SURVEY hazard 15) */
static float wrap_half(float x, float tw2) {
    if (x > tw2) return x - tw2;
    if (x < -tw2) return x + tw2;
    return x;
}

/* massContained + massContainedChan: worker k sums particles k, k+W, ... in
 * float64; the partial sums are added in worker order (the reference adds them
 * in channel-arrival order, i.e. nondeterministically). */
double orc_mass_contained(double totalWidth, const float *xs, const float *ms, int64_t n,
                          const double *cs, int order, const float c[3],
                          double rLow, double rHigh, int workers) {
    float tw2 = (float)totalWidth / 2;
    float low2 = (float)(rLow * rLow), high2 = (float)(rHigh * rHigh);
    if (workers < 1) workers = 1;
    double total = 0.0;
    for (int w = 0; w < workers; w++) {
        double sum = 0.0;
        for (int64_t i = w; i < n; i += workers) {
            float x = xs[3 * i], y = xs[3 * i + 1], z = xs[3 * i + 2];
            x = x - c[0]; y = y - c[1]; z = z - c[2];
            x = wrap_half(x, tw2); y = wrap_half(y, tw2); z = wrap_half(z, tw2);
            float r2 = x * x + y * y + z * z;
            if (r2 < low2 || (r2 < high2 && orc_shell_contains(cs, order, (double)x, (double)y, (double)z)))
                sum += (double)ms[i];
        }
        total += sum;
    }
    return total;
}

/* appendShellParticles + appendShellParticlesChan, cmd/stats.go:482-520, 545-597, with one
 * worker: indices (into the block) of the particles within shellWidth * sphereR of the surface,
 * or inside it when shellWidth < 0.  out has room for n entries; returns the count. */
int64_t orc_shell_particles(double totalWidth, const float *xs, int64_t n, const double *cs, int order,
                            const float c[3], float sphereR, double rLow, double rHigh, double shellWidth,
                            int64_t *out) {
    float tw2 = (float)totalWidth / 2;
    double delta = (double)sphereR * shellWidth;
    if (shellWidth < 0) delta = 0;
    rLow -= delta;
    rHigh += delta;
    float low2 = (float)(rLow * rLow), high2 = (float)(rHigh * rHigh);
    int64_t m = 0;
    for (int64_t i = 0; i < n; i++) {
        float x = xs[3 * i], y = xs[3 * i + 1], z = xs[3 * i + 2];
        x = x - c[0]; y = y - c[1]; z = z - c[2];
        x = wrap_half(x, tw2); y = wrap_half(y, tw2); z = wrap_half(z, tw2);
        float r2 = x * x + y * y + z * z;
        if (shellWidth < 0) {
            if (r2 < high2) {
                double r = sqrt((double)r2);
                double phi = atan2((double)y, (double)x);
                double theta = acos((double)z / r);
                double rs = orc_penna_func(cs, order, order, 2, phi, theta);
                if (rs > r) out[m++] = i;
            }
        } else if (r2 > low2 && r2 < high2) {
            double r = sqrt((double)r2);
            double phi = atan2((double)y, (double)x);
            double theta = acos((double)z / r);
            double rs = orc_penna_func(cs, order, order, 2, phi, theta);
            if (rs + delta > r && rs - delta < r) out[m++] = i;
        }
    }
    return m;
}

static double wrap_dist(double x1, double x2, double width) {
    double dist = x1 - x2;
    if (dist > width / 2) return dist - width;
    if (dist < width / -2) return dist + width;
    return dist;
}

static int in_range(double x, double r, double low, double width, double tw) {
    return (x + r > low && x - r < low + width) ||
           (wrap_dist(x, low, tw) > -r && wrap_dist(x, low + width, tw) < r);
}

/* cmd/stats.go:685-692 */
int orc_sphere_sheet_intersect(const float c[3], float r, const float origin[3],
                               const float width[3], double tw) {
    return in_range((double)c[0], (double)r, (double)origin[0], (double)width[0], tw) &&
           in_range((double)c[1], (double)r, (double)origin[1], (double)width[1], tw) &&
           in_range((double)c[2], (double)r, (double)origin[2], (double)width[2], tw);
}
