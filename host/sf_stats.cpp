// sf_stats.cpp -- `shellfish stats` (cmd/stats.go): host orchestration of the two
// device passes of the mode, sfk_shell_props (Shell.Volume / SurfaceArea / Axes /
// RadialRange, los/analyze/shell.go:58-283) and sfk_mass_* (massContained,
// cmd/stats.go:451-543), and the catalog text around them.
#include "sf_stats.hpp"

#include <algorithm>
#include <cstdio>
#include <map>
#include <memory>

#include "../include/shellfish_b200.h"
#include "sf_catalog.hpp"
#include "sf_config.hpp"

namespace sfhost {

std::string StatsConfig::ReadConfig(const std::string &fname, const std::vector<std::string> &flags) {
    ConfigVars vars("stats.config");   // cmd/stats.go:109-146
    vars.Strings(&values, "Values", {});
    vars.Int(&monteCarloSamples, "MonteCarloSamples", 50 * 1000);
    vars.String(&exclusionStrategy, "ExclusionStrategy", "none");
    vars.Int(&order, "Order", 3);
    vars.String(&shellParticleFile, "ShellParticleFile", "");
    vars.Float(&shellWidth, "ShellWidth", 0);
    vars.Bool(&skipMass, "SkipMass", false);
    shellFilter = false;
    std::string err;
    if (fname.empty()) {
        if (flags.empty()) return "";
        if (!(err = vars.ReadFlags(flags)).empty()) return err;
        shellFilter = shellWidth != 0 && !shellParticleFile.empty();
        return validate();
    }
    if (!(err = vars.ReadConfig(fname)).empty()) return err;
    if (!(err = vars.ReadFlags(flags)).empty()) return err;
    shellFilter = shellWidth != 0 && !shellParticleFile.empty();
    return validate();
}

std::string StatsConfig::validate() const {   // cmd/stats.go:148-172
    for (size_t i = 0; i < values.size(); i++) {
        const std::string &v = values[i];
        if (v != "snap" && v != "id" && v != "m_sp" && v != "r_sp" && v != "a_sp" && v != "b_sp" && v != "c_sp" &&
            v != "SA_sp/V_sp")
            return "Item " + std::to_string(i) + " of variable 'Values' is set to '" + v + "', which I don't recognize.";
    }
    if (exclusionStrategy != "none" && exclusionStrategy != "contain" && exclusionStrategy != "overlap")
        return "variable 'ExclusionStrategy' set to '" + exclusionStrategy + "', which I don't recognize.";
    if (monteCarloSamples <= 0) {
        // fmt.Errorf("The variable '%s' was set to %g", name, int64): Go prints the verb mismatch
        return "The variable 'MonteCarloSamples' was set to %!g(int64=" + std::to_string(monteCarloSamples) + ")";
    }
    return "";
}

const char *StatsConfig::ExampleConfig() {
    return "[stats.config]\n\n"
           "# All fields are optional; the values shown are the defaults (cmd/stats.go:112-118).\n"
           "# MonteCarloSamples is accepted for compatibility: the B200 build integrates the shell\n"
           "# properties over a fixed 128 x 256 quadrature rule instead of random samples.\n"
           "MonteCarloSamples = 50000\nExclusionStrategy = none\nOrder = 3\n# SkipMass = false";
}

namespace {

int find_order(size_t len) {   // findOrder, cmd/stats.go:423-434 (panics on "Impossible")
    for (int i = 1; 2 * static_cast<size_t>(i) * i <= len; i++)
        if (2 * static_cast<size_t>(i) * i == len) return i;
    return -1;
}

}  // namespace

std::string RunStats(const GlobalConfig &g, const StatsConfig &c, const std::string &stdinData, int device,
                     std::vector<std::string> &outLines) {
    // cmd/stats.go:191-207
    const size_t P2 = static_cast<size_t>(2 * c.order * c.order);
    std::vector<int> floatIdx(4 + P2);
    for (size_t i = 0; i < floatIdx.size(); i++) floatIdx[i] = static_cast<int>(i) + 2;
    std::vector<std::vector<int64_t>> intCols;
    std::vector<std::vector<double>> floatCols;
    std::string err = Parse(stdinData, {0, 1}, floatIdx, intCols, floatCols);
    if (!err.empty()) return err;
    if (intCols.empty() || intCols[0].empty()) return "No input IDs.";
    const std::vector<int64_t> &ids = intCols[0], &snaps = intCols[1];
    const size_t n = ids.size();
    const int order = find_order(P2);
    if (order < 0 || order > 4) return "Order above 4 is not supported.";

    // output columns, cmd/stats.go:209-219
    std::vector<std::vector<double>> cols(12, std::vector<double>(n, 0.0));   // masses rads vols sas as bs cs ax ay az rmins rmaxes

    std::vector<std::vector<int64_t>> shellParticles(n);   // per input row, :221
    std::map<int64_t, std::vector<size_t>> idxBins;   // binCoeffsBySnap + sort.Ints
    for (size_t i = 0; i < n; i++) idxBins[snaps[i]].push_back(i);
    int64_t firstSnap = -1;   // the reference opens snaps[0] (:232); a -1 there indexes out of range
    for (size_t i = 0; i < n && firstSnap == -1; i++) firstSnap = snaps[i];
    if (firstSnap == -1) return "Every row of the catalog has Snapshot = -1: nothing to do.";

    Catalogs cat;
    if (!(err = cat.Init(g.particle, g.SnapshotType == "gotetra")).empty()) return err;
    Context ioCtx;
    ioCtx.LGadgetNPartNum = g.LGadgetNpartNum;
    ioCtx.GadgetDMTypeIndices = g.GadgetDMTypeIndices;
    ioCtx.GadgetDMSingleMassIndices = g.GadgetSingleMassIndices;
    ioCtx.GadgetMassUnits = g.GadgetMassUnits;
    ioCtx.GadgetPositionUnits = g.GadgetPositionUnits;
    auto snap_ok = [&](int64_t s) { return s >= g.particle.SnapMin && s <= g.particle.SnapMax; };
    if (!snap_ok(firstSnap))
        return "Snapshot " + std::to_string(firstSnap) + " is outside [SnapMin, SnapMax] = [" +
               std::to_string(g.particle.SnapMin) + ", " + std::to_string(g.particle.SnapMax) + "].";
    std::unique_ptr<VectorBuffer> buf;
    if (!(err = NewVectorBuffer(g.SnapshotType, cat.ParticleCatalog(static_cast<int>(firstSnap), 0), g.Endianness, ioCtx, buf)).empty())
        return err;

    sfk_params p;
    sfk_default_params(&p);
    p.rings = 3; p.device = device;   // only the stats entry points of the context are used
    sfk_ctx *sfk = nullptr;
    struct Guard { sfk_ctx *&c; ~Guard() { if (c) sfk_destroy(c); } } guard{sfk};
    if (sfk_create(&sfk, &p) != 0) return sfk_last_error(sfk);

    for (const auto &kv : idxBins) {
        const int64_t snap = kv.first;
        if (snap == -1) continue;   // cmd/stats.go:250-252
        if (!snap_ok(snap))
            return "Snapshot " + std::to_string(snap) + " is outside [SnapMin, SnapMax] = [" +
                   std::to_string(g.particle.SnapMin) + ", " + std::to_string(g.particle.SnapMax) + "].";
        const std::vector<size_t> &idxs = kv.second;
        const size_t nh = idxs.size();
        std::vector<double> coeffs(nh * P2);
        std::vector<float> centers(3 * nh), radii(nh);
        for (size_t j = 0; j < nh; j++) {
            for (size_t k = 0; k < P2; k++) coeffs[j * P2 + k] = floatCols[4 + k][idxs[j]];
            // boundingSpheres, cmd/stats.go:407-421
            for (int d = 0; d < 3; d++) centers[3 * j + d] = static_cast<float>(floatCols[d][idxs[j]]);
            radii[j] = static_cast<float>(floatCols[3][idxs[j]]);
        }
        // cmd/stats.go:271-289: Volume, R_sp, SurfaceArea, Axes, rangeSp
        std::vector<double> props(nh * SFK_SHELL_PROPS);
        if (sfk_shell_props(sfk, static_cast<int64_t>(nh), order, coeffs.data(), props.data()) != 0)
            return sfk_last_error(sfk);
        std::vector<double> rLow(nh), rHigh(nh);
        for (size_t j = 0; j < nh; j++) {
            const double *q = &props[j * SFK_SHELL_PROPS];
            const size_t o = idxs[j];
            cols[1][o] = q[0]; cols[2][o] = q[1]; cols[3][o] = q[2];
            cols[4][o] = q[3]; cols[5][o] = q[4]; cols[6][o] = q[5];
            cols[7][o] = q[6]; cols[8][o] = q[7]; cols[9][o] = q[8];
            cols[10][o] = q[9]; cols[11][o] = q[10];
            rLow[j] = q[9]; rHigh[j] = q[10];   // the second rangeSp call, :309-313
        }
        if (c.skipMass) continue;   // :316

        std::vector<Header> hds;
        std::vector<std::string> files;
        if (!(err = ReadHeaders(static_cast<int>(snap), *buf, cat, g.MemoDir, hds, files)).empty()) return err;
        if (sfk_mass_begin(sfk, static_cast<int64_t>(nh), order, centers.data(), coeffs.data(), rLow.data(), rHigh.data()) != 0)
            return sfk_last_error(sfk);
        // ShellParticleFile: IDs of the particles within ShellWidth * R200m of each surface, :326-341
        if (c.shellFilter && sfk_mass_set_band(sfk, c.shellWidth, radii.data(), rLow.data(), rHigh.data()) != 0)
            return sfk_last_error(sfk);
        for (size_t i = 0; i < hds.size(); i++) {
            // binSphereIntersections, :694-708: a block no R200m sphere touches is not read
            bool any = false;
            for (size_t j = 0; j < nh && !any; j++)
                any = sfk_sphere_block_hit(&centers[3 * j], radii[j], hds[i].Origin, hds[i].Width, hds[i].TotalWidth) == 1;
            if (!any) continue;
            const float *xs, *ms;
            int64_t np;
            if (!(err = buf->Read(files[i], &xs, &ms, &np)).empty()) return err;
            if (sfk_mass_push_block(sfk, xs, ms, np, hds[i].TotalWidth) != 0) return sfk_last_error(sfk);
            if (c.shellFilter) {
                int64_t nPairs = 0, nIds = 0;
                const int64_t *pairs = nullptr, *pids = nullptr;
                if (sfk_mass_take_band(sfk, &nPairs, &pairs) != 0) return sfk_last_error(sfk);
                if (nPairs > 0) {
                    if (!(err = buf->ReadIDs(files[i], &pids, &nIds)).empty()) return err;
                    for (int64_t k = 0; k < nPairs; k++) {
                        const int64_t halo = pairs[k] >> 32, idx = pairs[k] & 0xffffffffll;
                        if (idx >= nIds) return "read " + files[i] + ": fewer particle IDs than positions";
                        shellParticles[idxs[halo]].push_back(pids[idx]);
                    }
                }
            }
        }
        std::vector<double> masses(nh);
        if (sfk_mass_finish(sfk, masses.data()) != 0) return sfk_last_error(sfk);
        for (size_t j = 0; j < nh; j++) cols[0][idxs[j]] = masses[j];
    }

    if (c.shellFilter && !c.skipMass) {   // writeShellParticles, :344-346, 599-640
        if (!(err = WriteFilter(c.shellParticleFile, g.Endianness, snaps, ids, shellParticles)).empty()) return err;
    }

    // cmd/stats.go:342-358
    std::vector<int> ord(14);
    for (int i = 0; i < 14; i++) ord[i] = i;
    std::vector<std::string> lines = FormatCols({ids, snaps}, cols, ord);
    outLines.clear();
    outLines.push_back(CommentString(
        {"ID", "Snapshot"},
        {"M_sp [M_sun/h]", "R_sp [cMpc/h]", "Volume [cMpc^3/h^3]", "Surface Area [cMpc^2/h^2]", "Major Axis [cMpc/h]",
         "Intermediate Axis [cMpc/h]", "Minor Axis [cMpc/h]", "Ax", "Ay", "Az", "RMin [cMpc/h]", "RMax [cMpc/h]"},
        ord, std::vector<int>(14, 1)));
    outLines.insert(outLines.end(), lines.begin(), lines.end());
    return "";
}

// ---- `shellfish prof`, angular-fraction profiles ----------------------------------------

std::string ProfConfig::ReadConfig(const std::string &fname, const std::vector<std::string> &flags) {
    ConfigVars vars("prof.config");   // cmd/prof.go:103-151
    vars.Int(&bins, "Bins", 150);
    vars.Int(&order, "Order", 3);
    vars.Int(&samples, "Samples", 50 * 1000);
    vars.Float(&rMaxMult, "RMaxMult", 3.0);
    vars.Float(&rMinMult, "RMinMult", 0.03);
    vars.Int(&medianPixelLevel, "MedianPixelLevel", 3);
    vars.Float(&percentile, "Percentile", 50);
    vars.String(&pType, "ProfileType", "");
    std::string err;
    if (fname.empty()) {
        if (flags.empty()) return "";
        if (!(err = vars.ReadFlags(flags)).empty()) return err;
    } else {
        if (!(err = vars.ReadConfig(fname)).empty()) return err;
        if (!(err = vars.ReadFlags(flags)).empty()) return err;
    }
    if (pType.empty()) return "The variable 'ProfileType' was not set.";
    if (pType != "density" && pType != "median-density" && pType != "median-error" && pType != "contained-density" &&
        pType != "angular-fraction" && pType != "bound-density")
        return "The varaiable 'ProfileType' was set to '" + pType + "'.";   // sic, cmd/prof.go:147
    return validate();
}

std::string ProfConfig::validate() const {   // cmd/prof.go:153-169, with its copy-paste slips
    char buf[200];
    if (bins < 0) { snprintf(buf, sizeof buf, "The variable 'Bins' was set to %lld.", static_cast<long long>(bins)); return buf; }
    if (rMinMult <= 0) { snprintf(buf, sizeof buf, "The variable 'RMinMult' was set to %s.", go_g6(rMinMult).c_str()); return buf; }
    if (rMaxMult <= 0) { snprintf(buf, sizeof buf, "The variable 'RMinMult' was set to %s.", go_g6(rMinMult).c_str()); return buf; }
    if (medianPixelLevel < 0)
        return "The variable 'MedianPixelLevel' was set to %!g(int64=" + std::to_string(medianPixelLevel) + ").";
    return "";
}

const char *ProfConfig::ExampleConfig() {
    return "[prof.config]\n\n"
           "# ProfileType is required.  This build provides angular-fraction, the one type for which\n"
           "# the reference prints a catalog (cmd/prof.go:414 returns early for the particle-based types).\n"
           "ProfileType = angular-fraction\n\n"
           "# Optional (defaults shown, cmd/prof.go:106-112).  Samples is accepted for compatibility: the\n"
           "# fractions count one direction per equal-area stratum (1024 x 256) instead of random samples.\n"
           "Bins = 150\nOrder = 3\nSamples = 50000\nRMaxMult = 3.0\nRMinMult = 0.03";
}

std::string RunProf(const GlobalConfig &g, const ProfConfig &c, const std::string &stdinData, int device,
                    std::vector<std::string> &outLines) {
    (void)g;
    if (c.pType != "angular-fraction")
        return "ProfileType = " + c.pType + " is not provided by this build: the reference's prof mode returns "
               "without output for the particle-based profile types (cmd/prof.go:414).";
    // cmd/prof.go:227-262
    const size_t P2 = static_cast<size_t>(2 * c.order * c.order);
    std::vector<int> floatIdx(4 + P2);
    for (size_t i = 0; i < floatIdx.size(); i++) floatIdx[i] = static_cast<int>(i) + 2;
    std::vector<std::vector<int64_t>> intCols;
    std::vector<std::vector<double>> floatCols;
    std::string err = Parse(stdinData, {0, 1}, floatIdx, intCols, floatCols);
    if (!err.empty()) return err;
    if (intCols.empty() || intCols[0].empty()) return "No input IDs.";
    const std::vector<int64_t> &ids = intCols[0], &snaps = intCols[1];
    const size_t n = ids.size();
    const int bins = static_cast<int>(c.bins);
    if (c.order <= 0 || c.order > 4) return "Order above 4 is not supported.";

    std::vector<double> coeffs(n * P2), rMin(n), rMax(n);
    for (size_t i = 0; i < n; i++) {
        for (size_t k = 0; k < P2; k++) coeffs[i * P2 + k] = floatCols[4 + k][i];
        rMin[i] = floatCols[3][i] * c.rMinMult;   // cmd/prof.go:641-642
        rMax[i] = floatCols[3][i] * c.rMaxMult;
    }
    std::vector<double> rs(n * std::max(bins, 0)), fs(rs.size());
    if (bins > 0) {
        sfk_params p;
        sfk_default_params(&p);
        p.rings = 3; p.device = device;
        sfk_ctx *sfk = nullptr;
        struct Guard { sfk_ctx *&c; ~Guard() { if (c) sfk_destroy(c); } } guard{sfk};
        if (sfk_create(&sfk, &p) != 0) return sfk_last_error(sfk);
        if (sfk_angular_fraction(sfk, static_cast<int64_t>(n), static_cast<int32_t>(c.order), coeffs.data(), rMin.data(),
                                 rMax.data(), bins, rs.data(), fs.data()) != 0)
            return sfk_last_error(sfk);
    }
    // angularFractionMain, cmd/prof.go:628-661
    std::vector<std::vector<double>> cols(2 * static_cast<size_t>(std::max(bins, 0)), std::vector<double>(n));
    for (int j = 0; j < bins; j++)
        for (size_t i = 0; i < n; i++) {
            cols[j][i] = rs[i * bins + j];
            cols[bins + j][i] = fs[i * bins + j];
        }
    std::vector<int> ord(cols.size() + 2);
    for (size_t i = 0; i < ord.size(); i++) ord[i] = static_cast<int>(i);
    std::vector<std::string> lines = FormatCols({ids, snaps}, cols, ord);
    outLines.clear();
    outLines.push_back(CommentString({"ID", "Snapshot", "R [cMpc/h]", "Volume Fraction Contained"}, {}, {0, 1, 2, 3},
                                     {1, 1, bins, bins}));
    outLines.insert(outLines.end(), lines.begin(), lines.end());
    return "";
}

}  // namespace sfhost
