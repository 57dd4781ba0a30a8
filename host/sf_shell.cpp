// sf_shell.cpp -- see sf_shell.hpp.
#include "sf_shell.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <condition_variable>
#include <map>
#include <mutex>
#include <thread>
#include <sys/stat.h>

#include "../include/shellfish_b200.h"
#include "sf_catalog.hpp"
#include "sf_config.hpp"

namespace sfhost {

const char *kSourceVersion = "1.0.4";

namespace {

bool parse_version(const std::string &s, int v[3]) {   // version/version.go:16-44
    size_t a = s.find('.'), b = a == std::string::npos ? a : s.find('.', a + 1);
    if (a == std::string::npos || b == std::string::npos || s.find('.', b + 1) != std::string::npos) return false;
    int64_t x[3];
    if (!go_atoi(s.substr(0, a), x[0]) || !go_atoi(s.substr(a + 1, b - a - 1), x[1]) || !go_atoi(s.substr(b + 1), x[2]))
        return false;
    for (int i = 0; i < 3; i++) { if (x[i] < 0) return false; v[i] = static_cast<int>(x[i]); }
    return true;
}

std::string validate_dir(const std::string &name) {   // cmd/cmd.go:372-381 (creates a missing directory)
    struct stat st;
    if (stat(name.c_str(), &st) != 0) {
        std::string cur;
        for (size_t i = 0; i <= name.size(); i++) {
            if (i == name.size() || name[i] == '/') {
                if (!cur.empty() && stat(cur.c_str(), &st) != 0 && mkdir(cur.c_str(), 0755) != 0)
                    return "mkdir " + cur + ": " + std::strerror(errno);
            }
            if (i < name.size()) cur += name[i];
        }
        return "";
    }
    if (!S_ISDIR(st.st_mode)) return name + " is not a directory.";
    return "";
}

std::string strip_spaces(const std::string &s) {
    std::string o;
    for (char c : s) if (c != ' ') o += c;
    return o;
}

bool in_list(const std::string &x, const std::vector<std::string> &xs) {
    return std::find(xs.begin(), xs.end(), x) != xs.end();
}

}  // namespace

std::string GlobalConfig::ReadConfig(const std::string &fname) {
    ConfigVars vars("config");
    vars.String(&Version, "Version", kSourceVersion);
    vars.String(&particle.SnapshotFormat, "SnapshotFormat", "");
    vars.String(&SnapshotType, "SnapshotType", "");
    vars.String(&HaloDir, "HaloDir", "");
    vars.String(&HaloType, "HaloType", "nil");
    vars.String(&TreeDir, "TreeDir", "");
    vars.String(&TreeType, "TreeType", "nil");
    vars.String(&MemoDir, "MemoDir", "");
    vars.Strings(&HaloValueNames, "HaloValueNames", {});
    vars.Ints(&HaloValueColumns, "HaloValueColumns", {});
    vars.Strings(&HaloValueComments, "HaloValueComments", {});
    vars.String(&HaloPositionUnits, "HaloPositionUnits", "");
    vars.String(&HaloRadiusUnits, "HaloRadiusUnits", "");
    vars.String(&HaloMassUnits, "HaloMassUnits", "");
    vars.Strings(&particle.SnapshotFormatMeanings, "SnapshotFormatMeanings", {});
    vars.String(&particle.ScaleFactorFile, "ScaleFactorFile", "");
    vars.Ints(&particle.BlockMins, "BlockMins", {});
    vars.Ints(&particle.BlockMaxes, "BlockMaxes", {});
    vars.Int(&particle.SnapMin, "SnapMin", -1);
    vars.Int(&particle.SnapMax, "SnapMax", -1);
    vars.String(&Endianness, "Endianness", "SystemOrder");
    vars.Bool(&ValidateFormats, "ValidateFormats", false);
    vars.Int(&Threads, "Threads", -1);
    vars.String(&Logging, "Logging", "nil");
    vars.Ints(&GadgetDMTypeIndices, "GadgetDMTypeIndices", {1});
    vars.Ints(&GadgetSingleMassIndices, "GadgetSingleMassIndices", {1});
    vars.Float(&GadgetPositionUnits, "GadgetPositionUnits", 1.0);
    vars.Float(&GadgetMassUnits, "GadgetMassUnits", 1.0);
    vars.Int(&LGadgetNpartNum, "LGadgetNpartNum", 2);
    vars.Float(&NilSnapOmegaM, "NilSnapOmegaM", -1);
    vars.Float(&NilSnapOmegaL, "NilSnapOmegaL", -1);
    vars.Float(&NilSnapH100, "NilSnapH100", -1);
    vars.Floats(&NilSnapScaleFactors, "NilSnapScaleFactors", {});
    vars.Float(&NilSnapTotalWidth, "NilSnapTotalWidth", -1);
    std::string err = vars.ReadConfig(fname);
    if (!err.empty()) return err;
    return validate();
}

std::string GlobalConfig::validate() {   // cmd/cmd.go:148-356
    int v[3], sv[3];
    if (!parse_version(Version, v))
        return "I couldn't parse the 'Version' variable: Version string does not take the form of three "
               "period-separated non-negative numbers";
    parse_version(kSourceVersion, sv);
    if ((sv[0] == 0 || v[0] == 0) && (v[0] != sv[0] || v[1] != sv[1] || v[2] != sv[2]))
        return "The 'Version' variable is set to " + Version + ", but the version of the source is " + kSourceVersion;
    if (sv[0] == 1 && v[0] == 1 && (v[1] > sv[1] || v[2] > sv[2]))
        return "You are attempting to run a config file from the future: The 'Version' variable is set to " +
               Version + ", but the version of the source code is " + kSourceVersion;
    static const std::vector<std::string> snapTypes{"gotetra", "LGadget-2", "Gadget-2", "ARTIO", "Bolshoi", "BolshoiP", "nil"};
    if (SnapshotType.empty()) return "The 'SnapshotType variable isn't set.'";
    if (!in_list(SnapshotType, snapTypes))
        return "The 'SnapshotType' variable is set to '" + SnapshotType + "', which I don't recognize.";
    if (HaloType.empty()) return "The 'HaloType' variable isn't set.'";
    if (HaloType != "Text" && HaloType != "nil")
        return "The 'HaloType' variable is set to '" + HaloType + "', which I don't recognize.";
    if (HaloType != "nil") {
        static const std::vector<std::string> units{"cMpc/h", "ckpc/h", "pMpc/h", "pkpc/h", "cMpc", "ckpc", "pMpc", "pkpc"};
        HaloPositionUnits = strip_spaces(HaloPositionUnits);
        if (HaloPositionUnits.empty()) return "The 'HaloPositionUnits' variable isn't set.";
        if (!in_list(HaloPositionUnits, units))
            return "The 'HaloPositionUnits variable is set to '" + HaloPositionUnits +
                   "', which I don't support. Only supported units are ckpc/h and cMpc/h";
        HaloRadiusUnits = strip_spaces(HaloRadiusUnits);
        if (HaloRadiusUnits.empty()) return "The 'HaloRadiusUnits' variable isn't set.";
        if (!in_list(HaloRadiusUnits, units))
            return "The 'HaloRadiusUnits variable is set to '" + HaloPositionUnits +
                   "', which I don't support. Only supported units are ckpc/h, cMpc/h, pkpc/h, pMpc/h, ckpc, cMpc, "
                   "pkpc, and pMpc";
        HaloMassUnits = strip_spaces(HaloMassUnits);
        if (HaloMassUnits.empty()) return "The 'HaloMassUnits' variable isn't set.";
        if (HaloMassUnits != "Msun/h")
            return "The 'HaloMassUnits variable is set to '" + HaloPositionUnits + "', which I don't understand.";
    } else {
        HaloRadiusUnits = HaloPositionUnits = "cMpc/h";
        HaloMassUnits = "Msun/h";
    }
    if (TreeType.empty()) return "The 'TreeType variable isn't set.'";
    if (TreeType != "consistent-trees" && TreeType != "nil")
        return "The 'TreeType' variable is set to '" + TreeType + "', which I don't recognize.";
    std::string err;
    if (HaloType != "nil") {
        if (HaloDir.empty()) return "The 'HaloDir' variable isn't set.";
        if (!(err = validate_dir(HaloDir)).empty()) return "The 'HaloDir' variable is set to '" + HaloDir + "', but " + err;
        if (TreeDir.empty()) return "The 'TreeDir' variable isn't set.";
        if (!(err = validate_dir(TreeDir)).empty()) return "The 'TreeDir' variable is set to '" + TreeDir + "', but " + err;
    }
    if (MemoDir.empty()) return "The 'MemoDir' variable isn't set.";
    if (!(err = validate_dir(MemoDir)).empty()) return "The 'MemoDir' variable is set to '" + MemoDir + "', but " + err;
    if (HaloType != "nil") {
        if (HaloValueNames.empty()) return "The 'HaloValueNames' variable isn't set.";
        if (HaloValueColumns.empty()) return "The 'HaloValueColumns' variable isn't set.";
        for (const char *need : {"ID", "X", "Y", "Z", "M200m"})
            if (!in_list(need, HaloValueNames))
                return std::string("'HaloValueNames' does not contain the '") + need + "' name.";
    }
    if (Endianness.empty()) return "The variable 'Endianness' was not set.";
    if (Endianness != "LittleEndian" && Endianness != "BigEndian" && Endianness != "SystemOrder")
        return "The variable 'Endianness' must be sent to either 'SystemOrder', 'LittleEndian', or 'BigEndian'.";
    char buf[256];
    if (HaloValueNames.size() != HaloValueColumns.size()) {
        snprintf(buf, sizeof buf, "len(HaloValueNames) = %zu, but len(HaloValueColumns = %zu)", HaloValueNames.size(),
                 HaloValueColumns.size());
        return buf;
    }
    if (HaloValueNames.size() != HaloValueComments.size()) {
        snprintf(buf, sizeof buf, "len(HaloValueNames) = %zu, but len(HaloValueComments = %zu)", HaloValueNames.size(),
                 HaloValueComments.size());
        return buf;
    }
    if (LGadgetNpartNum > 2 || LGadgetNpartNum <= 0) {
        snprintf(buf, sizeof buf, "GadgetNpartNum set to %lld, but the only valid values are 1 and 2.",
                 static_cast<long long>(LGadgetNpartNum));
        return buf;
    }
    return "";
}

std::string ShellConfig::ReadConfig(const std::string &fname, const std::vector<std::string> &flags) {
    ConfigVars vars("shell.config");   // cmd/shell.go:102-141
    vars.Int(&subsampleFactor, "SubsampleFactor", 1);
    vars.Int(&radialBins, "RadialBins", 256);
    vars.Int(&spokes, "Spokes", 256);
    vars.Int(&rings, "Rings", 100);
    vars.Float(&rMaxMult, "RMaxMult", 3);
    vars.Float(&rMinMult, "RMinMult", 0.3);
    vars.Float(&rKernelMult, "RKernelMult", 0.2);
    vars.Float(&eta, "Eta", 10);
    vars.Int(&order, "Order", 3);
    vars.Int(&levels, "Levels", 3);
    vars.Int(&smoothingWindow, "SmoothingWindow", 121);
    vars.Float(&losSlopeCutoff, "LOSSlopeCutoff", 0.0);
    vars.Float(&backgroundRhoMult, "BackgroundRhoMult", 0.5);
    vars.Bool(&percentileProfile, "PercentileProfile", false);
    vars.Float(&percentile, "Percentile", 50.0);
    std::string err;
    if (fname.empty()) {
        if (flags.empty()) return "";
        if (!(err = vars.ReadFlags(flags)).empty()) return err;
        return validate();
    }
    if (!(err = vars.ReadConfig(fname)).empty()) return err;
    if (!(err = vars.ReadFlags(flags)).empty()) return err;
    return validate();
}

std::string ShellConfig::validate() const {   // cmd/shell.go:143-187
    char buf[256];
    auto bi = [&](const char *n, int64_t v) { snprintf(buf, sizeof buf, "The variable '%s' was set to %lld.", n, static_cast<long long>(v)); return std::string(buf); };
    auto bf = [&](const char *n, double v) { snprintf(buf, sizeof buf, "The variable '%s' was set to %g.", n, v); return std::string(buf); };
    if (subsampleFactor <= 0) return bi("SubsampleFactor", subsampleFactor);
    if (radialBins <= 0) return bi("RadialBins", radialBins);
    if (spokes <= 0) return bi("Spokes", spokes);
    if (rings <= 0) return bi("Rings", rings);
    if (rMaxMult <= 0) return bf("RMaxMult", rMaxMult);
    if (rMinMult <= 0) return bf("RMinMult", rMinMult);
    if (rKernelMult <= 0) return bf("RKernelMult", rKernelMult);
    if (eta <= 0) return bf("Eta", eta);
    if (order <= 0) return bi("Order", order);
    if (levels <= 0) return bi("Levels", levels);
    if (smoothingWindow <= 0) return bi("SmoothingWindow", smoothingWindow);
    if (rMinMult >= rMaxMult) {
        snprintf(buf, sizeof buf, "The variable '%s' was set to %g, but the variable '%s' was set to %g.", "RMinMult",
                 rMinMult, "RMaxMult", rMaxMult);
        return buf;
    }
    return "";
}

const char *ShellConfig::ExampleConfig() {
    return "[shell.config]\n\n"
           "# All fields are optional; the values shown are the defaults (cmd/shell.go:105-119).\n"
           "SubsampleFactor = 1\nRadialBins = 256\nSpokes = 256\nRings = 100\nRMaxMult = 3.0\nRMinMult = 0.3\n"
           "RKernelMult = 0.2\nEta = 10.0\nOrder = 3\nLevels = 3\nSmoothingWindow = 121\nLOSSlopeCutoff = 0.0\n"
           "BackgroundRhoMult = 0.5";
}

void PartitionHalos(const std::vector<double> &weights, const std::vector<int64_t> &groups, int nRanks,
                    std::vector<int> &owner) {
    // group weights, groups in ascending label order (numpy.unique)
    std::map<int64_t, double> gw;
    for (size_t i = 0; i < weights.size(); i++) gw[groups[i]] += weights[i];
    std::vector<std::pair<int64_t, double>> gs(gw.begin(), gw.end());
    std::vector<size_t> order(gs.size());
    for (size_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return gs[a].second > gs[b].second; });
    std::vector<double> loads(std::max(nRanks, 1), 0.0);
    std::map<int64_t, int> gOwner;
    for (size_t k : order) {
        int r = 0;
        for (int q = 1; q < nRanks; q++) if (loads[q] < loads[r]) r = q;   // first minimum, like argmin
        gOwner[gs[k].first] = r;
        loads[r] += gs[k].second;
    }
    owner.resize(weights.size());
    for (size_t i = 0; i < weights.size(); i++) owner[i] = gOwner[groups[i]];
}

void PartitionHalosSpatial(const std::vector<double> &weights, const std::vector<double> &x, const std::vector<double> &y,
                           const std::vector<double> &z, double totalWidth, int nRanks, std::vector<int> &owner) {
    const size_t n = weights.size();
    owner.assign(n, 0);
    if (n == 0 || nRanks <= 1) return;
    const int bits = 10;
    std::vector<int64_t> key(n);
    for (size_t i = 0; i < n; i++) {
        const double c[3] = {x[i], y[i], z[i]};
        int64_t q[3], k = 0;
        for (int ax = 0; ax < 3; ax++) {
            const double v = std::floor(c[ax] / totalWidth * static_cast<double>(1 << bits));
            q[ax] = static_cast<int64_t>(std::min(std::max(v, 0.0), static_cast<double>((1 << bits) - 1)));
        }
        for (int b = 0; b < bits; b++)
            for (int ax = 0; ax < 3; ax++) k |= ((q[ax] >> b) & 1) << (3 * b + (2 - ax));
        key[i] = k;
    }
    std::vector<size_t> order(n);
    for (size_t i = 0; i < n; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return key[a] < key[b]; });
    std::vector<double> cum(n);
    double run = 0;
    for (size_t k = 0; k < n; k++) { run += weights[order[k]]; cum[k] = run; }   // numpy.cumsum: sequential
    const double total = run > 0 ? run : 1.0;
    for (size_t k = 0; k < n; k++) {
        const double wv = weights[order[k]];
        const int64_t r = static_cast<int64_t>((cum[k] - wv / 2) / total * static_cast<double>(nRanks));
        owner[order[k]] = static_cast<int>(std::min<int64_t>(r, nRanks - 1));
    }
}

std::string RunShell(const GlobalConfig &g, const ShellConfig &c, const std::string &stdinData, uint64_t seed,
                     const std::vector<int> &devices, std::vector<std::string> &outLines) {
    if (devices.empty()) return "No GPU selected.";
    const int G = static_cast<int>(devices.size());
    // catalog.Parse(stdin, {0,1}, {2,3,4,5}), cmd/shell.go:209-218
    std::vector<std::vector<int64_t>> intCols;
    std::vector<std::vector<double>> coords;
    std::string err = Parse(stdinData, {0, 1}, {2, 3, 4, 5}, intCols, coords);
    if (!err.empty()) return err;
    const std::vector<int64_t> &ids = intCols[0], &snaps = intCols[1];
    if (ids.empty()) return "No input IDs.";
    // cmd/shell.go:223-231 sizes percentile rows as RadialBins, but haloAnalysis replaces each
    // processed row by calcPercentile's 2*RadialBins slice (:516,:632): radii then densities
    const size_t rowLen = c.percentileProfile ? static_cast<size_t>(2 * c.radialBins) : static_cast<size_t>(c.order * c.order * 2);
    std::vector<std::vector<double>> out(ids.size(), std::vector<double>(rowLen, 0.0));

    Catalogs cat;
    if (!(err = cat.Init(g.particle, g.SnapshotType == "gotetra")).empty()) return err;
    Context ctx;
    ctx.LGadgetNPartNum = g.LGadgetNpartNum;
    ctx.GadgetDMTypeIndices = g.GadgetDMTypeIndices;
    ctx.GadgetDMSingleMassIndices = g.GadgetSingleMassIndices;
    ctx.GadgetMassUnits = g.GadgetMassUnits;
    ctx.GadgetPositionUnits = g.GadgetPositionUnits;
    // loop(), cmd/shell.go:288-372
    std::map<int64_t, std::vector<size_t>> idxBins;   // binBySnap, sorted like sort.Ints
    for (size_t i = 0; i < snaps.size(); i++) idxBins[snaps[i]].push_back(i);
    // The reference opens snaps[0] (:234) and reads the headers of sortedSnaps[0] (:305) before it
    // skips Snap == -1 rows (:325), i.e. it indexes the file table with -1 and panics whenever such
    // a row is present.  Here the first real snapshot serves both purposes.
    int64_t firstSnap = -1;
    for (const auto &kv : idxBins)
        if (kv.first != -1) { firstSnap = kv.first; break; }
    if (firstSnap == -1) return "Every row of the halo table has Snapshot = -1: nothing to do.";
    if (firstSnap < g.particle.SnapMin || firstSnap > g.particle.SnapMax)
        return "Snapshot " + std::to_string(firstSnap) + " is outside [SnapMin, SnapMax] = [" +
               std::to_string(g.particle.SnapMin) + ", " + std::to_string(g.particle.SnapMax) + "].";
    std::unique_ptr<VectorBuffer> buf, buf2;
    if (!(err = NewVectorBuffer(g.SnapshotType, cat.ParticleCatalog(static_cast<int>(firstSnap), 0), g.Endianness, ctx, buf)).empty())
        return err;
    std::vector<Header> hds0;
    std::vector<std::string> files;
    if (!(err = ReadHeaders(static_cast<int>(firstSnap), *buf, cat, g.MemoDir, hds0, files)).empty()) return err;
    if (hds0.empty()) return "The snapshot has no blocks.";
    const float minMass = buf->MinMass();
    const bool uniformMass = buf->UniformMass();   // LGadget-2, gotetra: ms[i] == minMass, not transferred

    sfk_params p;
    sfk_default_params(&p);
    p.radial_bins = c.radialBins; p.spokes = c.spokes; p.rings = c.rings;
    p.r_max_mult = c.rMaxMult; p.r_min_mult = c.rMinMult; p.r_kernel_mult = c.rKernelMult;
    p.eta = c.eta; p.order = c.order; p.smoothing_window = c.smoothingWindow; p.levels = c.levels;
    p.subsample_factor = c.subsampleFactor; p.los_slope_cutoff = c.losSlopeCutoff;
    p.background_rho_mult = c.backgroundRhoMult; p.percentile_profile = c.percentileProfile;
    p.percentile = c.percentile; p.seed = seed;
    const auto tCreate = std::chrono::steady_clock::now();
    std::vector<sfk_ctx *> ctxs(G, nullptr);
    struct Guard { std::vector<sfk_ctx *> &c; ~Guard() { for (sfk_ctx *x : c) if (x) sfk_destroy(x); } } guard{ctxs};
    {   // CUDA context creation takes a second per device: one thread each
        std::vector<int> rcs(G, 0);
        std::vector<std::thread> th;
        for (int d = 0; d < G; d++)
            th.emplace_back([&, d] {
                sfk_params q = p;
                q.device = devices[d];
                rcs[d] = sfk_create(&ctxs[d], &q);
            });
        for (std::thread &t : th) t.join();
        for (int d = 0; d < G; d++)
            if (rcs[d] != 0) return ctxs[d] ? sfk_last_error(ctxs[d]) : "sfk_create failed";
    }
    sfk_ctx *sfk = ctxs[0];
    const sfk_cosmo cosmo0{hds0[0].Cosmo.Z, hds0[0].Cosmo.OmegaM, hds0[0].Cosmo.OmegaL, hds0[0].Cosmo.H100};

    // Logging = performance (cmd/shell.go:352-367): elapsed time after sphereLoop and after
    // haloAnalysis of every snapshot, on stderr like Go's log package; the time spent inside
    // buf.Read (file I/O) and the device stage times are reported next to it.
    const bool perf = g.Logging == "performance";
    const auto tStart = std::chrono::steady_clock::now();
    auto since = [&](std::chrono::steady_clock::time_point t0) {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    };
    if (perf) std::fprintf(stderr, "Setup: %.6gs creating %d GPU context(s)\n", since(tCreate), G);
    for (const auto &kv : idxBins) {
        const int64_t snap = kv.first;
        if (snap == -1) continue;
        if (snap < g.particle.SnapMin || snap > g.particle.SnapMax)
            return "Snapshot " + std::to_string(snap) + " is outside [SnapMin, SnapMax] = [" +
                   std::to_string(g.particle.SnapMin) + ", " + std::to_string(g.particle.SnapMax) + "].";
        const std::vector<size_t> &idxs = kv.second;
        double readSeconds = 0.0;
        int64_t blocksRead = 0, particlesRead = 0;
        std::vector<double> x(idxs.size()), y(idxs.size()), z(idxs.size()), r(idxs.size());
        for (size_t i = 0; i < idxs.size(); i++) {
            x[i] = coords[0][idxs[i]]; y[i] = coords[1][idxs[i]]; z[i] = coords[2][idxs[i]]; r[i] = coords[3][idxs[i]];
        }
        for (sfk_ctx *cx : ctxs) {
            if (sfk_begin_snapshot(cx, &cosmo0, minMass) != 0) return sfk_last_error(cx);
            if (sfk_add_halos(cx, static_cast<int64_t>(idxs.size()), x.data(), y.data(), z.data(), r.data()) != 0)
                return sfk_last_error(cx);
        }
        // sphereLoop, cmd/shell.go:376-415
        std::vector<Header> hds;
        if (!(err = ReadHeaders(static_cast<int>(snap), *buf, cat, g.MemoDir, hds, files)).empty()) return err;
        // several GPUs: halos partitioned by estimated particle count (R200m^3 plus a constant
        // for the per-halo analysis) into equal-weight runs along a Z-curve, so that a GPU's
        // halos are spatially compact and touch few blocks
        std::vector<int> owner(idxs.size(), 0);
        if (G > 1) {
            std::vector<double> w(idxs.size());
            double mean = 0;
            for (size_t i = 0; i < idxs.size(); i++) { w[i] = r[i] > 0 ? r[i] * r[i] * r[i] : 0.0; mean += w[i]; }
            mean = idxs.empty() ? 1.0 : std::max(mean / static_cast<double>(idxs.size()), 1e-300);
            for (size_t i = 0; i < idxs.size(); i++) w[i] = w[i] / mean + 0.25;
            PartitionHalosSpatial(w, x, y, z, hds.empty() ? 1.0 : hds[0].TotalWidth, G, owner);
            std::vector<uint8_t> mine(idxs.size());
            for (int d = 0; d < G; d++) {
                for (size_t i = 0; i < idxs.size(); i++) mine[i] = owner[i] == d;
                if (sfk_set_owned(ctxs[d], static_cast<int64_t>(mine.size()), mine.data()) != 0) return sfk_last_error(ctxs[d]);
            }
        }
        // binIntersections first (cmd/shell.go:599-611): the blocks some halo touches, and for each
        // the GPUs that own such a halo
        std::vector<uint8_t> hit(idxs.size());
        std::vector<size_t> needed;
        std::vector<std::vector<char>> wantsOf;
        for (size_t i = 0; i < hds.size(); i++) {
            int64_t nHit = 0;
            if (sfk_block_halo_mask(sfk, hds[i].Origin, hds[i].Width, hds[i].TotalWidth, hit.data(), &nHit) != 0)
                return sfk_last_error(sfk);
            if (nHit == 0) continue;
            std::vector<char> wants(G, 0);
            for (size_t k = 0; k < hit.size(); k++) if (hit[k]) wants[owner[k]] = 1;
            needed.push_back(i);
            wantsOf.push_back(wants);
        }
        // Every needed block is read once, by a reader thread that runs one block ahead of the
        // pushes through two VectorBuffers (a buffer's arrays are reused by its next Read,
        // io/io.go:57-67), and is pushed to every GPU that wants it.
        if (!buf2 && !(err = NewVectorBuffer(g.SnapshotType, cat.ParticleCatalog(static_cast<int>(firstSnap), 0),
                                             g.Endianness, ctx, buf2)).empty())
            return err;
        struct Slot { const float *xs = nullptr, *ms = nullptr; int64_t n = 0; std::string err; double seconds = 0; };
        Slot slots[2];
        std::mutex mu;
        std::condition_variable cv;
        size_t produced = 0, consumed = 0;   // blocks read / blocks pushed
        bool stop = false;
        VectorBuffer *bufs[2] = {buf.get(), buf2.get()};
        std::thread reader([&] {
            for (size_t k = 0; k < needed.size(); k++) {
                {
                    std::unique_lock<std::mutex> lk(mu);
                    cv.wait(lk, [&] { return stop || k < consumed + 2; });
                    if (stop) return;
                }
                Slot &s = slots[k & 1];
                const auto t0 = std::chrono::steady_clock::now();
                s.err = bufs[k & 1]->Read(files[needed[k]], &s.xs, &s.ms, &s.n);
                s.seconds = since(t0);
                std::lock_guard<std::mutex> lk(mu);
                produced = k + 1;
                cv.notify_all();
                if (!s.err.empty()) return;
            }
        });
        struct Joiner { std::thread &t; std::mutex &m; std::condition_variable &c; bool &stop;
                        ~Joiner() { { std::lock_guard<std::mutex> lk(m); stop = true; } c.notify_all(); if (t.joinable()) t.join(); } }
            joiner{reader, mu, cv, stop};
        for (size_t k = 0; k < needed.size(); k++) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return produced > k; });
            }
            Slot &s = slots[k & 1];
            if (!s.err.empty()) return s.err;
            readSeconds += s.seconds; blocksRead++; particlesRead += s.n;
            const Header &h = hds[needed[k]];
            const sfk_cosmo hc{h.Cosmo.Z, h.Cosmo.OmegaM, h.Cosmo.OmegaL, h.Cosmo.H100};
            for (int d = 0; d < G; d++)
                if (wantsOf[k][d] && sfk_push_block(ctxs[d], s.xs, uniformMass ? nullptr : s.ms, s.n, &hc, h.TotalWidth, h.Origin,
                                                    h.Width) != 0)
                    return sfk_last_error(ctxs[d]);
            std::lock_guard<std::mutex> lk(mu);
            consumed = k + 1;
            cv.notify_all();
        }
        if (perf) {
            std::fprintf(stderr, "Snap %lld, sphereLoop ended\nTime: %.6gs\n", static_cast<long long>(snap), since(tStart));
            std::fprintf(stderr, "Read: %.6gs in buf.Read, overlapped with the pushes (%lld blocks, %lld particles)\n", readSeconds,
                         static_cast<long long>(blocksRead), static_cast<long long>(particlesRead));
        }
        // haloAnalysis, cmd/shell.go:502-528: a failed fit leaves the zero row
        std::vector<std::vector<double>> cs(G, std::vector<double>(idxs.size() * rowLen));
        std::vector<std::vector<int32_t>> status(G, std::vector<int32_t>(idxs.size()));
        std::vector<int> rcs(G, 0);
        if (G == 1) {
            rcs[0] = sfk_finish_snapshot(sfk, cs[0].data(), status[0].data());
        } else {   // one host thread per GPU
            std::vector<std::thread> th;
            for (int d = 0; d < G; d++)
                th.emplace_back([&, d] { rcs[d] = sfk_finish_snapshot(ctxs[d], cs[d].data(), status[d].data()); });
            for (std::thread &t : th) t.join();
        }
        for (int d = 0; d < G; d++) if (rcs[d] != 0) return sfk_last_error(ctxs[d]);
        for (size_t i = 0; i < idxs.size(); i++) {
            const std::vector<double> &src = cs[owner[i]];
            std::copy(src.begin() + i * rowLen, src.begin() + (i + 1) * rowLen, out[idxs[i]].begin());
        }
        if (perf) {
            std::fprintf(stderr, "Snap %lld, haloAnalysis ended\nTime: %.6gs\n", static_cast<long long>(snap), since(tStart));
            sfk_stats st{};
            for (int d = 0; d < G; d++)
            if (sfk_get_stats(ctxs[d], &st) == 0)
                std::fprintf(stderr, "Device: cull %.3f ms, hits %.3f ms, bin %.3f ms, los %.3f ms, filter %.3f ms, fit %.3f ms; "
                                     "%lld candidates, %lld intersections, %lld launches\n",
                             st.ms_cull, st.ms_hits, st.ms_bin, st.ms_los, st.ms_filter, st.ms_fit,
                             static_cast<long long>(st.n_candidates), static_cast<long long>(st.n_intersections),
                             static_cast<long long>(st.launches));
        }
    }

    // cmd/shell.go:246-262
    std::vector<std::vector<double>> floatCols = coords;
    for (size_t k = 0; k < rowLen; k++) {
        std::vector<double> col(ids.size());
        for (size_t i = 0; i < ids.size(); i++) col[i] = out[i][k];
        floatCols.push_back(col);
    }
    std::vector<int> order(2 + 4 + rowLen);
    for (size_t i = 0; i < order.size(); i++) order[i] = static_cast<int>(i);
    std::vector<std::string> lines = FormatCols({ids, snaps}, floatCols, order);
    outLines.clear();
    outLines.push_back(CommentString({"ID", "Snapshot"}, {"X [cMpc/h]", "Y [cMpc/h]", "Z [cMpc/h]", "R200m [cMpc/h]", "P_ijk"},
                                     {0, 1, 2, 3, 4, 5, 6}, {1, 1, 1, 1, 1, 1, static_cast<int>(rowLen)}));
    outLines.insert(outLines.end(), lines.begin(), lines.end());
    if (perf) std::fprintf(stderr, "Catalog formatted\nTime: %.6gs\n", since(tStart));
    {   // freeing tens of GB per device takes a while: one thread per context
        std::vector<std::thread> th;
        for (int d = 0; d < G; d++) th.emplace_back([&, d] { sfk_destroy(ctxs[d]); ctxs[d] = nullptr; });
        for (std::thread &t : th) t.join();
    }
    if (perf) std::fprintf(stderr, "Contexts released\nTime: %.6gs\n", since(tStart));
    return "";
}

}  // namespace sfhost
