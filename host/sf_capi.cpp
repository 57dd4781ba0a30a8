// sf_capi.cpp -- C entry points over the host modules for the CPU test tier
// (ctypes).  No CUDA here: libsfhost.so loads without a GPU.
#include <cstring>
#include <string>
#include <vector>

#include "sf_catalog.hpp"
#include "sf_config.hpp"
#include "sf_shell.hpp"
#include "sf_stats.hpp"
#include "sf_snapshot.hpp"

using namespace sfhost;

namespace { std::string g_text; const char *keep(const std::string &s) { g_text = s; return g_text.c_str(); } }

extern "C" {

const char *sfh_go_g6(double v) { return keep(go_g6(v)); }

// names: '\n'-separated; order/sizes: int arrays
const char *sfh_comment_string(const char *intNames, const char *floatNames, const int *order, int nOrder,
                               const int *sizes, int nSizes) {
    auto split = [](const char *s) {
        std::vector<std::string> out;
        std::string cur;
        for (const char *p = s; *p; p++) { if (*p == '\n') { out.push_back(cur); cur.clear(); } else cur += *p; }
        if (!cur.empty()) out.push_back(cur);
        return out;
    };
    return keep(CommentString(split(intNames), split(floatNames), std::vector<int>(order, order + nOrder),
                              std::vector<int>(sizes, sizes + nSizes)));
}

// ints: [nInt][rows], floats: [nFloat][rows]; returns '\n'-joined lines
const char *sfh_format_cols(const int64_t *ints, int nInt, const double *floats, int nFloat, int rows,
                            const int *order, int nOrder) {
    std::vector<std::vector<int64_t>> ic(nInt);
    std::vector<std::vector<double>> fc(nFloat);
    for (int i = 0; i < nInt; i++) ic[i].assign(ints + static_cast<size_t>(i) * rows, ints + static_cast<size_t>(i + 1) * rows);
    for (int i = 0; i < nFloat; i++) fc[i].assign(floats + static_cast<size_t>(i) * rows, floats + static_cast<size_t>(i + 1) * rows);
    std::string out;
    for (const std::string &l : FormatCols(ic, fc, std::vector<int>(order, order + nOrder))) { out += l; out += '\n'; }
    return keep(out);
}

// Parses `data`; writes up to cap rows of each requested column; returns rows or -1 (error text via sfh_last_text)
int sfh_parse_catalog(const char *data, const int *icols, int nI, const int *fcols, int nF, int64_t *iout, double *fout, int cap) {
    std::vector<std::vector<int64_t>> ic;
    std::vector<std::vector<double>> fc;
    std::string err = Parse(data, std::vector<int>(icols, icols + nI), std::vector<int>(fcols, fcols + nF), ic, fc);
    if (!err.empty()) { keep(err); return -1; }
    const int rows = static_cast<int>(nI ? ic[0].size() : (nF ? fc[0].size() : 0));
    for (int j = 0; j < nI; j++) for (int i = 0; i < rows && i < cap; i++) iout[static_cast<size_t>(j) * cap + i] = ic[j][i];
    for (int j = 0; j < nF; j++) for (int i = 0; i < rows && i < cap; i++) fout[static_cast<size_t>(j) * cap + i] = fc[j][i];
    return rows;
}
const char *sfh_last_text() { return g_text.c_str(); }

void sfh_partition_spatial(const double *weights, const double *xyz, int n, double totalWidth, int nRanks, int *owner) {
    std::vector<double> w(weights, weights + n), x(n), y(n), z(n);
    for (int i = 0; i < n; i++) { x[i] = xyz[3 * i]; y[i] = xyz[3 * i + 1]; z[i] = xyz[3 * i + 2]; }
    std::vector<int> own;
    PartitionHalosSpatial(w, x, y, z, totalWidth, nRanks, own);
    for (int i = 0; i < n; i++) owner[i] = own[i];
}

// PartitionHalos of sf_shell.cpp (LPT with affinity groups)
void sfh_partition_halos(const double *weights, const int64_t *groups, int n, int nRanks, int *owner) {
    std::vector<int> own;
    PartitionHalos(std::vector<double>(weights, weights + n), std::vector<int64_t>(groups, groups + n), nRanks, own);
    for (int i = 0; i < n; i++) owner[i] = own[i];
}

// makeTestConfig of parse/config_test.go:300-321: one variable of every kind
// under the header [config].  text == NULL skips the file step; flags is
// '\n'-separated.  Lists come back joined by ','.
const char *sfh_read_test_config(const char *text, const char *fname, const char *flags, int64_t *num, double *flt,
                                 int *okay, char *word, char *nums, char *floats, char *okays, char *words, int cap) {
    ConfigVars vars("config");
    int64_t vi; double vf; std::string vs; bool vb;
    std::vector<int64_t> vis; std::vector<double> vfs; std::vector<std::string> vss; std::vector<bool> vbs;
    vars.Int(&vi, "num", 0); vars.Ints(&vis, "nums", {}); vars.Float(&vf, "float", 0); vars.Floats(&vfs, "floats", {});
    vars.Bool(&vb, "okay", false); vars.Bools(&vbs, "okays", {}); vars.String(&vs, "word", "");
    vars.Strings(&vss, "words", {});
    std::string err = text ? vars.ReadConfigText(text, fname) : "";
    if (err.empty() && flags && *flags) {
        std::vector<std::string> args;
        std::string cur;
        for (const char *p = flags; ; p++) {
            if (*p == '\n' || !*p) { args.push_back(cur); cur.clear(); if (!*p) break; } else cur += *p;
        }
        err = vars.ReadFlags(args);
    }
    auto put = [cap](char *dst, const std::string &v) { std::strncpy(dst, v.c_str(), cap - 1); dst[cap - 1] = 0; };
    *num = vi; *flt = vf; *okay = vb;
    put(word, vs);
    std::string t;
    for (size_t k = 0; k < vis.size(); k++) t += (k ? "," : "") + std::to_string(vis[k]);
    put(nums, t); t.clear();
    for (size_t k = 0; k < vfs.size(); k++) t += (k ? "," : "") + go_g6(vfs[k]);
    put(floats, t); t.clear();
    for (size_t k = 0; k < vbs.size(); k++) t += std::string(k ? "," : "") + (vbs[k] ? "true" : "false");
    put(okays, t); t.clear();
    for (size_t k = 0; k < vss.size(); k++) t += (k ? "," : "") + vss[k];
    put(words, t);
    return keep(err);
}

const char *sfh_read_shell_config(const char *fname, const char *flags, int64_t *ints7, double *floats7, int *pp) {
    ShellConfig c;
    std::vector<std::string> args;
    if (flags && *flags) {
        std::string cur;
        for (const char *p = flags; ; p++) {
            if (*p == '\n' || !*p) { args.push_back(cur); cur.clear(); if (!*p) break; } else cur += *p;
        }
    }
    std::string err = c.ReadConfig(fname ? fname : "", args);
    ints7[0] = c.subsampleFactor; ints7[1] = c.radialBins; ints7[2] = c.spokes; ints7[3] = c.rings;
    ints7[4] = c.order; ints7[5] = c.levels; ints7[6] = c.smoothingWindow;
    floats7[0] = c.rMaxMult; floats7[1] = c.rMinMult; floats7[2] = c.rKernelMult; floats7[3] = c.eta;
    floats7[4] = c.losSlopeCutoff; floats7[5] = c.backgroundRhoMult; floats7[6] = c.percentile;
    *pp = c.percentileProfile;
    return keep(err);
}

const char *sfh_read_stats_config(const char *fname, const char *flags, int64_t *ints2, double *width, int *bools2,
                                  char *strategy, int cap) {
    StatsConfig c;
    std::vector<std::string> args;
    if (flags && *flags) {
        std::string cur;
        for (const char *p = flags; ; p++) {
            if (*p == '\n' || !*p) { args.push_back(cur); cur.clear(); if (!*p) break; } else cur += *p;
        }
    }
    std::string err = c.ReadConfig(fname ? fname : "", args);
    ints2[0] = c.monteCarloSamples; ints2[1] = c.order;
    *width = c.shellWidth;
    bools2[0] = c.skipMass; bools2[1] = c.shellFilter;
    std::strncpy(strategy, c.exclusionStrategy.c_str(), cap - 1); strategy[cap - 1] = 0;
    return keep(err);
}

const char *sfh_read_prof_config(const char *fname, const char *flags, int64_t *ints4, double *floats3, char *ptype, int cap) {
    ProfConfig c;
    std::vector<std::string> args;
    if (flags && *flags) {
        std::string cur;
        for (const char *p = flags; ; p++) {
            if (*p == '\n' || !*p) { args.push_back(cur); cur.clear(); if (!*p) break; } else cur += *p;
        }
    }
    std::string err = c.ReadConfig(fname ? fname : "", args);
    ints4[0] = c.bins; ints4[1] = c.order; ints4[2] = c.samples; ints4[3] = c.medianPixelLevel;
    floats3[0] = c.rMaxMult; floats3[1] = c.rMinMult; floats3[2] = c.percentile;
    std::strncpy(ptype, c.pType.c_str(), cap - 1); ptype[cap - 1] = 0;
    return keep(err);
}

const char *sfh_read_global_config(const char *fname, char *snapType, int cap, int64_t *blocks) {
    GlobalConfig g;
    std::string err = g.ReadConfig(fname);
    std::strncpy(snapType, g.SnapshotType.c_str(), cap - 1); snapType[cap - 1] = 0;
    if (err.empty()) {
        Catalogs cat;
        err = cat.Init(g.particle, g.SnapshotType == "gotetra");
        *blocks = cat.Blocks();
    }
    return keep(err);
}

// expands SnapshotFormat for (snap, block) -> file name
const char *sfh_catalog_name(const char *format, const char *meanings, const int64_t *mins, const int64_t *maxes,
                             int nBlockDims, int64_t snapMin, int64_t snapMax, int snap, int block, int *nBlocks) {
    ParticleInfo info;
    info.SnapshotFormat = format;
    std::string cur;
    for (const char *p = meanings; ; p++) { if (*p == ',' || !*p) { info.SnapshotFormatMeanings.push_back(trim_spaces(cur)); cur.clear(); if (!*p) break; } else cur += *p; }
    info.BlockMins.assign(mins, mins + nBlockDims); info.BlockMaxes.assign(maxes, maxes + nBlockDims);
    info.SnapMin = snapMin; info.SnapMax = snapMax;
    Catalogs cat;
    std::string err = cat.Init(info, false);
    if (!err.empty()) { *nBlocks = -1; return keep(err); }
    *nBlocks = cat.Blocks();
    return keep(cat.ParticleCatalog(snap, block));
}

// Reads one block with the named reader.  Returns N or -1; data copied to xs/ms (cap entries).
// dmTypes / singleMass: GadgetDMTypeIndices / GadgetSingleMassIndices as bit masks over the six species
// (0 = the reference's defaults, {1} and {1})
int64_t sfh_read_block(const char *type, const char *fname, const char *endianness, int64_t npartNum, double posUnits,
                       double massUnits, float *xs, float *ms, int64_t cap, void *header72, float *minMass,
                       int64_t *total, int dmTypes, int singleMass) {
    Context ctx;
    ctx.LGadgetNPartNum = npartNum; ctx.GadgetPositionUnits = posUnits; ctx.GadgetMassUnits = massUnits;
    if (dmTypes) { ctx.GadgetDMTypeIndices.clear(); for (int i = 0; i < 6; i++) if (dmTypes >> i & 1) ctx.GadgetDMTypeIndices.push_back(i); }
    if (singleMass) { ctx.GadgetDMSingleMassIndices.clear(); for (int i = 0; i < 6; i++) if (singleMass >> i & 1) ctx.GadgetDMSingleMassIndices.push_back(i); }
    std::unique_ptr<VectorBuffer> buf;
    std::string err = NewVectorBuffer(type, fname, endianness, ctx, buf);
    if (!err.empty()) { keep(err); return -1; }
    const float *px, *pm;
    int64_t n;
    if (!(err = buf->Read(fname, &px, &pm, &n)).empty()) { keep(err); return -1; }
    for (int64_t i = 0; i < n && i < cap; i++) { std::memcpy(xs + 3 * i, px + 3 * i, 12); ms[i] = pm[i]; }
    Header hd;
    if (!(err = buf->ReadHeader(fname, hd)).empty()) { keep(err); return -1; }
    std::memcpy(header72, &hd, sizeof hd);
    *minMass = buf->MinMass();
    if (!(err = buf->TotalParticles(fname, *total)).empty()) { keep(err); return -1; }
    return n;
}

// Particle IDs of a block (VectorBuffer::ReadIDs); returns the count or -1
int64_t sfh_read_ids(const char *type, const char *fname, const char *endianness, int64_t npartNum, int64_t *ids, int64_t cap,
                     int dmTypes) {
    Context ctx;
    ctx.LGadgetNPartNum = npartNum;
    if (dmTypes) { ctx.GadgetDMTypeIndices.clear(); for (int i = 0; i < 6; i++) if (dmTypes >> i & 1) ctx.GadgetDMTypeIndices.push_back(i); }
    std::unique_ptr<VectorBuffer> buf;
    std::string err = NewVectorBuffer(type, fname, endianness, ctx, buf);
    if (!err.empty()) { keep(err); return -1; }
    const int64_t *p;
    int64_t n;
    if (!(err = buf->ReadIDs(fname, &p, &n)).empty()) { keep(err); return -1; }
    for (int64_t i = 0; i < n && i < cap; i++) ids[i] = p[i];
    return n;
}

// io.WriteFilter; particles are concatenated, lens[i] per halo.  Returns 0 or -1
int sfh_write_filter(const char *path, const char *endianness, int n, const int64_t *snaps, const int64_t *ids,
                     const int64_t *lens, const int64_t *particles) {
    std::vector<std::vector<int64_t>> ps(n);
    int64_t at = 0;
    for (int i = 0; i < n; i++) { ps[i].assign(particles + at, particles + at + lens[i]); at += lens[i]; }
    std::string err = WriteFilter(path, endianness, std::vector<int64_t>(snaps, snaps + n), std::vector<int64_t>(ids, ids + n), ps);
    if (!err.empty()) { keep(err); return -1; }
    return 0;
}

void sfh_bounding_box(const float *xs, int64_t n, double tw, float *origin, float *width) { boundingBox(xs, n, tw, origin, width); }

}  // extern "C"
