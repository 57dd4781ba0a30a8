// sf_snapshot.cpp -- see sf_snapshot.hpp.
#include "sf_snapshot.hpp"

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sys/stat.h>

#include "sf_config.hpp"

namespace sfk { double rho_average(double H0, double omegaM, double omegaL, double z); }

namespace sfhost {

namespace {

bool read_file(const std::string &path, std::vector<unsigned char> &out, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "open " + path + ": " + std::strerror(errno); return false; }
    std::fseek(f, 0, SEEK_END);
    long sz = std::ftell(f);
    std::fseek(f, 0, SEEK_SET);
    out.resize(sz > 0 ? sz : 0);
    size_t got = sz > 0 ? std::fread(out.data(), 1, sz, f) : 0;
    std::fclose(f);
    if (static_cast<long>(got) != sz) { err = "read " + path + ": short read"; return false; }
    return true;
}

// The first `want` bytes of a file (fewer if it is shorter): headers are parsed from these.
bool read_prefix(const std::string &path, size_t want, std::vector<unsigned char> &out, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "open " + path + ": " + std::strerror(errno); return false; }
    out.resize(want);
    out.resize(std::fread(out.data(), 1, want, f));
    std::fclose(f);
    return true;
}

// An open snapshot file read in pieces: only the blocks the shell path needs (positions,
// masses) are transferred, the velocity and id blocks are skipped with a seek.
struct FileReader {
    FILE *f = nullptr;
    int64_t size = 0;
    ~FileReader() { if (f) std::fclose(f); }
    bool open(const std::string &path, std::string &err) {
        f = std::fopen(path.c_str(), "rb");
        if (!f) { err = "open " + path + ": " + std::strerror(errno); return false; }
        std::fseek(f, 0, SEEK_END);
        size = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        return true;
    }
    bool read_at(int64_t offset, void *dst, size_t bytes) {
        if (offset < 0 || offset + static_cast<int64_t>(bytes) > size) return false;
        if (std::fseek(f, static_cast<long>(offset), SEEK_SET) != 0) return false;
        return bytes == 0 || std::fread(dst, 1, bytes, f) == bytes;
    }
};

template <typename T> void swap_array(T *dst, size_t n) {
    for (size_t k = 0; k < n; k++) {
        unsigned char *b = reinterpret_cast<unsigned char *>(&dst[k]);
        for (size_t i = 0; i < sizeof(T) / 2; i++) std::swap(b[i], b[sizeof(T) - 1 - i]);
    }
}

// A cursor over a file image; `swap` = the file's byte order differs from the host's.
struct Cursor {
    const unsigned char *p, *end;
    bool swap;
    bool ok = true;
    template <typename T> T get(bool forceLittle = false) {
        T v{};
        if (p + sizeof(T) > end) { ok = false; return v; }
        std::memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        if (swap && !forceLittle) {
            unsigned char *b = reinterpret_cast<unsigned char *>(&v);
            for (size_t i = 0; i < sizeof(T) / 2; i++) std::swap(b[i], b[sizeof(T) - 1 - i]);
        }
        return v;
    }
    template <typename T> void array(T *dst, size_t n) {
        if (p + n * sizeof(T) > end) { ok = false; return; }
        std::memcpy(dst, p, n * sizeof(T));
        p += n * sizeof(T);
        if (swap)
            for (size_t k = 0; k < n; k++) {
                unsigned char *b = reinterpret_cast<unsigned char *>(&dst[k]);
                for (size_t i = 0; i < sizeof(T) / 2; i++) std::swap(b[i], b[sizeof(T) - 1 - i]);
            }
    }
};

bool needs_swap(const std::string &flag) {   // io/lgadget2.go:194-203 (host is little endian)
    return flag == "BigEndian";
}

// ---- (L)Gadget-2 headers, io/lgadget2.go:12-23 and io/gadget2.go:11-25 (256 bytes each)
// Fortran framing: [i32 256][header][i32 256][i32 bytes of the position block] -> positions start here
constexpr int64_t kGadgetPosOffset = 4 + 256 + 4 + 4;
struct LGadget2Header {
    uint32_t NPart[6]; double Mass[6]; double Time, Redshift; int32_t FlagSfr, FlagFeedback;
    uint32_t NPartTotal[6]; int32_t FlagCooling, NumFiles; double BoxSize, Omega0, OmegaLambda, HubbleParam;
    int32_t FlagStellarAge, HashTabSize;
};
struct Gadget2Header {
    uint32_t NPart[6]; double Mass[6]; double Time, Redshift; int32_t FlagSfr, FlagFeedback;
    uint32_t NumPartTotal[6]; int32_t FlagCooling, NumFiles; double BoxSize, Omega0, OmegaLambda, HubbleParam;
    int32_t FlagStellarAge, FlagMetals; uint32_t NumPartTotalHW[6]; int32_t FlagEntropyICs;
};

void parse_common(Cursor &c, bool little, uint32_t *npart, double *mass, double &time, double &z, uint32_t *tot,
                  int32_t &numFiles, double &box, double &om, double &ol, double &h) {
    for (int i = 0; i < 6; i++) npart[i] = c.get<uint32_t>(little);
    for (int i = 0; i < 6; i++) mass[i] = c.get<double>(little);
    time = c.get<double>(little); z = c.get<double>(little);
    c.get<int32_t>(little); c.get<int32_t>(little);
    for (int i = 0; i < 6; i++) tot[i] = c.get<uint32_t>(little);
    c.get<int32_t>(little); numFiles = c.get<int32_t>(little);
    box = c.get<double>(little); om = c.get<double>(little); ol = c.get<double>(little); h = c.get<double>(little);
}

LGadget2Header parse_lgadget(Cursor &c, bool little) {
    LGadget2Header h{};
    const unsigned char *start = c.p;
    parse_common(c, little, h.NPart, h.Mass, h.Time, h.Redshift, h.NPartTotal, h.NumFiles, h.BoxSize, h.Omega0,
                 h.OmegaLambda, h.HubbleParam);
    h.FlagStellarAge = c.get<int32_t>(little); h.HashTabSize = c.get<int32_t>(little);
    c.p = start + 256;
    if (c.p > c.end) c.ok = false;
    return h;
}

Gadget2Header parse_gadget(Cursor &c, bool little) {
    Gadget2Header h{};
    const unsigned char *start = c.p;
    parse_common(c, little, h.NPart, h.Mass, h.Time, h.Redshift, h.NumPartTotal, h.NumFiles, h.BoxSize, h.Omega0,
                 h.OmegaLambda, h.HubbleParam);
    h.FlagStellarAge = c.get<int32_t>(little); h.FlagMetals = c.get<int32_t>(little);
    for (int i = 0; i < 6; i++) h.NumPartTotalHW[i] = c.get<uint32_t>(little);
    h.FlagEntropyICs = c.get<int32_t>(little);
    c.p = start + 256;
    if (c.p > c.end) c.ok = false;
    return h;
}

std::string corrupt(const std::string &path) {
    return "Corruption detected in the file " + path + ". I can't analyze it.";
}

// io/lgadget2.go:123-143 and io/gadget2.go:297-320: wrap into [0, L), reject garbage
bool fix_positions(float *xs, int64_t n, float tw) {
    for (int64_t i = 0; i < 3 * n; i++) {
        if (xs[i] < 0) xs[i] += tw; else if (xs[i] >= tw) xs[i] -= tw;
        if (std::isnan(xs[i]) || std::isinf(xs[i]) || xs[i] < -tw || xs[i] > 2 * tw) return false;
    }
    return true;
}

// ------------------------------------------------------------ LGadget-2
class LGadget2Buffer : public VectorBuffer {
public:
    LGadget2Buffer(bool swap, Context ctx) : swap_(swap), ctx_(std::move(ctx)) {}
    std::string init(const std::string &path) {   // NewLGadget2Buffer, io/lgadget2.go:188-223
        LGadget2Header hd{};
        std::string err = header(path, hd);
        if (!err.empty()) return err;
        int64_t tot = 0;
        err = TotalParticles(path, tot);
        if (!err.empty()) return err;
        mass_ = calcUniformMass(tot, hd.BoxSize, Cosmology{hd.Redshift, hd.Omega0, hd.OmegaLambda, hd.HubbleParam});
        return "";
    }
    std::string Read(const std::string &fname, const float **xs, const float **ms, int64_t *n) override {
        // file = [i32 hdr(256) i32] [i32 pos(12 N) i32] [i32 vel(12 N) i32] [i32 ids(8 N) i32]: only the
        // header and the position block are transferred (readLGadget2Particles, io/lgadget2.go:84-150)
        FileReader fr;
        std::string err;
        if (!fr.open(fname, err)) return err;
        unsigned char head[kGadgetPosOffset];
        if (!fr.read_at(0, head, sizeof head)) return "read " + fname + ": unexpected EOF";
        Cursor c{head, head + sizeof head, swap_};
        c.get<int32_t>();
        LGadget2Header gh = parse_lgadget(c, true);   // the header is always little endian, :100
        int64_t count = 0;
        err = particle_num(gh.NPart, count);
        if (!err.empty()) return err;
        // a header that promises more particles than the file holds must not size the buffers
        if (count < 0 || kGadgetPosOffset + 12 * count > fr.size) return "read " + fname + ": unexpected EOF";
        xs_.resize(3 * count);
        if (!fr.read_at(kGadgetPosOffset, xs_.data(), 12 * static_cast<size_t>(count))) return "read " + fname + ": unexpected EOF";
        if (swap_) swap_array(xs_.data(), xs_.size());
        if (!fix_positions(xs_.data(), count, static_cast<float>(gh.BoxSize))) return corrupt(fname);
        ms_.assign(count, mass_);
        *xs = xs_.data(); *ms = ms_.data(); *n = count;
        return "";
    }
    std::string ReadIDs(const std::string &fname, const int64_t **ids, int64_t *n) override {
        // [hdr] [pos 12 N] [vel 12 N] [ids 8 N], io/lgadget2.go:107-116
        FileReader fr;
        std::string err;
        if (!fr.open(fname, err)) return err;
        unsigned char head[kGadgetPosOffset];
        if (!fr.read_at(0, head, sizeof head)) return "read " + fname + ": unexpected EOF";
        Cursor c{head, head + sizeof head, swap_};
        c.get<int32_t>();
        LGadget2Header gh = parse_lgadget(c, true);
        int64_t count = 0;
        err = particle_num(gh.NPart, count);
        if (!err.empty()) return err;
        const int64_t at = kGadgetPosOffset + 12 * count + 4 + (4 + 12 * count + 4) + 4;
        if (count < 0 || at + 8 * count > fr.size) return "read " + fname + ": unexpected EOF";
        ids_.resize(count);
        if (!fr.read_at(at, ids_.data(), 8 * static_cast<size_t>(count))) return "read " + fname + ": unexpected EOF";
        if (swap_) swap_array(ids_.data(), ids_.size());
        *ids = ids_.data(); *n = count;
        return "";
    }
    std::string ReadHeader(const std::string &fname, Header &out) override {   // :259-273 + postprocess :25-40
        LGadget2Header hd{};
        std::string err = header(fname, hd);
        if (!err.empty()) return err;
        const float *xs, *ms;
        int64_t n;
        err = Read(fname, &xs, &ms, &n);
        if (!err.empty()) return err;
        out.TotalWidth = hd.BoxSize;
        err = particle_num(hd.NPart, out.N);
        if (!err.empty()) return err;
        out.Cosmo = Cosmology{hd.Redshift, hd.Omega0, hd.OmegaLambda, hd.HubbleParam};
        boundingBox(xs, n, hd.BoxSize, out.Origin, out.Width);
        return "";
    }
    float MinMass() const override { return mass_; }
    bool UniformMass() const override { return true; }
    std::string TotalParticles(const std::string &fname, int64_t &n) override {
        LGadget2Header hd{};
        std::string err = header(fname, hd);
        if (!err.empty()) return err;
        return particle_num(hd.NPartTotal, n);
    }
private:
    std::string header(const std::string &path, LGadget2Header &hd) {   // readLGadget2Header, :70-82
        std::vector<unsigned char> img;
        std::string err;
        if (!read_prefix(path, kGadgetPosOffset, img, err)) return err;
        Cursor c{img.data(), img.data() + img.size(), swap_};
        c.get<int32_t>();
        hd = parse_lgadget(c, true);
        return c.ok ? "" : "unexpected EOF";
    }
    std::string particle_num(const uint32_t npart[6], int64_t &n) const {   // lgadgetParticleNum, :42-57
        if (ctx_.LGadgetNPartNum == 2) {
            if (npart[0] > 100 * 1000)
                return "Simulation contains too many particles. This is probably because GadgetNpartNum is set to "
                       "2 when it should be set to 1.";
            n = static_cast<int64_t>(npart[1]) + (static_cast<int64_t>(npart[0]) << 32);
        } else {
            n = npart[0];
        }
        return "";
    }
    bool swap_;
    Context ctx_;
    float mass_ = 0;
    std::vector<float> xs_, ms_;
    std::vector<int64_t> ids_;
};

// -------------------------------------------------------------- Gadget-2
class Gadget2Buffer : public VectorBuffer {
public:
    Gadget2Buffer(bool swap, Context ctx) : swap_(swap), ctx_(std::move(ctx)) {}
    std::string init(const std::string &path) { Gadget2Header hd; return header(path, hd); }
    std::string Read(const std::string &fname, const float **xs, const float **ms, int64_t *n) override {
        std::string err;
        // file = [i32 hdr i32] [i32 pos i32] [i32 vel i32] [i32 ids i32] [i32 masses i32]: header, positions and
        // the mass block are transferred, velocities and ids are skipped (readGadget2Particles, io/gadget2.go:74-130)
        FileReader fr;
        if (!fr.open(fname, err)) return err;
        unsigned char head[kGadgetPosOffset];
        if (!fr.read_at(0, head, sizeof head)) return "read " + fname + ": unexpected EOF";
        Cursor c{head, head + sizeof head, swap_};
        c.get<int32_t>();
        Gadget2Header gh = parse_gadget(c, false);   // honours Endianness here, io/gadget2.go:83
        int64_t totalN = 0, dmN = 0, multiN = 0;
        for (int i = 0; i < 6; i++) {
            totalN += gh.NPart[i];
            if (isDM(i)) dmN += gh.NPart[i];
            if (isMulti(i)) multiN += gh.NPart[i];
        }
        // a header that promises more particles than the file holds must not size the buffers
        if (totalN < 0 || kGadgetPosOffset + 24 * totalN > fr.size) return "read " + fname + ": unexpected EOF";
        xs_.resize(3 * totalN);
        ms_.assign(totalN, 0.f);
        std::vector<float> multi(multiN);
        int64_t at = kGadgetPosOffset;
        if (!fr.read_at(at, xs_.data(), 12 * static_cast<size_t>(totalN))) return "read " + fname + ": unexpected EOF";
        if (swap_) swap_array(xs_.data(), xs_.size());
        at += 12 * totalN + 4;                 // positions + their end marker
        at += 4 + 12 * totalN + 4;             // velocities
        int32_t idBytes = 0;                   // ids: 8 or 4 bytes each, :110-120
        if (!fr.read_at(at, &idBytes, 4)) return "read " + fname + ": unexpected EOF";
        if (swap_) swap_array(&idBytes, 1);
        at += 4;
        if (totalN > 0) at += (idBytes / totalN == 8 ? 8 : (idBytes / totalN == 4 ? 4 : 0)) * totalN;
        at += 4;
        at += 4;                               // start marker of the mass block, read unconditionally, :123-125
        if (!fr.read_at(at, multi.data(), 4 * static_cast<size_t>(multiN))) return "read " + fname + ": unexpected EOF";
        if (swap_) swap_array(multi.data(), multi.size());
        if (at + 4 * multiN + 4 > fr.size) return "read " + fname + ": unexpected EOF";   // its end marker
        // unpackMass :202-233, then packVec / packFloat32 :235-295 keep the DM species
        int64_t off = 0, moff = 0, dst = 0;
        for (int i = 0; i < 6; i++) {
            const int64_t cnt = gh.NPart[i];
            for (int64_t k = 0; k < cnt; k++)
                ms_[off + k] = isMulti(i) ? multi[moff + k] : static_cast<float>(gh.Mass[i]);
            if (isMulti(i)) moff += cnt;
            if (isDM(i)) {
                std::memmove(&xs_[3 * dst], &xs_[3 * off], sizeof(float) * 3 * cnt);
                std::memmove(&ms_[dst], &ms_[off], sizeof(float) * cnt);
                dst += cnt;
            }
            off += cnt;
        }
        xs_.resize(3 * dmN);
        ms_.resize(dmN);
        // fix, :297-322
        if (!fix_positions(xs_.data(), dmN, static_cast<float>(gh.BoxSize))) return corrupt(fname);
        const float pu = static_cast<float>(ctx_.GadgetPositionUnits), mu = static_cast<float>(ctx_.GadgetMassUnits);
        for (float &v : xs_) v *= pu;
        for (float &v : ms_) v *= mu;
        *xs = xs_.data(); *ms = ms_.data(); *n = dmN;
        return "";
    }
    std::string ReadIDs(const std::string &fname, const int64_t **ids, int64_t *n) override {
        // ids of every species, 8 or 4 bytes each (io/gadget2.go:109-120), then packInt64 keeps the DM species (:133)
        FileReader fr;
        std::string err;
        if (!fr.open(fname, err)) return err;
        unsigned char head[kGadgetPosOffset];
        if (!fr.read_at(0, head, sizeof head)) return "read " + fname + ": unexpected EOF";
        Cursor c{head, head + sizeof head, swap_};
        c.get<int32_t>();
        Gadget2Header gh = parse_gadget(c, false);
        int64_t totalN = 0;
        for (int i = 0; i < 6; i++) totalN += gh.NPart[i];
        int64_t at = kGadgetPosOffset + 12 * totalN + 4 + (4 + 12 * totalN + 4);
        int32_t idBytes = 0;
        if (totalN < 0 || !fr.read_at(at, &idBytes, 4)) return "read " + fname + ": unexpected EOF";
        if (swap_) swap_array(&idBytes, 1);
        at += 4;
        std::vector<int64_t> all(totalN);
        const int64_t each = totalN > 0 ? idBytes / totalN : 8;
        if (each == 8) {
            if (!fr.read_at(at, all.data(), 8 * static_cast<size_t>(totalN))) return "read " + fname + ": unexpected EOF";
            if (swap_) swap_array(all.data(), all.size());
        } else if (each == 4) {
            std::vector<int32_t> i32(totalN);
            if (!fr.read_at(at, i32.data(), 4 * static_cast<size_t>(totalN))) return "read " + fname + ": unexpected EOF";
            if (swap_) swap_array(i32.data(), i32.size());
            for (int64_t k = 0; k < totalN; k++) all[k] = i32[k];
        } else {
            return "Cannot read particle IDs of " + fname + ": " + std::to_string(each) + " bytes per ID.";
        }
        ids_.clear();
        int64_t off = 0;
        for (int i = 0; i < 6; i++) {
            if (isDM(i)) ids_.insert(ids_.end(), all.begin() + off, all.begin() + off + gh.NPart[i]);
            off += gh.NPart[i];
        }
        *ids = ids_.data(); *n = static_cast<int64_t>(ids_.size());
        return "";
    }
    std::string ReadHeader(const std::string &fname, Header &out) override {   // :386-400 + postprocess :29-48
        Gadget2Header hd{};
        std::string err = header(fname, hd);
        if (!err.empty()) return err;
        const float *xs, *ms;
        int64_t n;
        err = Read(fname, &xs, &ms, &n);
        if (!err.empty()) return err;
        out.TotalWidth = hd.BoxSize * ctx_.GadgetPositionUnits;
        out.N = 0;
        for (int64_t i : ctx_.GadgetDMTypeIndices) out.N += hd.NPart[i];
        out.Cosmo = Cosmology{hd.Redshift, hd.Omega0, hd.OmegaLambda, hd.HubbleParam};
        boundingBox(xs, n, out.TotalWidth, out.Origin, out.Width);
        return "";
    }
    float MinMass() const override { return 0.f; }   // never assigned in the reference, io/gadget2.go:327,402
    std::string TotalParticles(const std::string &fname, int64_t &n) override {
        Gadget2Header hd{};
        std::string err = header(fname, hd);
        if (!err.empty()) return err;
        n = 0;
        for (int64_t i : ctx_.GadgetDMTypeIndices)
            n += static_cast<int64_t>(hd.NumPartTotal[i]) + (static_cast<int64_t>(hd.NumPartTotalHW[i]) << 32);
        return "";
    }
private:
    bool isDM(int i) const { for (int64_t j : ctx_.GadgetDMTypeIndices) if (j == i) return true; return false; }
    bool isMulti(int i) const { for (int64_t j : ctx_.GadgetDMSingleMassIndices) if (j == i) return false; return true; }
    std::string header(const std::string &path, Gadget2Header &hd) {   // readGadget2Header: always little endian, :60
        std::vector<unsigned char> img;
        std::string err;
        if (!read_prefix(path, kGadgetPosOffset, img, err)) return err;
        Cursor c{img.data(), img.data() + img.size(), swap_};
        c.get<int32_t>();
        hd = parse_gadget(c, true);
        return c.ok ? "" : "unexpected EOF";
    }
    bool swap_;
    Context ctx_;
    std::vector<float> xs_, ms_;
    std::vector<int64_t> ids_;
};

// --------------------------------------------------------------- gotetra
struct RawGotetraHeader {   // io/gotetra.go:109-120, 152 bytes
    Cosmology Cosmo;
    int64_t Count, CountWidth, SegmentWidth, GridWidth, GridCount, Idx, Cells;
    double Mass, TotalWidth;
    float Origin[3], Width[3], VelocityOrigin[3], VelocityWidth[3];
};
static_assert(sizeof(RawGotetraHeader) == 152, "rawGotetraHeader layout");

class GotetraBuffer : public VectorBuffer {
public:
    std::string init(const std::string &fname) {   // NewGotetraBuffer, io/gotetra.go:24-47
        RawGotetraHeader hd;
        std::vector<unsigned char> img;
        std::string err = load(fname, hd, img, nullptr);
        if (!err.empty()) return err;
        // the reference indexes sheet[x + y*GridWidth + z*GridWidth^2] for x, y, z < SegmentWidth
        // (io/gotetra.go:73-85) and panics on a header that does not fit; refuse it here
        if (!(hd.SegmentWidth > 0 && hd.SegmentWidth <= hd.GridWidth && hd.GridWidth <= 4096 &&
              hd.GridWidth * hd.GridWidth * hd.GridWidth == hd.GridCount)) {
            char buf[256];
            snprintf(buf, sizeof buf, "Corruption detected in %s: SegmentWidth = %lld, GridWidth = %lld, GridCount = %lld.",
                     fname.c_str(), static_cast<long long>(hd.SegmentWidth), static_cast<long long>(hd.GridWidth),
                     static_cast<long long>(hd.GridCount));
            return buf;
        }
        sw_ = hd.SegmentWidth; gw_ = hd.GridWidth;
        mass_ = calcUniformMass(hd.Count, hd.TotalWidth, hd.Cosmo);
        return "";
    }
    std::string Read(const std::string &fname, const float **xs, const float **ms, int64_t *n) override {
        RawGotetraHeader hd;
        std::vector<unsigned char> img;
        Cursor c{nullptr, nullptr, false};
        std::string err = load(fname, hd, img, &c);
        if (!err.empty()) return err;
        if (hd.GridCount != gw_ * gw_ * gw_) {
            char buf[256];
            snprintf(buf, sizeof buf, "Position buffer has length %lld, but file %s has %lld vectors.",
                     static_cast<long long>(gw_ * gw_ * gw_), fname.c_str(), static_cast<long long>(hd.GridCount));
            return buf;
        }
        if (hd.GridCount < 0 || 12 * hd.GridCount > static_cast<int64_t>(img.size())) return "read " + fname + ": unexpected EOF";
        if (c.p + 12 * hd.GridCount > c.end) return "read " + fname + ": unexpected EOF";
        xs_.resize(3 * sw_ * sw_ * sw_);
        if (static_cast<int64_t>(ms_.size()) != sw_ * sw_ * sw_) ms_.assign(sw_ * sw_ * sw_, mass_);
        // the SegmentWidth^3 corner of the GridWidth^3 sheet (:73-85), one x-row at a time
        // straight out of the file image
        const unsigned char *sheet = c.p;
        for (int64_t z = 0; z < sw_; z++)
            for (int64_t y = 0; y < sw_; y++)
                std::memcpy(&xs_[3 * (y + z * sw_) * sw_], sheet + 12 * (y * gw_ + z * gw_ * gw_), 12 * sw_);
        if (c.swap)
            for (float &v : xs_) {
                unsigned char *b = reinterpret_cast<unsigned char *>(&v);
                std::swap(b[0], b[3]); std::swap(b[1], b[2]);
            }
        *xs = xs_.data(); *ms = ms_.data(); *n = sw_ * sw_ * sw_;
        return "";
    }
    std::string ReadIDs(const std::string &, const int64_t **ids, int64_t *n) override {
        // the gotetra reader allocates its ID slice and never fills it (io/gotetra.go:40,87): zeros
        ids_.assign(sw_ * sw_ * sw_, 0);
        *ids = ids_.data(); *n = sw_ * sw_ * sw_;
        return "";
    }
    std::string ReadHeader(const std::string &fname, Header &out) override {   // :213-226 + postprocess :122-128
        RawGotetraHeader hd;
        std::vector<unsigned char> img;
        std::string err = load(fname, hd, img, nullptr);
        if (!err.empty()) return err;
        out.Cosmo = hd.Cosmo;
        out.N = hd.SegmentWidth * hd.SegmentWidth * hd.SegmentWidth;
        out.TotalWidth = hd.TotalWidth;
        std::memcpy(out.Origin, hd.Origin, sizeof out.Origin);
        std::memcpy(out.Width, hd.Width, sizeof out.Width);
        return "";
    }
    float MinMass() const override { return mass_; }
    bool UniformMass() const override { return true; }
    std::string TotalParticles(const std::string &, int64_t &n) override { n = -1; return ""; }
private:
    // loadSheetHeader, io/gotetra.go:177-211
    std::string load(const std::string &fname, RawGotetraHeader &hd, std::vector<unsigned char> &img, Cursor *rest) {
        std::string err;
        // without `rest` only the header is wanted: 8 bytes of flags + the 152-byte struct
        if (!(rest ? read_file(fname, img, err) : read_prefix(fname, 8 + sizeof(RawGotetraHeader), img, err))) return err;
        Cursor c{img.data(), img.data() + img.size(), false};
        const int32_t flag = c.get<int32_t>();
        if (flag == -1) c.swap = true;          // endianness(): 0 little, -1 big, :145-153
        else if (flag != 0) return "Unrecognized endianness flag.";
        const int32_t size = c.get<int32_t>();
        if (size != static_cast<int32_t>(sizeof(RawGotetraHeader))) {
            char buf[128];
            snprintf(buf, sizeof buf, "Expected catalog.SheetHeader size of %zu, found %d.", sizeof(RawGotetraHeader), size);
            return buf;
        }
        hd.Cosmo.Z = c.get<double>(); hd.Cosmo.OmegaM = c.get<double>();
        hd.Cosmo.OmegaL = c.get<double>(); hd.Cosmo.H100 = c.get<double>();
        hd.Count = c.get<int64_t>(); hd.CountWidth = c.get<int64_t>(); hd.SegmentWidth = c.get<int64_t>();
        hd.GridWidth = c.get<int64_t>(); hd.GridCount = c.get<int64_t>(); hd.Idx = c.get<int64_t>();
        hd.Cells = c.get<int64_t>();
        hd.Mass = c.get<double>(); hd.TotalWidth = c.get<double>();
        for (float *arr : {hd.Origin, hd.Width, hd.VelocityOrigin, hd.VelocityWidth})
            for (int i = 0; i < 3; i++) arr[i] = c.get<float>();
        if (!c.ok) return "unexpected EOF";
        hd.Count = hd.CountWidth * hd.CountWidth * hd.CountWidth;   // "a bug in the current gotetra version", :206-208
        if (rest) *rest = c;
        return "";
    }
    int64_t sw_ = 0, gw_ = 0;
    float mass_ = 0;
    std::vector<float> xs_, ms_;
    std::vector<int64_t> ids_;
};

}  // namespace

std::string WriteFilter(const std::string &path, const std::string &endianness, const std::vector<int64_t> &snaps,
                        const std::vector<int64_t> &ids, const std::vector<std::vector<int64_t>> &particles) {
    const bool big = needs_swap(endianness);   // flagToOrder, io/filter_data.go:20-34 (the host is little endian)
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) return "open " + path + ": " + std::strerror(errno);
    auto put32 = [&](int32_t v) { if (big) swap_array(&v, 1); std::fwrite(&v, 4, 1, f); };
    auto put64 = [&](int64_t v) { if (big) swap_array(&v, 1); std::fwrite(&v, 8, 1, f); };
    put32(big ? -1 : 0);
    put32(static_cast<int32_t>(snaps.size()));
    const int64_t baseOffset = 4 + 4 + static_cast<int64_t>(snaps.size()) * 32;
    for (size_t i = 0; i < snaps.size(); i++) {   // haloInfo{ID, Snap, StartByte, Len}
        put64(ids[i]); put64(snaps[i]);
        put64(baseOffset);                        // totLen stays 0 in the reference, :52
        put64(static_cast<int64_t>(particles[i].size()));
    }
    for (const std::vector<int64_t> &p : particles) {
        if (!big) { if (!p.empty()) std::fwrite(p.data(), 8, p.size(), f); }
        else for (int64_t v : p) put64(v);
    }
    const bool ok = std::ferror(f) == 0;
    std::fclose(f);
    return ok ? "" : "write " + path + ": " + std::strerror(errno);
}

float calcUniformMass(int64_t count, double tw, const Cosmology &c) {
    const double rhoM0 = sfk::rho_average(c.H100 * 100, c.OmegaM, c.OmegaL, 0);
    const double mTot = (tw * tw * tw) * rhoM0;
    return static_cast<float>(mTot / static_cast<double>(count));
}

void boundingBox(const float *xs, int64_t n, double totalWidth, float origin[3], float width[3]) {
    for (int j = 0; j < 3; j++) { origin[j] = n > 0 ? xs[j] : 0.f; width[j] = 0.f; }
    const float tw = static_cast<float>(totalWidth), tw2 = static_cast<float>(totalWidth) / 2;
    for (int64_t i = 0; i < n; i++)
        for (int j = 0; j < 3; j++) {
            float x = xs[3 * i + j];
            const float x0 = origin[j], w = width[j];
            if (x > x0 && x < x0 + w) continue;
            if (x - x0 > tw2) x -= tw; else if (x - x0 < -tw2) x += tw;
            if (x < x0) { width[j] += x0 - x; origin[j] = x; }
            else if (x - x0 > w) width[j] = x - x0;
        }
    for (int j = 0; j < 3; j++)
        if (width[j] > tw2) { width[j] = tw; origin[j] = 0; }
}

std::string NewVectorBuffer(const std::string &snapshotType, const std::string &fname,
                            const std::string &endianness, const Context &ctx,
                            std::unique_ptr<VectorBuffer> &out) {
    if (snapshotType == "LGadget-2") {
        auto b = std::make_unique<LGadget2Buffer>(needs_swap(endianness), ctx);
        std::string err = b->init(fname);
        if (!err.empty()) return err;
        out = std::move(b);
        return "";
    }
    if (snapshotType == "Gadget-2") {
        auto b = std::make_unique<Gadget2Buffer>(needs_swap(endianness), ctx);
        std::string err = b->init(fname);
        if (!err.empty()) return err;
        out = std::move(b);
        return "";
    }
    if (snapshotType == "gotetra") {
        auto b = std::make_unique<GotetraBuffer>();
        std::string err = b->init(fname);
        if (!err.empty()) return err;
        out = std::move(b);
        return "";
    }
    if (snapshotType == "ARTIO" || snapshotType == "Bolshoi" || snapshotType == "BolshoiP" || snapshotType == "nil")
        return "SnapshotType '" + snapshotType + "' is not built in the B200 host (only LGadget-2, Gadget-2, gotetra).";
    return "SnapshotType '" + snapshotType + "' not recognized.";
}

// ------------------------------------------------------------------ env

std::string go_sprintf(const std::string &format, const std::vector<FormatArg> &args) {
    std::string out;
    size_t ai = 0;
    for (size_t i = 0; i < format.size(); i++) {
        if (format[i] != '%') { out += format[i]; continue; }
        size_t j = i + 1;
        if (j < format.size() && format[j] == '%') { out += '%'; i = j; continue; }
        std::string spec = "%";
        while (j < format.size() && (format[j] == '-' || format[j] == '+' || format[j] == ' ' || format[j] == '0' ||
                                     (format[j] >= '1' && format[j] <= '9') || format[j] == '.'))
            spec += format[j++];
        if (j >= format.size()) { out += "%!(NOVERB)"; break; }
        const char verb = format[j];
        char buf[512];
        if (ai >= args.size()) {
            out += std::string("%!") + verb + "(MISSING)";
        } else if ((verb == 'd') && args[ai].isInt) {
            snprintf(buf, sizeof buf, (spec + "lld").c_str(), static_cast<long long>(args[ai].i));
            out += buf; ai++;
        } else if ((verb == 's' || verb == 'v') && !args[ai].isInt) {
            snprintf(buf, sizeof buf, (spec + "s").c_str(), args[ai].s.c_str());
            out += buf; ai++;
        } else if (verb == 'v' && args[ai].isInt) {
            out += std::to_string(args[ai].i); ai++;
        } else if (verb == 's' && args[ai].isInt) {
            out += "%!s(int=" + std::to_string(args[ai].i) + ")"; ai++;
        } else if (verb == 'd' && !args[ai].isInt) {
            out += "%!d(string=" + args[ai].s + ")"; ai++;
        } else {
            out += std::string("%!") + verb + "(?)"; ai++;
        }
        i = j;
    }
    return out;
}

std::string Catalogs::Init(const ParticleInfo &info, bool gotetra) {
    if (gotetra && info.BlockMins.size() != 3)
        return "'BlockMins' had " + std::to_string(info.BlockMins.size()) +
               " elements, but 3 are required for gotetra catalogs.";
    snapMin_ = static_cast<int>(info.SnapMin);
    // GetColumn, cmd/env/env.go:61-118
    std::vector<std::vector<FormatArg>> cols;
    std::vector<bool> aligned;
    for (const std::string &m : info.SnapshotFormatMeanings) {
        std::vector<FormatArg> col;
        if (m == "ScaleFactor") {
            std::ifstream f(info.ScaleFactorFile);
            if (!f) return "open " + info.ScaleFactorFile + ": " + std::strerror(errno);
            std::string line;
            while (std::getline(f, line)) {
                line = trim_spaces(line);
                if (!line.empty()) col.push_back({false, 0, line});
            }
            if (static_cast<int64_t>(col.size()) != info.SnapMax - info.SnapMin + 1) {
                char buf[512];
                snprintf(buf, sizeof buf, "%s has %zu non-empty lines, but SnapMax = %lld and SnapMin = %lld.",
                         info.ScaleFactorFile.c_str(), col.size(), static_cast<long long>(info.SnapMax),
                         static_cast<long long>(info.SnapMin));
                return buf;
            }
            aligned.push_back(true);
        } else if (m == "Snapshot") {
            for (int64_t s = info.SnapMin; s <= info.SnapMax; s++) col.push_back({true, s, ""});
            aligned.push_back(true);
        } else if (m.size() >= 5 && m.compare(0, 5, "Block") == 0) {
            int64_t idx = 0;
            if (m.size() > 5 && !go_atoi(m.substr(5), idx)) return "strconv.Atoi: parsing \"" + m.substr(5) + "\": invalid syntax";
            if (idx < 0 || static_cast<size_t>(idx) >= info.BlockMins.size() || static_cast<size_t>(idx) >= info.BlockMaxes.size())
                return "SnapshotFormatMeanings refers to Block" + std::to_string(idx) + ", but BlockMins/BlockMaxes are shorter.";
            for (int64_t b = info.BlockMins[idx]; b <= info.BlockMaxes[idx]; b++) col.push_back({true, b, ""});
            aligned.push_back(false);
        } else {
            return "I don't recognize the SnapshotFormatMeanings entry '" + m + "'.";
        }
        cols.push_back(col);
    }
    // interleave, cmd/env/env.go:187-234
    int64_t snaps = -1, nonSnaps = 1;
    for (size_t i = 0; i < cols.size(); i++) {
        if (aligned[i]) {
            if (snaps == -1) snaps = cols[i].size();
            else if (snaps != static_cast<int64_t>(cols[i].size())) return "Snapshot-aligned columns of unequal height.";
        } else {
            nonSnaps *= cols[i].size();
        }
    }
    if (snaps == -1) snaps = 1;
    names_.clear();
    for (int64_t snap = 0; snap < snaps; snap++) {
        std::vector<std::string> names;
        for (int64_t row = 0; row < nonSnaps; row++) {
            std::vector<FormatArg> args(cols.size());
            int64_t divSize = 1, modSize = 1;
            for (size_t col = 0; col < cols.size(); col++) {
                if (aligned[col]) {
                    args[col] = cols[col][snap];
                } else {
                    divSize *= modSize;
                    modSize = cols[col].size();
                    args[col] = cols[col][(row / divSize) % modSize];
                }
            }
            names.push_back(go_sprintf(info.SnapshotFormat, args));
        }
        names_.push_back(names);
    }
    return "";
}

// ----------------------------------------------------------------- memo

std::string ReadHeaders(int snap, VectorBuffer &buf, const Catalogs &cat, const std::string &memoDir,
                        std::vector<Header> &hds, std::vector<std::string> &files) {
    struct stat st;
    if (stat(memoDir.c_str(), &st) != 0) return "stat " + memoDir + ": " + std::strerror(errno);
    const std::string memoFile = memoDir + "/hd_snap" + std::to_string(snap) + ".dat";
    const int nb = cat.Blocks();
    files.resize(nb);
    hds.assign(nb, Header{});
    for (int i = 0; i < nb; i++) files[i] = cat.ParticleCatalog(snap, i);
    if (stat(memoFile.c_str(), &st) != 0) {
        for (int i = 0; i < nb; i++) {
            std::string err = buf.ReadHeader(files[i], hds[i]);
            if (!err.empty()) return err;
        }
        FILE *f = std::fopen(memoFile.c_str(), "wb");
        if (!f) return "open " + memoFile + ": " + std::strerror(errno);
        std::fwrite(hds.data(), sizeof(Header), hds.size(), f);   // binary.Write(LittleEndian, []io.Header)
        std::fclose(f);
    } else {
        FILE *f = std::fopen(memoFile.c_str(), "rb");
        if (!f) return "open " + memoFile + ": " + std::strerror(errno);
        size_t got = std::fread(hds.data(), sizeof(Header), hds.size(), f);
        (void)got;   // the reference ignores a short memo file as well
        std::fclose(f);
    }
    return "";
}

}  // namespace sfhost
