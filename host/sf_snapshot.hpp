// sf_snapshot.hpp -- snapshot readers behind the VectorBuffer interface, block
// headers, file-name expansion and the header memo: host-side mirror of
// io/io.go, io/{lgadget2,gadget2,gotetra}.go, cmd/env/ and
// cmd/memo/memoize.go:242-302.  The three formats BASELINE's configs name.
#pragma once

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

namespace sfhost {

// io.CosmologyHeader / io.Header, io/io.go:71-83.  72 bytes, written verbatim
// (little endian) to MemoDir/hd_snap<snap>.dat like binary.Write does.
struct Cosmology { double Z, OmegaM, OmegaL, H100; };
struct Header {
    Cosmology Cosmo;
    int64_t N;
    double TotalWidth;
    float Origin[3], Width[3];
};
static_assert(sizeof(Header) == 72, "io.Header layout");

// io.Context, io/io.go:85-98 (fields the three readers use)
struct Context {
    int64_t LGadgetNPartNum = 2;
    std::vector<int64_t> GadgetDMTypeIndices{1}, GadgetDMSingleMassIndices{1};
    double GadgetMassUnits = 1.0, GadgetPositionUnits = 1.0;
};

// io.VectorBuffer, io/io.go:58-67 (positions and masses only: the shell path
// never looks at velocities or ids).  Errors are returned as text, "" = ok.
class VectorBuffer {
public:
    virtual ~VectorBuffer() = default;
    // xs: [N][3] Mpc/h, ms: [N] Msun/h.  The arrays are owned by the buffer
    // and reused by the next Read, like the reference's slices.
    virtual std::string Read(const std::string &fname, const float **xs, const float **ms, int64_t *n) = 0;
    // The particle IDs Read's fourth return value holds (only `stats` with ShellParticleFile
    // asks for them): ids[i] belongs to xs[i] of the same file.  Owned by the buffer.
    virtual std::string ReadIDs(const std::string &fname, const int64_t **ids, int64_t *n) = 0;
    virtual std::string ReadHeader(const std::string &fname, Header &out) = 0;
    virtual float MinMass() const = 0;
    // true when Read fills ms with MinMass() for every particle (LGadget-2, gotetra): the caller may
    // then push blocks without a mass array (sfk_push_block with mass == NULL)
    virtual bool UniformMass() const { return false; }
    virtual std::string TotalParticles(const std::string &fname, int64_t &n) = 0;
};

// cmd/utils.go:8-44 getVectorBuffer
std::string NewVectorBuffer(const std::string &snapshotType, const std::string &fname,
                            const std::string &endianness, const Context &ctx,
                            std::unique_ptr<VectorBuffer> &out);

// io.WriteFilter, io/filter_data.go:36-63: the ShellParticleFile of `shellfish stats`.  Layout:
// int32 endianness flag (0 little, -1 big), int32 halo count, per halo {ID, Snap, StartByte, Len}
// int64, then every halo's particle IDs.  StartByte is the same for every halo -- the reference
// never advances totLen -- and ReadFilter ignores it; kept, so that the bytes are the reference's.
std::string WriteFilter(const std::string &path, const std::string &endianness, const std::vector<int64_t> &snaps,
                        const std::vector<int64_t> &ids, const std::vector<std::vector<int64_t>> &particles);

// io/io.go:263-304
void boundingBox(const float *xs, int64_t n, double totalWidth, float origin[3], float width[3]);
// io/gotetra.go:50-54
float calcUniformMass(int64_t count, double tw, const Cosmology &c);

// env.ParticleInfo + Catalogs, cmd/env/env.go:47-56,140-151 and Init{LGadget2,Gadget2,Gotetra}
struct ParticleInfo {
    std::string SnapshotFormat, ScaleFactorFile;
    std::vector<std::string> SnapshotFormatMeanings;
    std::vector<int64_t> BlockMins, BlockMaxes;
    int64_t SnapMin = -1, SnapMax = -1;
};
class Catalogs {
public:
    std::string Init(const ParticleInfo &info, bool gotetra);
    int Blocks() const { return names_.empty() ? 0 : static_cast<int>(names_[0].size()); }
    const std::string &ParticleCatalog(int snap, int block) const { return names_.at(snap - snapMin_).at(block); }
private:
    int snapMin_ = 0;
    std::vector<std::vector<std::string>> names_;
};
// fmt.Sprintf(format, args...) for the verbs snapshot names use (%d %0Nd %s %%);
// every argument is either an int or a string
struct FormatArg { bool isInt; int64_t i; std::string s; };
std::string go_sprintf(const std::string &format, const std::vector<FormatArg> &args);

// memo.ReadHeaders, cmd/memo/memoize.go:260-302
std::string ReadHeaders(int snap, VectorBuffer &buf, const Catalogs &cat, const std::string &memoDir,
                        std::vector<Header> &hds, std::vector<std::string> &files);

}  // namespace sfhost
