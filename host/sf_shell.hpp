// sf_shell.hpp -- GlobalConfig (cmd/cmd.go:46-356), ShellConfig
// (cmd/shell.go:24-187) and the `shell` mode (cmd/shell.go:192-372) on top of
// the sm_100a library (include/shellfish_b200.h).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "sf_snapshot.hpp"

namespace sfhost {

extern const char *kSourceVersion;   // version/version.go:12

struct GlobalConfig {
    ParticleInfo particle;
    std::string HaloDir, TreeDir;
    std::string Version, SnapshotType, HaloType, TreeType, MemoDir;
    std::vector<std::string> HaloValueNames, HaloValueComments;
    std::vector<int64_t> HaloValueColumns;
    std::string HaloPositionUnits, HaloRadiusUnits, HaloMassUnits;
    std::string Endianness, Logging;
    bool ValidateFormats = false;
    int64_t Threads = -1;
    std::vector<int64_t> GadgetDMTypeIndices, GadgetSingleMassIndices;
    double GadgetPositionUnits = 1, GadgetMassUnits = 1;
    int64_t LGadgetNpartNum = 2;
    double NilSnapOmegaM = -1, NilSnapOmegaL = -1, NilSnapH100 = -1, NilSnapTotalWidth = -1;
    std::vector<double> NilSnapScaleFactors;

    std::string ReadConfig(const std::string &fname);   // "" or the reference's error text
    std::string validate();
};

struct ShellConfig {
    int64_t radialBins, spokes, rings;
    double rMaxMult, rMinMult, rKernelMult;
    bool percentileProfile;
    double percentile;
    double eta;
    int64_t order, smoothingWindow, levels, subsampleFactor;
    double losSlopeCutoff, backgroundRhoMult;

    std::string ReadConfig(const std::string &fname, const std::vector<std::string> &flags);
    std::string validate() const;
    static const char *ExampleConfig();
};

// ShellConfig.Run: stdin halo table -> catalog lines (header comment first).
// seed replaces cmd/cmd.go:658 randSeed; device is the CUDA ordinal.
std::string RunShell(const GlobalConfig &g, const ShellConfig &c, const std::string &stdinData, uint64_t seed,
                     int device, std::vector<std::string> &outLines);

}  // namespace sfhost
