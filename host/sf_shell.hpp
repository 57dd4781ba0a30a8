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
// seed replaces cmd/cmd.go:658 randSeed; devices are the CUDA ordinals to use, one context
// (and, for haloAnalysis, one host thread) per entry.  With several entries the halos of a
// snapshot are partitioned by estimated particle count (PartitionHalos) -- the reference
// spreads each halo's particles over all cores instead (cmd/shell.go:311-314, 427-431) --
// every block is read once and pushed to the contexts whose halos touch it, and the rows
// come back in input order.  The catalog does not depend on the device list.
std::string RunShell(const GlobalConfig &g, const ShellConfig &c, const std::string &stdinData, uint64_t seed,
                     const std::vector<int> &devices, std::vector<std::string> &outLines);

// Longest-processing-time partition of halos over nRanks: weights[i] is the cost estimate of
// halo i, groups[i] an affinity label (the block holding its centre); halos of one group go
// to one rank.  Same result as shellfish_b200/shard.py partition_halos.  owner[i] in [0, nRanks).
void PartitionHalos(const std::vector<double> &weights, const std::vector<int64_t> &groups, int nRanks,
                    std::vector<int> &owner);
// Partition along a space-filling curve (what RunShell uses): halos sorted by the Z-curve key
// of their centre (10 bits per axis over totalWidth), cut into nRanks contiguous runs of equal
// weight.  Same result as shellfish_b200/shard.py partition_spatial.
void PartitionHalosSpatial(const std::vector<double> &weights, const std::vector<double> &x, const std::vector<double> &y,
                           const std::vector<double> &z, double totalWidth, int nRanks, std::vector<int> &owner);

}  // namespace sfhost
