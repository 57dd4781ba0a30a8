// sf_stats.hpp -- the `stats` mode (cmd/stats.go:22-360) on top of the sm_100a
// library: shell catalog on stdin -> M_sp, R_sp, volume, surface area, axes,
// radial range per halo.
#pragma once

#include <cstdint>
#include <string>
#include <vector>

#include "sf_shell.hpp"

namespace sfhost {

struct StatsConfig {
    std::vector<std::string> values;
    int64_t monteCarloSamples;
    std::string exclusionStrategy;
    int64_t order;
    bool skipMass;
    bool shellFilter;
    std::string shellParticleFile;
    double shellWidth;

    std::string ReadConfig(const std::string &fname, const std::vector<std::string> &flags);
    std::string validate() const;
    static const char *ExampleConfig();
};

// StatsConfig.Run, cmd/stats.go:174-360: stdin = the catalog `shellfish shell`
// printed; outLines = header comment + one row per halo.
std::string RunStats(const GlobalConfig &g, const StatsConfig &c, const std::string &stdinData, int device,
                     std::vector<std::string> &outLines);

// ProfConfig, cmd/prof.go:23-186.  Only ProfileType = angular-fraction is built: it is
// the one profile type for which the reference prints anything (Run returns nil, nil
// after the first particle block for every other type, cmd/prof.go:414).
struct ProfConfig {
    int64_t bins, order, samples;
    double rMaxMult, rMinMult;
    int64_t medianPixelLevel;
    double percentile;
    std::string pType;

    std::string ReadConfig(const std::string &fname, const std::vector<std::string> &flags);
    std::string validate() const;
    static const char *ExampleConfig();
};

// ProfConfig.Run for the angular-fraction type (cmd/prof.go:188-275, 628-661): stdin =
// the `shell` catalog, output = bin radii and the fraction of solid angle in which the
// shell lies beyond them.
std::string RunProf(const GlobalConfig &g, const ProfConfig &c, const std::string &stdinData, int device,
                    std::vector<std::string> &outLines);

}  // namespace sfhost
