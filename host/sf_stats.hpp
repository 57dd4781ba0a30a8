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

}  // namespace sfhost
