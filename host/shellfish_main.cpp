// shellfish_main.cpp -- `shellfish shell [shell.config] [--Flag value ...]` and
// `shellfish stats [stats.config] [--Flag value ...]`: the reference's CLI
// protocol (shellfish.go:305-450) for the two modes this repository rebuilds, so
// that `... | shellfish shell | shellfish stats` runs end to end.  shell: stdin is
// the halo table `ID Snap X Y Z R200m`, stdout the Penna coefficient catalog;
// stats: stdin is that catalog, stdout the M_sp / R_sp / shape catalog.  Errors
// go to stderr and "Shellfish terminating." to stdout so that the next pipe stage
// aborts.
//
// New knobs are environment variables (config files stay valid for the Go
// binary): SHELLFISH_SEED pins cmd/cmd.go:658's wall-clock seed,
// SHELLFISH_GPUS / SHELLFISH_GPU pick the CUDA devices (default: all of the node).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>

#include <algorithm>

#include "../include/shellfish_b200.h"
#include "sf_config.hpp"
#include "sf_shell.hpp"
#include "sf_stats.hpp"

using namespace sfhost;

namespace {

[[noreturn]] void die(const std::string &mode, const std::string &msg) {
    std::fprintf(stderr, "Error running mode %s:\n%s\n", mode.c_str(), msg.c_str());
    std::printf("Shellfish terminating.\n");
    std::exit(1);
}

bool file_equal_config(const GlobalConfig &a, const GlobalConfig &b) {   // configEqual, shellfish.go:566-600
    return a.Version == b.Version && a.particle.SnapshotFormat == b.particle.SnapshotFormat &&
           a.SnapshotType == b.SnapshotType && a.HaloDir == b.HaloDir && a.HaloType == b.HaloType &&
           a.TreeDir == b.TreeDir && a.particle.BlockMins == b.particle.BlockMins &&
           a.particle.BlockMaxes == b.particle.BlockMaxes && a.particle.SnapMin == b.particle.SnapMin &&
           a.particle.SnapMax == b.particle.SnapMax &&
           a.particle.SnapshotFormatMeanings == b.particle.SnapshotFormatMeanings &&
           a.HaloValueNames == b.HaloValueNames && a.HaloValueColumns == b.HaloValueColumns &&
           a.HaloPositionUnits == b.HaloPositionUnits && a.HaloRadiusUnits == b.HaloRadiusUnits &&
           a.HaloMassUnits == b.HaloMassUnits && a.Endianness == b.Endianness;
}

std::string check_memo_dir(const std::string &memoDir, const std::string &configFile) {   // shellfish.go:491-521
    const std::string memoConfig = memoDir + "/memo.config";
    std::ifstream probe(memoConfig);
    if (!probe) {
        std::ifstream src(configFile, std::ios::binary);
        std::ofstream dst(memoConfig, std::ios::binary);
        if (!src || !dst) return "could not copy " + configFile + " to " + memoConfig;
        dst << src.rdbuf();
        return "";
    }
    GlobalConfig a, b;
    std::string err;
    if (!(err = a.ReadConfig(configFile)).empty()) return err;
    if (!(err = b.ReadConfig(memoConfig)).empty()) return err;
    if (!file_equal_config(a, b))
        return "You've changed the variables in the config file " + configFile +
               " in a way that would invlalidate the files Shellfish cached in " + memoDir +
               " (i.e. MemoDir) to speed up performance. Remove the contents of MemoDir (a copy of the old config "
               "is in " + memoConfig + ") or fix the MemoDir variable.";
    return "";
}

}  // namespace

int main(int argc, char **argv) {
    if (argc <= 1) {
        std::fprintf(stderr, "I was not supplied with a mode.\nFor help, type './shellfish help'.\n");
        return 1;
    }
    const std::string mode = argv[1];
    if (mode == "help") {
        if (argc == 3 && std::string(argv[2]) == "shell.config") std::printf("%s\n", ShellConfig::ExampleConfig());
        else if (argc == 3 && std::string(argv[2]) == "stats.config") std::printf("%s\n", StatsConfig::ExampleConfig());
        else if (argc == 3 && std::string(argv[2]) == "prof.config") std::printf("%s\n", ProfConfig::ExampleConfig());
        else std::printf("shellfish shell [shell.config] [--Flag value ...]\n"
                         "shellfish stats [stats.config] [--Flag value ...]\n"
                         "shellfish prof  [prof.config]  [--Flag value ...]   (B200 build: shell, stats and prof's angular-fraction)\n");
        return 0;
    }
    if (mode == "version") { std::printf("Shellfish version %s\n", kSourceVersion); return 0; }
    if (mode == "hello") { std::printf("Hello back at you! Installation was successful.\n"); return 0; }
    if (mode != "shell" && mode != "stats" && mode != "prof") {
        std::fprintf(stderr, "You passed me the mode '%s', which this B200 build does not provide.\n"
                             "For help, type './shellfish help'\n", mode.c_str());
        std::printf("Shellfish terminating.\n");
        return 1;
    }
    // all of stdin (shellfish.go:351-373)
    std::string stdinData((std::istreambuf_iterator<char>(std::cin)), std::istreambuf_iterator<char>());
    if (stdinData.empty()) return 0;
    size_t lineEnd = stdinData.find('\n');
    if (lineEnd == std::string::npos) lineEnd = stdinData.size();
    if (lineEnd >= 9 && stdinData.compare(0, 9, "Shellfish") == 0) {
        std::printf("%s\n", stdinData.substr(0, lineEnd).c_str());
        return 1;
    }
    std::vector<std::string> rest(argv + 2, argv + argc), flags;
    std::string configName;
    if (!rest.empty() && !rest[0].empty() && rest[0][0] != '-') { configName = rest[0]; flags.assign(rest.begin() + 1, rest.end()); }
    else flags = rest;

    const char *gName = std::getenv("SHELLFISH_GLOBAL_CONFIG");
    if (!gName || !*gName) die(mode, "$SHELLFISH_GLOBAL_CONFIG has not been set.");
    GlobalConfig g;
    std::string err = g.ReadConfig(gName);
    if (!err.empty()) die(mode, err);
    ShellConfig c;
    StatsConfig sc;
    ProfConfig pc;
    if (mode == "shell") err = c.ReadConfig(configName, flags);
    else if (mode == "stats") err = sc.ReadConfig(configName, flags);
    else err = pc.ReadConfig(configName, flags);
    if (!err.empty()) die(mode, err);
    if (!(err = check_memo_dir(g.MemoDir, gName)).empty()) die(mode, err);
    if (g.SnapshotType == "nil") {
        std::fprintf(stderr, "Cannot run mode %s with SnapshotType = nil\n", mode.c_str());
        std::printf("Shellfish terminating\n");
        return 1;
    }
    if (g.Logging != "nil" && g.Logging != "performance" && g.Logging != "debug") {
        std::fprintf(stderr, "Unrecognized logging mode, %s\n", g.Logging.c_str());
        return 1;
    }
    uint64_t seed = static_cast<uint64_t>(std::chrono::system_clock::now().time_since_epoch().count());
    if (const char *s = std::getenv("SHELLFISH_SEED")) seed = std::strtoull(s, nullptr, 10);
    // SHELLFISH_GPUS = comma-separated CUDA ordinals (or "all"), one context per entry;
    // SHELLFISH_GPU = one ordinal.  Default: every GPU of the node, as the reference uses every
    // core (Threads = -1, cmd/shell.go:311-314).  `stats` and `prof` use the first entry.
    std::vector<int> devices;
    if (const char *s = std::getenv("SHELLFISH_GPUS")) {
        if (std::string(s) != "all")
            for (const char *q = s; *q;) {
                char *end = nullptr;
                const long v = std::strtol(q, &end, 10);
                if (end == q) die(mode, std::string("Cannot parse $SHELLFISH_GPUS = '") + s + "'.");
                devices.push_back(static_cast<int>(v));
                q = *end == ',' ? end + 1 : end;
                if (*end && *end != ',') die(mode, std::string("Cannot parse $SHELLFISH_GPUS = '") + s + "'.");
            }
    } else if (const char *s1 = std::getenv("SHELLFISH_GPU")) {
        devices.push_back(std::atoi(s1));
    }
    if (devices.empty())
        for (int d = 0; d < std::max(1, sfk_device_count()); d++) devices.push_back(d);
    const int device = devices[0];
    if (g.Logging != "nil") {
        std::fprintf(stderr, "\n#####################\n## shellfish %s ##\n#####################\n", mode.c_str());
        if (mode == "shell") std::fprintf(stderr, "RNG Seed is %llu\n", static_cast<unsigned long long>(seed));
    }
    std::vector<std::string> lines;
    try {
        if (mode == "shell") err = RunShell(g, c, stdinData, seed, devices, lines);
        else if (mode == "stats") err = RunStats(g, sc, stdinData, device, lines);
        else err = RunProf(g, pc, stdinData, device, lines);
    } catch (const std::exception &e) {   // a Go panic ends the same way: message, then the pipe sentinel
        err = std::string("internal error: ") + e.what();
    }
    if (!err.empty()) die(mode, err);
    for (const std::string &l : lines) std::printf("%s\n", l.c_str());
    return 0;
}
