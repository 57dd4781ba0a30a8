// sf_catalog.hpp -- whitespace text tables: host-side mirror of
// cmd/catalog/catalog.go (Parse, FormatCols, CommentString).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace sfhost {

// fmt.Sprintf("%.6g", v) as Go prints it (NaN, +Inf, -Inf spelled Go's way).
std::string go_g6(double v);

// catalog.go:14-50
std::string CommentString(const std::vector<std::string> &intNames, const std::vector<std::string> &floatNames,
                          const std::vector<int> &order, const std::vector<int> &sizes);

// catalog.go:52-106: every column right-aligned to its own widest entry
std::vector<std::string> FormatCols(const std::vector<std::vector<int64_t>> &intCols,
                                    const std::vector<std::vector<double>> &floatCols,
                                    const std::vector<int> &order);

// catalog.go:143-150,261-326.  Returns "" or the reference's error text.
std::string Parse(const std::string &data, const std::vector<int> &icolIdxs, const std::vector<int> &fcolIdxs,
                  std::vector<std::vector<int64_t>> &icols, std::vector<std::vector<double>> &fcols);

}  // namespace sfhost
