// sf_catalog.cpp -- see sf_catalog.hpp; follows cmd/catalog/catalog.go.
#include "sf_catalog.hpp"

#include <cmath>
#include <cstdio>

#include "sf_config.hpp"

namespace sfhost {

std::string go_g6(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "+Inf" : "-Inf";
    char buf[64];
    snprintf(buf, sizeof buf, "%.6g", v);
    return buf;
}

std::string CommentString(const std::vector<std::string> &intNames, const std::vector<std::string> &floatNames,
                          const std::vector<int> &order, const std::vector<int> &sizes) {
    std::vector<std::string> tokens{"# Column contents:"};
    for (const auto &n : intNames) tokens.push_back(n);
    for (const auto &n : floatNames) tokens.push_back(n);
    std::vector<std::string> ordered{tokens[0]};
    std::vector<int> orderedSizes;
    for (int idx : order) {
        ordered.push_back(tokens.at(idx + 1));
        orderedSizes.push_back(sizes.at(idx));
    }
    int n = 0;
    for (size_t i = 1; i < ordered.size(); i++) {
        char buf[64];
        if (orderedSizes[i - 1] == 1) snprintf(buf, sizeof buf, "(%d)", n);
        else snprintf(buf, sizeof buf, "(%d-%d)", n, n + orderedSizes[i - 1] - 1);
        ordered[i] += buf;
        n += orderedSizes[i - 1];
    }
    std::string out;
    for (size_t i = 0; i < ordered.size(); i++) { if (i) out += " "; out += ordered[i]; }
    return out;
}

static std::vector<std::string> pad_left(const std::vector<std::string> &raw) {
    size_t width = 0;
    for (const auto &s : raw) width = s.size() > width ? s.size() : width;
    std::vector<std::string> out;
    for (const auto &s : raw) out.push_back(std::string(width - s.size(), ' ') + s);
    return out;
}

std::vector<std::string> FormatCols(const std::vector<std::vector<int64_t>> &intCols,
                                    const std::vector<std::vector<double>> &floatCols,
                                    const std::vector<int> &order) {
    if ((intCols.empty() && floatCols.empty()) || (!intCols.empty() && intCols[0].empty()) ||
        (!floatCols.empty() && floatCols[0].empty()))
        return {};
    std::vector<std::vector<std::string>> cols;
    for (const auto &c : intCols) {
        std::vector<std::string> raw;
        for (int64_t v : c) raw.push_back(std::to_string(v));
        cols.push_back(pad_left(raw));
    }
    for (const auto &c : floatCols) {
        std::vector<std::string> raw;
        for (double v : c) raw.push_back(go_g6(v));
        cols.push_back(pad_left(raw));
    }
    const size_t height = cols[0].size();
    std::vector<std::string> lines;
    for (size_t i = 0; i < height; i++) {
        std::string line;
        for (size_t j = 0; j < order.size(); j++) {
            if (j) line += " ";
            line += cols.at(order[j]).at(i);
        }
        lines.push_back(line);
    }
    return lines;
}

std::string Parse(const std::string &data, const std::vector<int> &icolIdxs, const std::vector<int> &fcolIdxs,
                  std::vector<std::vector<int64_t>> &icols, std::vector<std::vector<double>> &fcols) {
    // split on '\n', cut comments at '#', drop lines made only of spaces
    std::vector<std::string> lines;
    size_t start = 0;
    for (;;) {
        size_t nl = data.find('\n', start);
        std::string l = data.substr(start, nl == std::string::npos ? std::string::npos : nl - start);
        size_t c = l.find('#');
        if (c != std::string::npos) l = l.substr(0, c);
        bool blank = true;
        for (char ch : l) if (ch != ' ') { blank = false; break; }
        if (!blank) lines.push_back(l);
        if (nl == std::string::npos) break;
        start = nl + 1;
    }
    icols.assign(icolIdxs.size(), std::vector<int64_t>(lines.size()));
    fcols.assign(fcolIdxs.size(), std::vector<double>(lines.size()));
    if (lines.empty()) return "";
    auto fields = [](const std::string &l) {
        std::vector<std::string> w;
        size_t i = 0;
        while (i < l.size()) {
            while (i < l.size() && l[i] == ' ') i++;
            size_t j = i;
            while (j < l.size() && l[j] != ' ') j++;
            if (j > i) w.push_back(l.substr(i, j - i));
            i = j;
        }
        return w;
    };
    // the column count is fixed by the first row (bytes.Fields: any white space)
    size_t ncol = 0;
    {
        const std::string &l = lines[0];
        bool in = false;
        for (char ch : l) {
            const bool ws = ch == ' ' || ch == '\t' || ch == '\r' || ch == '\v' || ch == '\f';
            if (!ws && !in) ncol++;
            in = !ws;
        }
    }
    int maxCol = -1;
    for (int i : icolIdxs) maxCol = i > maxCol ? i : maxCol;
    for (int i : fcolIdxs) maxCol = i > maxCol ? i : maxCol;
    char buf[256];
    if (maxCol >= static_cast<int>(ncol)) {
        if (ncol == 1) snprintf(buf, sizeof buf, "Data has 1 column, but column %d was requested.", maxCol);
        else snprintf(buf, sizeof buf, "Data has %zu columns, but column %d was requested.", ncol, maxCol);
        return buf;
    }
    for (size_t i = 0; i < lines.size(); i++) {
        const std::vector<std::string> words = fields(lines[i]);
        if (words.size() != ncol) {
            snprintf(buf, sizeof buf, "Data on line %zu has %zu columns, not %zu.", i + 1, words.size(), ncol);
            return buf;
        }
        for (size_t j = 0; j < icolIdxs.size(); j++)
            if (!go_atoi(words[icolIdxs[j]], icols[j][i]))
                return "strconv.Atoi: parsing \"" + words[icolIdxs[j]] + "\": invalid syntax";
        for (size_t j = 0; j < fcolIdxs.size(); j++)
            if (!go_parse_float(words[fcolIdxs[j]], fcols[j][i]))
                return "strconv.ParseFloat: parsing \"" + words[fcolIdxs[j]] + "\": invalid syntax";
    }
    return "";
}

}  // namespace sfhost
