// sf_config.cpp -- see sf_config.hpp; follows parse/config.go.
#include "sf_config.hpp"

#include <cerrno>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

namespace sfhost {

std::string trim_spaces(const std::string &s) {
    size_t a = 0, b = s.size();
    while (a < b && s[a] == ' ') a++;
    while (b > a && s[b - 1] == ' ') b--;
    return s.substr(a, b - a);
}

std::string to_lower(const std::string &s) {
    std::string o = s;
    for (char &c : o) if (c >= 'A' && c <= 'Z') c = static_cast<char>(c - 'A' + 'a');
    return o;
}

// strconv.Atoi: optional sign, decimal digits only, must consume everything
bool go_atoi(const std::string &s, int64_t &out) {
    if (s.empty()) return false;
    size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0;
    if (i == s.size()) return false;
    for (size_t k = i; k < s.size(); k++) if (s[k] < '0' || s[k] > '9') return false;
    errno = 0;
    char *end = nullptr;
    long long v = std::strtoll(s.c_str(), &end, 10);
    if (errno == ERANGE || *end) return false;
    out = v;
    return true;
}

// strconv.ParseFloat(s, 64): decimal / exponent / inf / nan, whole string
bool go_parse_float(const std::string &s, double &out) {
    if (s.empty() || s[0] == ' ' || s.back() == ' ') return false;
    std::string l = to_lower(s);
    // Go rejects hex floats without a p exponent and C accepts some forms Go
    // does not; config files only hold plain decimals, so keep the common set
    for (char c : l)
        if (!((c >= '0' && c <= '9') || c == '+' || c == '-' || c == '.' || c == 'e' || c == 'i' ||
              c == 'n' || c == 'f' || c == 'a' || c == 't' || c == 'y' || c == '_'))
            return false;
    if (l.find('_') != std::string::npos) return false;
    errno = 0;
    char *end = nullptr;
    double v = std::strtod(s.c_str(), &end);
    if (end == s.c_str() || *end) return false;
    if (errno == ERANGE && (v == HUGE_VAL || v == -HUGE_VAL)) return false;   // Go: value out of range
    out = v;
    return true;
}

// strconv.ParseBool
bool go_parse_bool(const std::string &s, bool &out) {
    static const char *t[] = {"1", "t", "T", "TRUE", "true", "True"};
    static const char *f[] = {"0", "f", "F", "FALSE", "false", "False"};
    for (const char *k : t) if (s == k) { out = true; return true; }
    for (const char *k : f) if (s == k) { out = false; return true; }
    return false;
}

static std::vector<std::string> str_to_list(const std::string &a) {   // config.go:153-159
    std::vector<std::string> out;
    size_t start = 0;
    for (;;) {
        size_t c = a.find(',', start);
        out.push_back(trim_spaces(a.substr(start, c == std::string::npos ? std::string::npos : c - start)));
        if (c == std::string::npos) break;
        start = c + 1;
    }
    return out;
}

const char *ConfigVars::typeName(Type t) {
    switch (t) {
    case kInt: return "int";
    case kInts: return "int list";
    case kFloat: return "float";
    case kFloats: return "float list";
    case kString: return "string";
    case kStrings: return "string list";
    case kBool: return "bool";
    case kBools: return "bool list";
    }
    return "?";
}

void ConfigVars::add(const std::string &name, Type t, std::function<bool(const std::string &)> conv) {
    names_.push_back(name);
    types_.push_back(t);
    convs_.push_back(std::move(conv));
}

void ConfigVars::Int(int64_t *p, const std::string &name, int64_t def) {
    *p = def;
    add(name, kInt, [p](const std::string &s) { return go_atoi(s, *p); });
}
void ConfigVars::Float(double *p, const std::string &name, double def) {
    *p = def;
    add(name, kFloat, [p](const std::string &s) { return go_parse_float(s, *p); });
}
void ConfigVars::String(std::string *p, const std::string &name, const std::string &def) {
    *p = def;
    add(name, kString, [p](const std::string &s) { *p = trim_spaces(s); return true; });
}
void ConfigVars::Bool(bool *p, const std::string &name, bool def) {
    *p = def;
    add(name, kBool, [p](const std::string &s) { return go_parse_bool(s, *p); });
}
void ConfigVars::Ints(std::vector<int64_t> *p, const std::string &name, std::vector<int64_t> def) {
    *p = std::move(def);
    add(name, kInts, [p](const std::string &s) {
        p->clear();
        for (const std::string &tok : str_to_list(s)) {
            int64_t v;
            if (!go_atoi(tok, v)) return false;
            p->push_back(v);
        }
        return true;
    });
}
void ConfigVars::Floats(std::vector<double> *p, const std::string &name, std::vector<double> def) {
    *p = std::move(def);
    add(name, kFloats, [p](const std::string &s) {
        p->clear();
        for (const std::string &tok : str_to_list(s)) {
            double v;
            if (!go_parse_float(tok, v)) return false;
            p->push_back(v);
        }
        return true;
    });
}
void ConfigVars::Strings(std::vector<std::string> *p, const std::string &name, std::vector<std::string> def) {
    *p = std::move(def);
    add(name, kStrings, [p](const std::string &s) { *p = str_to_list(s); return true; });
}
void ConfigVars::Bools(std::vector<bool> *p, const std::string &name, std::vector<bool> def) {
    *p = std::move(def);
    add(name, kBools, [p](const std::string &s) {
        p->clear();
        for (const std::string &tok : str_to_list(s)) {
            bool v;
            if (!go_parse_bool(tok, v)) return false;
            p->push_back(v);
        }
        return true;
    });
}

int ConfigVars::find(const std::string &name) const {
    const std::string l = to_lower(name);
    for (size_t j = 0; j < names_.size(); j++) if (to_lower(names_[j]) == l) return static_cast<int>(j);
    return -1;
}

namespace {

// config.go:464-488
void remove_comments(const std::vector<std::string> &in, std::vector<std::string> &out, std::vector<int> &nums) {
    for (size_t i = 0; i < in.size(); i++) {
        std::string l = in[i];
        size_t c = l.find('#');
        if (c != std::string::npos) l = l.substr(0, c);
        l = trim_spaces(l);
        if (l.empty()) continue;
        out.push_back(l);
        nums.push_back(static_cast<int>(i));
    }
}

// config.go:490-509; returns the offending line or -1
int association_list(const std::vector<std::string> &lines, std::vector<std::string> &names,
                     std::vector<std::string> &vals) {
    for (size_t i = 0; i < lines.size(); i++) {
        size_t eq = lines[i].find('=');
        if (eq == std::string::npos) return static_cast<int>(i);
        std::string name = trim_spaces(lines[i].substr(0, eq));
        std::string val = lines[i].size() - 1 > eq ? lines[i].substr(eq + 1) : "";
        names.push_back(name);
        if (name.empty()) return static_cast<int>(i);
        vals.push_back(trim_spaces(val));
    }
    return -1;
}

std::string fmt(const char *f, ...) __attribute__((format(printf, 1, 2)));
std::string fmt(const char *f, ...) {
    char buf[2048];
    va_list ap;
    va_start(ap, f);
    vsnprintf(buf, sizeof buf, f, ap);
    va_end(ap);
    return buf;
}

}  // namespace

std::string ConfigVars::ReadConfig(const std::string &fname) {
    std::ifstream f(fname, std::ios::binary);
    if (!f) return "open " + fname + ": " + std::strerror(errno);   // os.PathError text
    std::stringstream ss;
    ss << f.rdbuf();
    return ReadConfigText(ss.str(), fname);
}

std::string ConfigVars::ReadConfigText(const std::string &text, const std::string &fname) {
    std::vector<std::string> raw;
    size_t start = 0;
    for (;;) {
        size_t nl = text.find('\n', start);
        raw.push_back(text.substr(start, nl == std::string::npos ? std::string::npos : nl - start));
        if (nl == std::string::npos) break;
        start = nl + 1;
    }
    std::vector<std::string> lines;
    std::vector<int> nums;
    remove_comments(raw, lines, nums);
    for (int &n : nums) n++;
    if (lines.empty() || lines[0] != "[" + name_ + "]")
        return fmt("I expected the config file %s to have the header [%s] at the top, but didn't find it.",
                   fname.c_str(), name_.c_str());
    lines.erase(lines.begin());
    std::vector<std::string> names, vals;
    int bad = association_list(lines, names, vals);
    if (bad != -1)
        return fmt("I could not parse line %d of the config file %s because it did not take the form of a "
                   "variable assignment.", nums[bad + 1], fname.c_str());
    for (size_t i = 0; i < names.size(); i++)
        if (find(names[i]) < 0)
            return fmt("Line %d of the config file %s assigns a value to the variable '%s', but config files of "
                       "type %s don't have that variable.", nums[i + 1], fname.c_str(), names[i].c_str(),
                       name_.c_str());
    for (size_t i = 0; i < names.size(); i++)
        for (size_t j = i + 1; j < names.size(); j++)
            if (to_lower(names[i]) == to_lower(names[j]))
                return fmt("Lines %d and %d of the config file %s both assign a value to the variable '%s'.",
                           nums[i + 1], nums[j + 1], fname.c_str(), names[i].c_str());
    for (size_t i = 0; i < names.size(); i++) {
        const int j = find(names[i]);
        if (!convs_[j](vals[i])) {
            const char *tn = typeName(types_[j]);
            // the reference indexes vals with j here (config.go:352); kept when in range
            const std::string &shown = static_cast<size_t>(j) < vals.size() ? vals[j] : vals[i];
            return fmt("I could not parse line %d of the config file %s because '%s' expects values of type %s "
                       "and '%s' cannnot be converted to %s %s.", nums[i + 1], fname.c_str(), names_[j].c_str(),
                       tn, shown.c_str(), tn[0] == 'i' ? "an" : "a", tn);
        }
    }
    return "";
}

std::string ConfigVars::ReadFlags(const std::vector<std::string> &args) {
    if (args.empty()) return "";
    for (const std::string &a : args)
        if (a.find('=') != std::string::npos) return fmt("The argument '%s' contains an equals sign.", a.c_str());
    std::vector<bool> isFlag(args.size());
    for (size_t i = 0; i < args.size(); i++) isFlag[i] = args[i].size() > 1 && args[i].compare(0, 2, "--") == 0;
    if (!isFlag[0]) return fmt("The argument '%s' does not have a flag.", args[0].c_str());
    auto varName = [](const std::string &f) { size_t k = 0; while (k < f.size() && f[k] == '-') k++; return f.substr(k); };
    std::vector<std::string> varNames{varName(args[0])}, values, cur;
    auto join = [](const std::vector<std::string> &v) {
        std::string o;
        for (size_t i = 0; i < v.size(); i++) { if (i) o += ","; o += v[i]; }
        return o;
    };
    for (size_t i = 1; i < args.size(); i++) {
        if (!isFlag[i]) { cur.push_back(args[i]); continue; }
        values.push_back(join(cur));
        cur.clear();
        varNames.push_back(varName(args[i]));
    }
    values.push_back(join(cur));
    for (size_t i = 0; i < values.size(); i++)
        if (values[i].empty())
            return fmt("The flag '%s' was supplied, but wasn't set to a value.", varNames[i].c_str());
    std::vector<std::string> lines, names, vals;
    for (size_t i = 0; i < values.size(); i++) lines.push_back(varNames[i] + "=" + values[i]);
    if (association_list(lines, names, vals) != -1) return "Internal error! A flag could not be parsed.";
    for (const std::string &n : names)
        if (find(n) < 0) return fmt("The flag '%s' cannot be set for this program.", n.c_str());
    for (size_t i = 0; i < names.size(); i++)
        for (size_t j = i + 1; j < names.size(); j++)
            if (to_lower(names[i]) == to_lower(names[j])) return fmt("The flag '%s' was assigned twice.", names[i].c_str());
    for (size_t i = 0; i < names.size(); i++) {
        const int j = find(names[i]);
        if (!convs_[j](vals[i])) {
            const char *tn = typeName(types_[j]);
            const std::string &shown = static_cast<size_t>(j) < vals.size() ? vals[j] : vals[i];
            return fmt("I could not parse the flag '%s', because it expects values of type %s and '%s' cannnot be "
                       "converted to %s %s.", names_[j].c_str(), tn, shown.c_str(), tn[0] == 'i' ? "an" : "a", tn);
        }
    }
    return "";
}

}  // namespace sfhost
