// sf_config.hpp -- `[name]`-headed `Var = value` config files and `--Flag v`
// overrides: host-side mirror of parse/config.go (same names, same rules, same
// error messages).
#pragma once

#include <cstdint>
#include <functional>
#include <string>
#include <vector>

namespace sfhost {

class ConfigVars {
public:
    explicit ConfigVars(std::string name) : name_(std::move(name)) {}
    // parse/config.go:221-275: register a variable with its default value
    void Int(int64_t *p, const std::string &name, int64_t def);
    void Float(double *p, const std::string &name, double def);
    void String(std::string *p, const std::string &name, const std::string &def);
    void Bool(bool *p, const std::string &name, bool def);
    void Ints(std::vector<int64_t> *p, const std::string &name, std::vector<int64_t> def);
    void Floats(std::vector<double> *p, const std::string &name, std::vector<double> def);
    void Strings(std::vector<std::string> *p, const std::string &name, std::vector<std::string> def);
    void Bools(std::vector<bool> *p, const std::string &name, std::vector<bool> def);

    const std::string &name() const { return name_; }

    // Returns "" on success, else the reference's error text.
    std::string ReadConfig(const std::string &fname);                 // parse/config.go:284
    std::string ReadConfigText(const std::string &text, const std::string &fname);
    std::string ReadFlags(const std::vector<std::string> &args);      // parse/config.go:361

private:
    enum Type { kInt, kInts, kFloat, kFloats, kString, kStrings, kBool, kBools };
    static const char *typeName(Type t);
    int find(const std::string &name) const;
    void add(const std::string &name, Type t, std::function<bool(const std::string &)> conv);
    std::string name_;
    std::vector<std::string> names_;
    std::vector<Type> types_;
    std::vector<std::function<bool(const std::string &)>> convs_;
};

// strconv.Atoi / ParseFloat / ParseBool as far as config files use them
bool go_atoi(const std::string &s, int64_t &out);
bool go_parse_float(const std::string &s, double &out);
bool go_parse_bool(const std::string &s, bool &out);
std::string trim_spaces(const std::string &s);   // strings.Trim(s, " ")
std::string to_lower(const std::string &s);

}  // namespace sfhost
