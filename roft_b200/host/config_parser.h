// config_parser.h - ROFT::ConfigParser: the front-end of the ROFT-tracker executable (src/roft/src/ConfigParser.cpp:8-169,
// src/roft/include/ConfigParser.h:21-72) without libconfig++ / tclap (absent from this image): a parser for the subset
// of the libconfig grammar the reference's configuration files use (config/config_fast_ycb.cfg, config_ho3d.cfg) -
// nested groups `name: { ... }`, scalar settings `name = value;` (integer, float, boolean, string), homogeneous arrays
// `[a, b, c]`, `#`, `//` and `/* */` comments - and the same command line: `--from <file>` selects the file and EVERY leaf
// is overridable as `--group::sub::key value` (booleans as true / false, arrays as "a,b,c"), exactly the options the
// reference auto-generates (ConfigParser.cpp:57-133).  Lookup keeps the reference's call syntax:
//     ConfigParser conf(argc, argv);  double t;  conf("sample_time", t);  std::vector<double> c;  conf("a.b.cov", c);
// with either '.' or '::' as path separator, and std::runtime_error on a missing setting or a type mismatch.
#pragma once

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace ROFT {

class ConfigParser {
public:
    enum class Type { Int, Float, Bool, String, Array };
    struct Setting {
        Type type = Type::String;
        Type array_type = Type::Float;
        std::vector<std::string> values;  // one entry for scalars, the elements for arrays (strings unquoted)
    };

    ConfigParser(const int& argc, char** argv, const std::string& file_path = "");
    // parse a text directly (tests); `overrides` are (path, value) pairs as they would come from the command line
    static ConfigParser from_string(const std::string& text, const std::vector<std::pair<std::string, std::string>>& overrides = {});

    void operator()(const std::string& path, double& value) const;
    void operator()(const std::string& path, int& value) const;
    void operator()(const std::string& path, bool& value) const;
    void operator()(const std::string& path, std::string& value) const;
    void operator()(const std::string& path, std::vector<double>& array) const;
    void operator()(const std::string& path, std::vector<int>& array) const;
    bool exists(const std::string& path) const { return settings_.count(normalise(path)) != 0; }
    const std::map<std::string, Setting>& settings() const { return settings_; }  // flattened "a.b.c" -> setting

private:
    ConfigParser() = default;
    void parse(const std::string& text, const std::string& origin);
    void override_setting(const std::string& path, const std::string& value);
    const Setting& lookup(const std::string& path) const;
    static std::string normalise(const std::string& path);
    std::map<std::string, Setting> settings_;
};

}  // namespace ROFT
