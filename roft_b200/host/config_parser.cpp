// config_parser.cpp - see config_parser.h
#include "config_parser.h"

#include <cctype>
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace ROFT {

namespace {

// hand-written token classifiers (no std::regex: this file is also loaded into processes that carry another C++ runtime's
// regex instantiations, e.g. a Python interpreter with torch, where the two collide)
bool is_int_token(const std::string& t) {
    std::size_t i = (t.size() && (t[0] == '-' || t[0] == '+')) ? 1 : 0;
    if (i >= t.size()) return false;
    for (; i < t.size(); ++i)
        if (!std::isdigit((unsigned char)t[i])) return false;
    return true;
}
bool is_hex_token(const std::string& t) {
    if (t.size() < 3 || t[0] != '0' || (t[1] != 'x' && t[1] != 'X')) return false;
    for (std::size_t i = 2; i < t.size(); ++i)
        if (!std::isxdigit((unsigned char)t[i])) return false;
    return true;
}
bool is_float_token(const std::string& t) {  // [-+]? (digits [. digits*] | . digits+) ([eE] [-+]? digits+)?
    std::size_t i = (t.size() && (t[0] == '-' || t[0] == '+')) ? 1 : 0;
    std::size_t nd = 0;
    while (i < t.size() && std::isdigit((unsigned char)t[i])) { ++i; ++nd; }
    if (i < t.size() && t[i] == '.') {
        ++i;
        while (i < t.size() && std::isdigit((unsigned char)t[i])) { ++i; ++nd; }
    }
    if (nd == 0) return false;
    if (i < t.size() && (t[i] == 'e' || t[i] == 'E')) {
        ++i;
        if (i < t.size() && (t[i] == '-' || t[i] == '+')) ++i;
        std::size_t ne = 0;
        while (i < t.size() && std::isdigit((unsigned char)t[i])) { ++i; ++ne; }
        if (ne == 0) return false;
    }
    return i == t.size();
}
std::string trim(const std::string& v) {
    std::size_t b = 0, e = v.size();
    while (b < e && std::isspace((unsigned char)v[b])) ++b;
    while (e > b && std::isspace((unsigned char)v[e - 1])) --e;
    return v.substr(b, e - b);
}

struct Lexer {
    const std::string& s;
    std::size_t i = 0;
    int line = 1;
    std::string origin;
    [[noreturn]] void fail(const std::string& what) const {
        throw std::runtime_error("ConfigParser::ctor. Parse error at " + origin + ":" + std::to_string(line) + " - " + what);
    }
    void skip() {  // whitespace and the three comment styles of libconfig
        for (;;) {
            while (i < s.size() && std::isspace((unsigned char)s[i])) {
                if (s[i] == '\n') ++line;
                ++i;
            }
            if (i < s.size() && s[i] == '#') {
                while (i < s.size() && s[i] != '\n') ++i;
            } else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '/') {
                while (i < s.size() && s[i] != '\n') ++i;
            } else if (i + 1 < s.size() && s[i] == '/' && s[i + 1] == '*') {
                i += 2;
                while (i + 1 < s.size() && !(s[i] == '*' && s[i + 1] == '/')) {
                    if (s[i] == '\n') ++line;
                    ++i;
                }
                if (i + 1 >= s.size()) fail("unterminated comment");
                i += 2;
            } else {
                return;
            }
        }
    }
    bool eof() { skip(); return i >= s.size(); }
    char peek() { skip(); return i < s.size() ? s[i] : '\0'; }
    bool accept(char c) {
        if (peek() == c) { ++i; return true; }
        return false;
    }
    void expect(char c) {
        if (!accept(c)) fail(std::string("syntax error: expected '") + c + "'");
    }
    std::string name() {
        skip();
        std::size_t b = i;
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '_' || s[i] == '-' || s[i] == '*')) ++i;
        if (b == i) fail("syntax error: expected a setting name");
        return s.substr(b, i - b);
    }
    // scalar token -> (type, text)
    std::pair<ConfigParser::Type, std::string> scalar() {
        skip();
        if (i < s.size() && s[i] == '"') {
            std::string out;
            for (;;) {  // adjacent strings concatenate, as in libconfig
                ++i;
                while (i < s.size() && s[i] != '"') {
                    if (s[i] == '\\' && i + 1 < s.size()) {
                        ++i;
                        out += s[i] == 'n' ? '\n' : s[i] == 't' ? '\t' : s[i];
                    } else {
                        if (s[i] == '\n') ++line;
                        out += s[i];
                    }
                    ++i;
                }
                if (i >= s.size()) fail("unterminated string");
                ++i;
                skip();
                if (!(i < s.size() && s[i] == '"')) break;
            }
            return {ConfigParser::Type::String, out};
        }
        std::size_t b = i;
        while (i < s.size() && (std::isalnum((unsigned char)s[i]) || s[i] == '.' || s[i] == '+' || s[i] == '-')) ++i;
        std::string tok = s.substr(b, i - b);
        if (tok.empty()) fail("syntax error: expected a value");
        std::string low = tok;
        for (char& c : low) c = (char)std::tolower((unsigned char)c);
        if (low == "true" || low == "false") return {ConfigParser::Type::Bool, low};
        std::string core = tok;
        while (!core.empty() && core.back() == 'L') core.pop_back();  // libconfig's 64-bit suffix
        if (is_int_token(core) || is_hex_token(core)) return {ConfigParser::Type::Int, core};
        if (is_float_token(tok)) return {ConfigParser::Type::Float, tok};
        fail("syntax error: bad value '" + tok + "'");
    }
};

void parse_group(Lexer& lx, const std::string& prefix, std::map<std::string, ConfigParser::Setting>& out, bool top) {
    for (;;) {
        if (top ? lx.eof() : lx.peek() == '}') return;
        if (lx.eof()) lx.fail("unexpected end of file inside a group");
        const std::string key = prefix.empty() ? lx.name() : prefix + "." + lx.name();
        if (!(lx.accept(':') || lx.accept('='))) lx.fail("syntax error: expected ':' or '=' after the setting name");
        if (lx.accept('{')) {
            parse_group(lx, key, out, false);
            lx.expect('}');
        } else if (lx.accept('[')) {
            ConfigParser::Setting s;
            s.type = ConfigParser::Type::Array;
            bool first = true;
            while (lx.peek() != ']') {
                if (!first) lx.expect(',');
                if (lx.peek() == ']') break;  // trailing comma
                auto v = lx.scalar();
                if (first) s.array_type = v.first;
                else if (v.first != s.array_type) {
                    // libconfig arrays are homogeneous; an integer literal in a float array is promoted
                    if ((s.array_type == ConfigParser::Type::Float && v.first == ConfigParser::Type::Int)) {
                    } else if (s.array_type == ConfigParser::Type::Int && v.first == ConfigParser::Type::Float) {
                        s.array_type = ConfigParser::Type::Float;
                    } else {
                        lx.fail("mismatched element type in array");
                    }
                }
                s.values.push_back(v.second);
                first = false;
            }
            lx.expect(']');
            out[key] = s;
        } else {
            auto v = lx.scalar();
            ConfigParser::Setting s;
            s.type = v.first;
            s.values = {v.second};
            out[key] = s;
        }
        if (!(lx.accept(';') || lx.accept(','))) { /* the terminator is optional in libconfig */ }
    }
}

}  // namespace

std::string ConfigParser::normalise(const std::string& path) {
    std::string out;
    for (std::size_t i = 0; i < path.size(); ++i) {
        if (path[i] == ':' && i + 1 < path.size() && path[i + 1] == ':') {
            out += '.';
            ++i;
        } else {
            out += path[i];
        }
    }
    return out;
}

void ConfigParser::parse(const std::string& text, const std::string& origin) {
    Lexer lx{text, 0, 1, origin};
    parse_group(lx, "", settings_, true);
}

void ConfigParser::override_setting(const std::string& path, const std::string& value) {
    auto it = settings_.find(normalise(path));
    if (it == settings_.end()) throw std::runtime_error("ConfigParser::ctor. Unknown option --" + path);
    Setting& s = it->second;
    auto check = [&](Type t, const std::string& v) {
        const bool ok = t == Type::Int ? is_int_token(v) : t == Type::Float ? is_float_token(v)
                        : t == Type::Bool ? (v == "true" || v == "false") : true;
        if (!ok) throw std::runtime_error("ConfigParser::ctor. Invalid value '" + v + "' for option --" + path);
    };
    if (s.type == Type::Array) {  // "a,b,c" (ConfigParser.cpp:57-101 validates the string with a regex, then splits it)
        std::vector<std::string> parts;
        std::stringstream ss(value);
        std::string item;
        while (std::getline(ss, item, ',')) {
            item = trim(item);
            check(s.array_type, item);
            parts.push_back(item);
        }
        if (parts.size() != s.values.size())
            throw std::runtime_error("ConfigParser::ctor. Option --" + path + " expects " + std::to_string(s.values.size()) + " comma separated values");
        s.values = parts;
    } else {
        check(s.type, value);
        s.values = {value};
    }
}

ConfigParser::ConfigParser(const int& argc, char** argv, const std::string& file_path) {
    // get_cfg_file_path: "--from <path>" wins over the default path (ConfigParser.cpp:136-157)
    std::string cfg_path = file_path;
    for (int i = 1; i + 1 < argc; ++i)
        if (std::string(argv[i]) == "--from") cfg_path = argv[i + 1];
    if (cfg_path.empty())
        throw std::runtime_error("ConfigParser::ctor. Please provide a valid configuration file using --from <path_to_cfg_file>");
    std::ifstream in(cfg_path);
    if (!in.is_open()) throw std::runtime_error("ConfigParser::ctor. I/O error while reading " + cfg_path + ".");
    std::stringstream ss;
    ss << in.rdbuf();
    parse(ss.str(), cfg_path);
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--from") { ++i; continue; }
        if (a.rfind("--", 0) != 0) throw std::runtime_error("ConfigParser::ctor. Unexpected argument " + a);
        if (i + 1 >= argc) throw std::runtime_error("ConfigParser::ctor. Missing value for option " + a);
        override_setting(a.substr(2), argv[++i]);
    }
}

ConfigParser ConfigParser::from_string(const std::string& text, const std::vector<std::pair<std::string, std::string>>& overrides) {
    ConfigParser c;
    c.parse(text, "<string>");
    for (const auto& o : overrides) c.override_setting(o.first, o.second);
    return c;
}

const ConfigParser::Setting& ConfigParser::lookup(const std::string& path) const {
    auto it = settings_.find(normalise(path));
    if (it == settings_.end()) throw std::runtime_error("ConfigParser::operator(). Error: cannot find the setting with name " + normalise(path));
    return it->second;
}

void ConfigParser::operator()(const std::string& path, double& value) const {
    const Setting& s = lookup(path);
    if (s.type != Type::Float && s.type != Type::Int) throw std::runtime_error("ConfigParser::operator(). Error: setting " + path + " is not a number");
    value = std::strtod(s.values[0].c_str(), nullptr);
}
void ConfigParser::operator()(const std::string& path, int& value) const {
    const Setting& s = lookup(path);
    if (s.type != Type::Int) throw std::runtime_error("ConfigParser::operator(). Error: setting " + path + " is not an integer");
    value = (int)std::strtol(s.values[0].c_str(), nullptr, 0);
}
void ConfigParser::operator()(const std::string& path, bool& value) const {
    const Setting& s = lookup(path);
    if (s.type != Type::Bool) throw std::runtime_error("ConfigParser::operator(). Error: setting " + path + " is not a boolean");
    value = s.values[0] == "true";
}
void ConfigParser::operator()(const std::string& path, std::string& value) const {
    const Setting& s = lookup(path);
    if (s.type != Type::String) throw std::runtime_error("ConfigParser::operator(). Error: setting " + path + " is not a string");
    value = s.values[0];
}
void ConfigParser::operator()(const std::string& path, std::vector<double>& array) const {
    const Setting& s = lookup(path);
    if (s.type != Type::Array) throw std::runtime_error("ConfigParser::operator(). Error: cannot find an array setting with name " + normalise(path));
    for (const std::string& v : s.values) array.push_back(std::strtod(v.c_str(), nullptr));
}
void ConfigParser::operator()(const std::string& path, std::vector<int>& array) const {
    const Setting& s = lookup(path);
    if (s.type != Type::Array || s.array_type != Type::Int)
        throw std::runtime_error("ConfigParser::operator(). Error: cannot find an integer array setting with name " + normalise(path));
    for (const std::string& v : s.values) array.push_back((int)std::strtol(v.c_str(), nullptr, 0));
}

}  // namespace ROFT

// ---- test hook (tests/test_host.py): parse a text with overrides and dump the flattened settings ---------------------------
extern "C" int rofth_config_dump(const char* text, const char* const* override_paths, const char* const* override_values, int n_overrides,
                                 char* out, int out_size) {
    try {
        std::vector<std::pair<std::string, std::string>> ov;
        for (int i = 0; i < n_overrides; ++i) ov.emplace_back(override_paths[i], override_values[i]);
        const ROFT::ConfigParser c = ROFT::ConfigParser::from_string(text, ov);
        std::string s;
        for (const auto& kv : c.settings()) {
            static const char* names[] = {"int", "float", "bool", "string", "array"};
            s += kv.first + "\t" + names[(int)kv.second.type] + "\t";
            for (std::size_t i = 0; i < kv.second.values.size(); ++i) s += (i ? "," : "") + kv.second.values[i];
            s += "\n";
        }
        if ((int)s.size() + 1 > out_size) return -2;
        std::copy(s.begin(), s.end(), out);
        out[s.size()] = '\0';
        return 0;
    } catch (const std::exception& e) {
        const std::string m = e.what();
        if ((int)m.size() + 1 <= out_size) { std::copy(m.begin(), m.end(), out); out[m.size()] = '\0'; }
        return -1;
    }
}
