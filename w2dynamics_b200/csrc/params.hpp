// Parameter-string handling of the engine boundary.
//
// Mirrors the semantics of the reference's Parameters.F90 (read_ParameterString 281-320,
// typed getters 77-224): a list of "key = value #\n" lines in the order Python sent them;
// lookup is a *prefix match on the first hit* (index(tmp%ID, ID) == 1), commas in values
// become blanks, tabs become blanks, everything after '#' is dropped, lines starting with
// '#' are skipped, integers may be written in scientific notation.
#pragma once
#include <cmath>
#include <cstdlib>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace w2b {

struct Params {
    std::vector<std::pair<std::string, std::string>> items;

    static std::string trim(const std::string& s) {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        if (a == std::string::npos) return "";
        return s.substr(a, b - a + 1);
    }
    void clear() { items.clear(); }
    void parse(const std::string& str) {
        items.clear();
        size_t pos = 0;
        while (pos < str.size()) {
            size_t nl = str.find('\n', pos);
            if (nl == std::string::npos) break;          // the reference only consumes newline-terminated lines
            std::string line = str.substr(pos, nl - pos);
            pos = nl + 1;
            if (!line.empty() && line[0] == '#') continue;
            for (char& c : line) if (c == '\t') c = ' ';
            size_t h = line.find('#');
            if (h != std::string::npos) line = line.substr(0, h);
            size_t eq = line.find('=');
            if (eq == std::string::npos) continue;
            std::string id = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
            for (char& c : val) if (c == ',') c = ' ';
            items.emplace_back(id, val);
        }
    }
    bool has(const std::string& id) const {
        for (auto& kv : items) if (kv.first.compare(0, id.size(), id) == 0) return true;
        return false;
    }
    const std::string& raw(const std::string& id) const {
        for (auto& kv : items) if (kv.first.compare(0, id.size(), id) == 0) return kv.second;
        throw std::runtime_error(id + " not found in parameter list!");
    }
    std::string str(const std::string& id) const { return raw(id); }
    std::string str_or(const std::string& id, const std::string& d) const { return has(id) ? raw(id) : d; }
    double real(const std::string& id) const {
        const std::string& v = raw(id);
        char* end = nullptr;
        double x = std::strtod(v.c_str(), &end);
        if (end == v.c_str()) throw std::runtime_error(id + " is not a real");
        return x;
    }
    double real_or(const std::string& id, double d) const { return has(id) ? real(id) : d; }
    long long integer(const std::string& id) const {
        const std::string& v = raw(id);
        char* end = nullptr;
        long long i = std::strtoll(v.c_str(), &end, 10);
        if (end != v.c_str() && (*end == '\0' || *end == ' ')) return i;
        double x = std::strtod(v.c_str(), &end);                 // scientific notation (Parameters.F90:94-101)
        if (end == v.c_str() || std::fabs(x - std::floor(x)) > 2.4e-7) throw std::runtime_error(id + " is not an integer");
        return (long long)x;
    }
    long long integer_or(const std::string& id, long long d) const { return has(id) ? integer(id) : d; }
    std::vector<int> integer_list(const std::string& id, size_t n) const {      // get_Integer_List (Parameters.F90:185-203): list-directed read of n integers
        std::istringstream in(raw(id));
        std::vector<int> v(n);
        for (size_t i = 0; i < n; ++i) if (!(in >> v[i])) throw std::runtime_error(id + " holds fewer than " + std::to_string(n) + " integers");
        return v;
    }
};

}  // namespace w2b
