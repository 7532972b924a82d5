// Minimal stand-in for the subset of boost::program_options the reference's drivers use (Boost is not in this
// image): `--name value`, `--name=value`, multitoken values, short aliases (-T), required / default values.
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

class CliArgs {
   public:
    CliArgs(int argc, char **argv, const std::map<std::string, std::string> &aliases = {}) {
        std::string cur;
        for (int i = 1; i < argc; ++i) {
            std::string a = argv[i];
            const bool is_opt = a.size() > 1 && a[0] == '-' && !(std::isdigit((unsigned char)a[1]) || a[1] == '.');
            if (is_opt) {
                auto eq = a.find('=');
                std::string name = a.substr(0, eq);
                auto al = aliases.find(name);
                if (al != aliases.end()) name = al->second;
                while (!name.empty() && name[0] == '-') name.erase(0, 1);
                cur = name;
                values_[cur];
                if (eq != std::string::npos) values_[cur].push_back(a.substr(eq + 1));
            } else if (!cur.empty()) {
                values_[cur].push_back(a);
            } else {
                throw std::runtime_error("unexpected positional argument '" + a + "'");
            }
        }
    }
    bool has(const std::string &name) const { return values_.count(name) != 0; }

    template <typename T>
    T get(const std::string &name) const {
        auto it = values_.find(name);
        if (it == values_.end() || it->second.empty())
            throw std::runtime_error("the option '--" + name + "' is required but missing");
        return convert<T>(it->second.front(), name);
    }
    template <typename T>
    T get(const std::string &name, const T &dflt) const {
        auto it = values_.find(name);
        if (it == values_.end() || it->second.empty()) return dflt;
        return convert<T>(it->second.front(), name);
    }
    template <typename T>
    std::vector<T> get_all(const std::string &name) const {
        auto it = values_.find(name);
        if (it == values_.end() || it->second.empty())
            throw std::runtime_error("the option '--" + name + "' is required but missing");
        std::vector<T> out;
        for (const auto &s : it->second) out.push_back(convert<T>(s, name));
        return out;
    }

   private:
    template <typename T>
    static T convert(const std::string &s, const std::string &name) {
        std::istringstream is(s);
        T v;
        is >> v;
        if (is.fail() || !is.eof()) throw std::runtime_error("the argument ('" + s + "') for option '--" + name + "' is invalid");
        return v;
    }
    std::map<std::string, std::vector<std::string>> values_;
};
