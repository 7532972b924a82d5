// Drop-in for the reference's efanna2e::Parameters (include/efanna2e/parameters.h:15-57): a
// string-keyed bag of values with typed Set/Get.  Same names, same exceptions
// (std::invalid_argument on a missing key, std::runtime_error on a failed conversion); written
// from scratch on an ordered map with a small conversion helper.
#pragma once
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <typeinfo>

namespace efanna2e {

class Parameters {
   public:
    template <typename T>
    void Set(const std::string &name, const T &value) {
        std::ostringstream os;
        os << value;
        values_[name] = os.str();
    }

    template <typename T>
    T Get(const std::string &name) const {
        auto it = values_.find(name);
        if (it == values_.end()) throw std::invalid_argument("Invalid parameter name.");
        return parse<T>(it->second);
    }

    template <typename T>
    T Get(const std::string &name, const T &fallback) const {
        auto it = values_.find(name);
        return it == values_.end() ? fallback : parse<T>(it->second);
    }

    bool Has(const std::string &name) const { return values_.count(name) != 0; }

   private:
    template <typename T>
    static T parse(const std::string &text) {
        std::istringstream is(text);
        T v;
        is >> v;
        if (is.fail() || !is.eof()) {
            throw std::runtime_error("Failed to convert value '" + text + "' to type: " + typeid(T).name());
        }
        return v;
    }
    std::map<std::string, std::string> values_;
};

}  // namespace efanna2e
