// TEST INFRASTRUCTURE ONLY (oracle/): a minimal stand-in for boost::dynamic_bitset<> so that the
// reference translation units under /root/reference/src compile without Boost (not installed in
// this image).  Only the members the reference actually calls are provided
// (ctor(n[,v]), brace-init {n, v}, reset, reserve, test, set, operator[], size).
// This is OUR code, not copied from Boost; behaviour matches the documented Boost semantics for
// the used subset (value-initialised to `v`'s low bits, bits beyond 64 zero).
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace boost {

template <typename Block = unsigned long, typename Alloc = void>
class dynamic_bitset {
   public:
    class reference {
       public:
        reference(uint64_t &w, uint64_t m) : w_(w), m_(m) {}
        operator bool() const { return (w_ & m_) != 0; }
        reference &operator=(bool v) {
            if (v) w_ |= m_; else w_ &= ~m_;
            return *this;
        }
        reference &operator=(const reference &o) { return *this = static_cast<bool>(o); }

       private:
        uint64_t &w_;
        uint64_t m_;
    };

    dynamic_bitset() : n_(0) {}
    explicit dynamic_bitset(size_t n, unsigned long v = 0) : n_(n), w_((n + 63) / 64, 0) {
        if (!w_.empty()) {
            w_[0] = v;
            if (n < 64) w_[0] &= ((uint64_t(1) << n) - 1);
        }
    }
    size_t size() const { return n_; }
    void reserve(size_t n) { w_.reserve((n + 63) / 64); }
    dynamic_bitset &reset() {
        for (auto &w : w_) w = 0;
        return *this;
    }
    dynamic_bitset &reset(size_t i) {
        w_[i >> 6] &= ~(uint64_t(1) << (i & 63));
        return *this;
    }
    bool test(size_t i) const { return (w_[i >> 6] >> (i & 63)) & 1; }
    dynamic_bitset &set(size_t i, bool v = true) {
        if (v) w_[i >> 6] |= (uint64_t(1) << (i & 63));
        else w_[i >> 6] &= ~(uint64_t(1) << (i & 63));
        return *this;
    }
    bool operator[](size_t i) const { return test(i); }
    reference operator[](size_t i) { return reference(w_[i >> 6], uint64_t(1) << (i & 63)); }

   private:
    size_t n_;
    std::vector<uint64_t> w_;
};

}  // namespace boost
