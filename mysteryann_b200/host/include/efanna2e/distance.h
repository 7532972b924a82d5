// Host-side FP32 distances with the same class names as the reference (include/efanna2e/distance.h)
// and bit-identical results, written as portable scalar-lane code instead of AVX-512 intrinsics.
//
// They are used ONLY by the CPU graph construction (BuildRoarGraph), which stays on the host exactly
// as in the reference; the search and kNN hot paths run these operation orders in CUDA
// (mysteryann_b200/csrc/rg_search.cu, rg_knn.cu).
//
// Operation order = what g++ -Ofast emits for the reference's intrinsics (DESIGN.md "Distances"):
// 16 independent lane sums with separate multiply and add, fold hi+lo to 8 lanes, fused
// multiply-add for the 8-, 4- and <4-wide tails, fold to 4 lanes, (x0+x1)+(x2+x3).
// Compile with -ffp-contract=off (host/Makefile does) so the compiler fuses nothing on its own.
#pragma once
#include <cmath>
#include <cstddef>

namespace efanna2e {

enum Metric { L2 = 0, INNER_PRODUCT = 1, FAST_L2 = 2, PQ = 3, COSINE = 4 };

class Distance {
   public:
    virtual float compare(const float *a, const float *b, unsigned length) const = 0;
    virtual ~Distance() {}
};

namespace detail {
typedef float v16f __attribute__((vector_size(64), aligned(4)));
typedef float v8f __attribute__((vector_size(32), aligned(4)));
typedef float v4f __attribute__((vector_size(16), aligned(4)));

template <typename V>
__attribute__((always_inline)) inline V loadu(const float *p) {
    V v;
    __builtin_memcpy(&v, p, sizeof(V));
    return v;
}
template <typename V>
__attribute__((always_inline)) inline V fused(V x, V y, V acc) {  // per-lane fma: one rounding, like vfmadd231ps
    constexpr int n = sizeof(V) / sizeof(float);
    V out;
    for (int l = 0; l < n; ++l) out[l] = __builtin_fmaf(x[l], y[l], acc[l]);
    return out;
}

// kSquaredDiff = false: sum a[i]*b[i];  true: sum (a[i]-b[i])^2.
// Vector-extension code: each `+`/`*` below is one IEEE operation per lane (the build disables contraction),
// whatever instruction set the clone is compiled for.
template <bool kSquaredDiff>
__attribute__((target_clones("arch=x86-64-v4", "arch=x86-64-v3", "default"))) inline float lane_ordered_sum(
    const float *a, const float *b, unsigned len) {
    v16f lanes = {0};
    unsigned i = 0;
    for (; i + 16 <= len; i += 16) {
        v16f x = loadu<v16f>(a + i), y = loadu<v16f>(b + i);
        if (kSquaredDiff) {
            x = x - y;
            y = x;
        }
        const v16f prod = x * y;
        lanes = lanes + prod;
    }
    v8f oct;
    for (int l = 0; l < 8; ++l) oct[l] = lanes[l + 8] + lanes[l];
    if (len - i >= 8) {
        v8f x = loadu<v8f>(a + i), y = loadu<v8f>(b + i);
        if (kSquaredDiff) {
            x = x - y;
            y = x;
        }
        oct = fused(x, y, oct);
        i += 8;
    }
    v4f quad;
    for (int l = 0; l < 4; ++l) quad[l] = oct[l + 4] + oct[l];
    if (len - i >= 4) {
        v4f x = loadu<v4f>(a + i), y = loadu<v4f>(b + i);
        if (kSquaredDiff) {
            x = x - y;
            y = x;
        }
        quad = fused(x, y, quad);
        i += 4;
    }
    if (len - i > 0) {  // zero-filled partial vector (masked_read, distance.h:23-36)
        v4f x = {0, 0, 0, 0}, y = {0, 0, 0, 0};
        for (unsigned l = 0; l < len - i; ++l) {
            x[l] = a[i + l];
            y[l] = b[i + l];
        }
        if (kSquaredDiff) {
            x = x - y;
            y = x;
        }
        quad = fused(x, y, quad);
    }
    const float lo = quad[0] + quad[1];
    const float hi = quad[2] + quad[3];
    return lo + hi;
}
}  // namespace detail

class DistanceL2 : public Distance {
   public:
    float compare(const float *a, const float *b, unsigned length) const override {
        return detail::lane_ordered_sum<true>(a, b, length);
    }
};

// Returns the NEGATED dot product, like the reference (distance.h:222).
class DistanceInnerProduct : public Distance {
   public:
    float compare(const float *a, const float *b, unsigned length) const override {
        return -detail::lane_ordered_sum<false>(a, b, length);
    }
};

}  // namespace efanna2e
