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
// kSquaredDiff = false: sum a[i]*b[i];  true: sum (a[i]-b[i])^2
template <bool kSquaredDiff>
inline float lane_ordered_sum(const float *a, const float *b, unsigned len) {
    float lanes[16] = {0};
    unsigned i = 0;
    for (; i + 16 <= len; i += 16) {
        for (int l = 0; l < 16; ++l) {
            const float x = kSquaredDiff ? a[i + l] - b[i + l] : a[i + l];
            const float y = kSquaredDiff ? x : b[i + l];
            const float prod = x * y;
            lanes[l] = lanes[l] + prod;
        }
    }
    float oct[8];
    for (int l = 0; l < 8; ++l) oct[l] = lanes[l + 8] + lanes[l];
    auto fused_tail = [&](float *acc, unsigned width, unsigned count) {
        for (unsigned l = 0; l < width; ++l) {
            float x = 0.f, y = 0.f;
            if (l < count) {
                x = kSquaredDiff ? a[i + l] - b[i + l] : a[i + l];
                y = kSquaredDiff ? x : b[i + l];
            }
            acc[l] = std::fmaf(x, y, acc[l]);
        }
    };
    if (len - i >= 8) {
        fused_tail(oct, 8, 8);
        i += 8;
    }
    float quad[4];
    for (int l = 0; l < 4; ++l) quad[l] = oct[l + 4] + oct[l];
    if (len - i >= 4) {
        fused_tail(quad, 4, 4);
        i += 4;
    }
    if (len - i > 0) fused_tail(quad, 4, len - i);
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
