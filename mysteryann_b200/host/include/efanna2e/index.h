// Drop-in for include/efanna2e/index.h: the abstract base of IndexBipartite.
#pragma once
#include <cstddef>

#include "distance.h"
#include "parameters.h"

namespace efanna2e {

class Index {
   public:
    explicit Index(size_t dimension, size_t n, Metric metric);
    virtual ~Index();

    bool HasBuilt() const { return has_built; }
    size_t GetDimension() const { return dimension_; }
    size_t GetSizeOfDataset() const { return nd_; }
    const float *GetBasePointSet() const { return data_bp_; }
    const float *GetSampledQuerySet() const { return data_sq_; }
    const Distance *GetDistance() const { return distance_; }

   protected:
    const size_t dimension_;
    const float *data_sq_ = nullptr;
    const float *data_bp_ = nullptr;
    size_t nd_;
    size_t nd_sq_ = 0;
    bool has_built = false;
    Distance *distance_ = nullptr;  // host distance for graph construction (src/index.cpp:8-26 picks it by metric)
    Metric metric_;
};

}  // namespace efanna2e
