// Drop-in for the parts of include/efanna2e/neighbor.h the RoarGraph path uses: Neighbor and the bounded
// sorted candidate pool NeighborPriorityQueue.  Host-side only (graph construction); the search kernels keep
// the same pool as packed 64-bit keys in shared memory (csrc/rg_search.cu).
#pragma once
#include <algorithm>
#include <cstddef>
#include <vector>

namespace efanna2e {

struct Neighbor {
    unsigned id = 0;
    float distance = 0.f;
    bool flag = false;  // "expanded"

    Neighbor() = default;
    Neighbor(unsigned id_, float distance_, bool f) : id(id_), distance(distance_), flag(f) {}

    // strict order (distance, id) - neighbor.h:29-31
    bool operator<(const Neighbor &o) const { return distance < o.distance || (distance == o.distance && id < o.id); }
    bool operator==(const Neighbor &o) const { return id == o.id; }  // neighbor.h:33 (id only)
};

// Sorted array of at most `capacity` neighbours plus a cursor to the closest unexpanded one
// (neighbor.h:138-223).  insert() keeps the array ordered, refuses an id that the binary search meets on its
// way (neighbor.h:161) and anything worse than a full pool's last entry (neighbor.h:151).
class NeighborPriorityQueue {
   public:
    NeighborPriorityQueue() = default;
    explicit NeighborPriorityQueue(size_t capacity) : cap_(capacity), slots_(capacity + 1) {}

    void insert(const Neighbor &nbr) {
        if (size_ == cap_ && slots_[size_ - 1] < nbr) return;
        size_t lo = 0, hi = size_;
        while (lo < hi) {
            const size_t mid = (lo + hi) / 2;
            if (nbr < slots_[mid]) hi = mid;
            else if (slots_[mid].id == nbr.id) return;
            else lo = mid + 1;
        }
        if (lo < cap_) std::move_backward(slots_.begin() + lo, slots_.begin() + size_, slots_.begin() + size_ + 1);
        slots_[lo] = Neighbor(nbr.id, nbr.distance, false);
        if (size_ < cap_) ++size_;
        if (lo < cursor_) cursor_ = lo;
    }

    Neighbor closest_unexpanded() {
        slots_[cursor_].flag = true;
        const size_t taken = cursor_;
        while (cursor_ < size_ && slots_[cursor_].flag) ++cursor_;
        return slots_[taken];
    }

    bool has_unexpanded_node() const { return cursor_ < size_; }
    size_t size() const { return size_; }
    size_t capacity() const { return cap_; }
    void reserve(size_t capacity) {
        if (capacity + 1 > slots_.size()) slots_.resize(capacity + 1);
        cap_ = capacity;
    }
    Neighbor &operator[](size_t i) { return slots_[i]; }
    Neighbor operator[](size_t i) const { return slots_[i]; }
    void clear() { size_ = cursor_ = 0; }

   private:
    size_t size_ = 0, cap_ = 0, cursor_ = 0;
    std::vector<Neighbor> slots_;
};

}  // namespace efanna2e
