// Drop-in for the reference's include/efanna2e/util.h: the fbin / ibin loaders and small helpers the CLI
// drivers call (same names, signatures, messages and exceptions), re-implemented on <cstdio> with
// chunked reads.  Formats (all little endian):
//   fbin      : u32 n, u32 d, f32[n*d]
//   truthset  : u32 n, u32 k, u32 ids[n*k], f32 dists[n*k]      (compute_groundtruth.cpp:325-343)
#pragma once
#include <malloc.h>

#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>

namespace efanna2e {

// rows are padded to a multiple of 8 floats and the buffer is 64-byte aligned (util.h:37-75, 179-211)
constexpr unsigned kRowAlignFloats = 8;
inline uint64_t padded_dim(uint64_t d) { return (d + kRowAlignFloats - 1) / kRowAlignFloats * kRowAlignFloats; }

namespace detail {
struct FileCloser {
    void operator()(FILE *f) const {
        if (f) fclose(f);
    }
};
using File = std::unique_ptr<FILE, FileCloser>;

inline File open_or_exit(const char *filename) {
    File f(fopen(filename, "rb"));
    if (!f) {
        std::cout << "open file error" << std::endl;  // util.h:87-90: the reference exits
        exit(-1);
    }
    return f;
}
inline uint64_t file_size(FILE *f) {
    long cur = ftell(f);
    fseek(f, 0, SEEK_END);
    uint64_t s = (uint64_t)ftell(f);
    fseek(f, cur, SEEK_SET);
    return s;
}
inline void read_header(FILE *f, unsigned &n, unsigned &d) {
    uint32_t h[2] = {0, 0};
    if (fread(h, sizeof(uint32_t), 2, f) != 2) throw std::runtime_error("Data file size wrong!");
    n = h[0];
    d = h[1];
}
template <typename T>
void check_count(const char *filename, FILE *f, unsigned n, unsigned d, unsigned expect_factor) {
    const uint64_t payload = file_size(f) - 2 * sizeof(uint32_t);
    const uint32_t contained = d ? (uint32_t)(payload / d / sizeof(T)) : 0;
    if (n * expect_factor != contained) {
        std::cerr << "filename: " << std::string(filename) << std::endl;
        std::cerr << "Data file size wrong! Get points " << contained << " but should have " << n << std::endl;
        throw std::runtime_error("Data file size wrong!");
    }
}
}  // namespace detail

// number of points and dimension of an fbin file (util.h:106-127)
template <typename T>
void load_meta(const char *filename, unsigned &points_num, unsigned &dim) {
    auto f = detail::open_or_exit(filename);
    detail::read_header(f.get(), points_num, dim);
    std::cout << "load meta from file: " << filename << " points_num: " << points_num << " dim: " << dim << std::endl;
    detail::check_count<T>(filename, f.get(), points_num, dim, 1);
}

// same for a truthset file, which holds ids AND distances, i.e. 2*n rows (util.h:84-105)
template <typename T>
void load_gt_meta(const char *filename, unsigned &points_num, unsigned &dim) {
    auto f = detail::open_or_exit(filename);
    detail::read_header(f.get(), points_num, dim);
    std::cout << "load gt from file: " << filename << " points_num: " << points_num << " dim: " << dim << std::endl;
    detail::check_count<T>(filename, f.get(), points_num, dim, 2);
}

// fbin payload into 8-float padded rows; points_num/dim come from load_meta (util.h:179-211).
// NOTE: like the reference, `dim` is NOT updated here - data_align() reports the padded length.
template <typename T>
void load_data(const char *filename, uint32_t &points_num, uint32_t &dim, T *&data) {
    std::cout << "load data from file: " << filename << std::endl;
    auto f = detail::open_or_exit(filename);
    fseek(f.get(), 2 * sizeof(uint32_t), SEEK_SET);
    const uint64_t n = points_num, d = dim, nd = padded_dim(d);
    data = static_cast<T *>(memalign(kRowAlignFloats * 8, n * nd * sizeof(T)));
    if (!data) throw std::bad_alloc();
    uint64_t got = 0;
    for (uint64_t i = 0; i < n; ++i) {
        got += fread(data + i * nd, sizeof(T), d, f.get());
        if (nd > d) memset(data + i * nd + d, 0, (nd - d) * sizeof(T));
    }
    if (got != n * d) {
        std::cerr << "Read file incompleted! filename:" << std::string(filename) << std::endl;
        throw std::runtime_error("Data file size wrong!");
    }
    std::cout << "load data from file: " << filename << " points_num: " << points_num << " dim: " << dim << std::endl;
}

// ids block then distances block of a truthset (util.h:129-155); arrays are new[]-allocated
template <typename T, typename T2>
void load_gt_data_with_dist(const char *filename, uint32_t &points_num, uint32_t &dim, T *&data, T2 *&res_dists) {
    auto f = detail::open_or_exit(filename);
    fseek(f.get(), 2 * sizeof(uint32_t), SEEK_SET);
    const uint64_t cnt = uint64_t(points_num) * dim;
    data = new T[cnt];
    res_dists = new T2[cnt];
    const uint64_t a = fread(data, sizeof(T), cnt, f.get());
    const uint64_t b = fread(res_dists, sizeof(T2), cnt, f.get());
    if (a != cnt || b != cnt) {
        std::cerr << "Read file incompleted!" << std::endl;
        throw std::runtime_error("Data file size wrong!");
    }
}

template <typename T>
void load_gt_data(const char *filename, uint32_t &points_num, uint32_t &dim, T *&data) {
    auto f = detail::open_or_exit(filename);
    fseek(f.get(), 2 * sizeof(uint32_t), SEEK_SET);
    const uint64_t cnt = uint64_t(points_num) * dim;
    data = new T[cnt];
    if (fread(data, sizeof(T), cnt, f.get()) != cnt) {
        std::cerr << "Read file incompleted!" << std::endl;
        throw std::runtime_error("Data file size wrong!");
    }
}

// Re-pack rows of `dim` floats into rows of padded_dim(dim) floats, 64-byte aligned; frees the input with
// delete[] semantics of the reference replaced by free()/delete[] detection is impossible, so the input must
// come from load_data (memalign) or new[] exactly as in the reference's drivers; `dim` becomes the padded length.
inline float *data_align(float *data_ori, unsigned point_num, unsigned &dim) {
    const uint64_t n = point_num, d = dim, nd = padded_dim(d);
    float *out = static_cast<float *>(memalign(kRowAlignFloats * 8, n * nd * sizeof(float)));
    if (!out) throw std::bad_alloc();
    for (uint64_t i = 0; i < n; ++i) {
        memcpy(out + i * nd, data_ori + i * d, d * sizeof(float));
        if (nd > d) memset(out + i * nd + d, 0, (nd - d) * sizeof(float));
    }
    dim = (unsigned)nd;
    std::cout << "new_dim: " << dim << std::endl;
    free(data_ori);  // load_data allocates with memalign (the reference calls delete[] on it, util.h:72)
    return out;
}

inline void prefetch_vector(const char *, size_t) {}  // CPU cache hint in the reference (util.h:77-80); no-op here

// L2-normalise one row in place (util.h:214-225)
template <typename T>
inline void normalize(T *arr, const size_t dim) {
    float sum = 0.0f;
    for (size_t i = 0; i < dim; i++) sum += arr[i] * arr[i];
    sum = std::sqrt(sum);
    for (size_t i = 0; i < dim; i++) arr[i] = (T)(arr[i] / sum);
}

}  // namespace efanna2e
