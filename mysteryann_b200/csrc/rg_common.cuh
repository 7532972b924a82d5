// Shared helpers for libroargraph_b200.so: error plumbing, key encoding, sm_100a PTX wrappers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/roargraph_b200.h"

namespace rg {

// ---- host-side error plumbing ---------------------------------------------------------------
std::string &last_error();
rg_status fail(rg_status code, const char *fmt, ...);

#define RG_CUDA_OK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            rg_status _c = (_e == cudaErrorMemoryAllocation) ? RG_ERR_OUT_OF_MEMORY                   \
                           : (_e == cudaErrorNoDevice || _e == cudaErrorInsufficientDriver)           \
                               ? RG_ERR_NO_DEVICE                                                     \
                               : RG_ERR_CUDA;                                                         \
            return rg::fail(_c, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,     \
                            __LINE__);                                                                \
        }                                                                                             \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; }
        if (cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// ---- (distance, id, expanded) packed into one sortable 64-bit key ---------------------------
// Order of the reference's Neighbor (include/efanna2e/neighbor.h:29-31): distance, then id.
//   bits 63..32  monotone image of the FP32 distance (-0 canonicalised to +0)
//   bits 31..1   id (ids < 2^31)
//   bit  0       "expanded" flag (Neighbor::flag); never takes part in ordering because two pool
//                entries never share (distance, id).
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float d) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(d + 0.0f);
#else
    union { float f; uint32_t u; } c; c.f = d + 0.0f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
    return (uint64_t(float_to_ordered(d)) << 32) | (uint64_t(id) << 1);
}
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return uint32_t(k >> 1) & 0x7fffffffu; }
__host__ __device__ __forceinline__ float key_dist(uint64_t k) { return ordered_to_float(uint32_t(k >> 32)); }

#ifdef __CUDACC__
// ---- PTX wrappers -----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Ampere-style 16-byte async copy global -> shared, L2 only (LDGSTS.BYPASS).
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// mbarrier + TMA bulk (1-D) copy global -> shared (SASS: UBLKCP), completion by transaction bytes.
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// L2 eviction-priority policies (createpolicy) and the hinted bulk copy; bulk L2 prefetch.
// (atom.cas takes no .L2::cache_hint on sm_100a - the visited-hash slabs are pinned with an access-policy window instead.).
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                              uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// asynchronous prefetch of `bytes` (multiple of 16, 16-byte aligned source) into L2 (SASS: UBLKPF)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;\n" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// one warp sorts P (a power of two) 64-bit keys in shared memory, ascending (bitonic network)
__device__ __forceinline__ void warp_sort_u64(uint64_t *s, uint32_t P, uint32_t lane) {
    for (uint32_t k = 2; k <= P; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < P; i += 32) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const uint64_t a = s[i], b = s[x];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) {
                        s[i] = b;
                        s[x] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}
__device__ __forceinline__ uint32_t pow2_at_least(uint32_t n) {
    uint32_t p = 1;
    while (p < n) p <<= 1;
    return p;
}
#endif  // __CUDACC__

}  // namespace rg
