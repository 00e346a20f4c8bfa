// common.cuh — shared device helpers for libqscuda (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qs {

// C(x,2), C(x,3), C(x,4) — QuartetLookupTable rank terms (reference src/quartet_lookup_table.hpp:141-168)
__host__ __device__ __forceinline__ uint64_t binom2(uint64_t x) { return x * (x - 1) / 2; }
__host__ __device__ __forceinline__ uint64_t binom3(uint64_t x) { return x < 3 ? 0 : x * (x - 1) * (x - 2) / 6; }
__host__ __device__ __forceinline__ uint64_t binom4(uint64_t x) { return x < 4 ? 0 : x * (x - 1) * (x - 2) * (x - 3) / 24; }
__host__ __device__ __forceinline__ uint64_t quartet_rank(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    return binom4(d) + binom3(c) + (b * (b - 1) / 2) + a;   // a<b<c<d
}

// ---- mbarrier + bulk-copy (TMA, non-tensor) PTX wrappers -----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// the same for a thread that has nothing else to do (the scan's producer lane): back off between polls instead of spinning in the issue slots
__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
    while (true) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
        __nanosleep(256);
    }
}
// global -> shared bulk copy (UBLKCP), completion signalled on an mbarrier of this CTA
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint4 lds128(const unsigned char* base, uint32_t byte_off) {
    return *reinterpret_cast<const uint4*>(base + byte_off);
}
__device__ __forceinline__ __half2 as_h2(uint32_t x) { return *reinterpret_cast<__half2*>(&x); }

}  // namespace qs
