// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), named barriers.
#pragma once
#include <cstdint>
#include <cuda.h> // CUtensorMap (type only; the encoder is fetched through cudaGetDriverEntryPoint)

namespace cvvp
{
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
                 "selp.u32 %0, 1, 0, P1;\n"
                 "}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return ok != 0;
}

// try_wait that may stay suspended for up to `ns` nanoseconds before it reports failure (the default limit is a few
// tens of cycles, so a waiting warp otherwise keeps issuing)
__device__ __forceinline__ bool mbar_try_wait_for(uint64_t *bar, uint32_t parity, uint32_t ns)
{
    uint32_t ok;
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
                 "selp.u32 %0, 1, 0, P1;\n"
                 "}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
                 : "memory");
    return ok != 0;
}

__device__ __forceinline__ uint64_t global_timer_ns()
{
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifndef CVVP_MBAR_SUSPEND_NS
#define CVVP_MBAR_SUSPEND_NS 2000u
#endif

// Wait for the phase with the given parity.  NOTE (parity aliasing): a waiter may only start waiting
// for fill q of a slot once fill q-1 of that slot has completed, otherwise the test passes early.
// A watchdog turns any protocol bug into a trap (sticky CUDA error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    if (mbar_try_wait(bar, parity))
        return;
    // The spin loop is kept to try_wait + counter: a waiting warp shares its scheduler (and the integer pipe) with
    // warps that have work, and reading the timer in every iteration made the waiters of median_pipe_kernel issue
    // as many integer instructions as the workers.  The timer is read once per 4096 failed tries.
    const uint64_t t0 = global_timer_ns();
    for (;;) {
#pragma unroll 1
        for (uint32_t spins = 0; spins < 4096u; ++spins)
            if (mbar_try_wait_for(bar, parity, CVVP_MBAR_SUSPEND_NS))
                return;
        if (global_timer_ns() - t0 > 2000000000ull)
            __trap();
    }
}

// L2 eviction-priority policies (createpolicy.fractional encodings used by cp.async.bulk .L2::cache_hint)
constexpr uint64_t kL2EvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kL2EvictLast = 0x14F0000000000000ull;
constexpr uint64_t kL2EvictNormal = 0x1000000000000000ull;

// 2-D tiled TMA load global -> shared, completion signalled on an mbarrier (complete_tx::bytes).
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *tmap, uint64_t *bar, int32_t x, int32_t y,
                                            uint64_t l2_policy)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
                 " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(smem_dst)),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y), "l"(l2_policy)
                 : "memory");
}

// Bulk prefetch of a contiguous global range into L2 (a hint: no destination, no completion to wait for).
// addr and bytes must be multiples of 16.
__device__ __forceinline__ void prefetch_l2_bulk(const void *gptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// bar.sync on a named barrier among `nthreads` threads (id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
} // namespace cvvp
