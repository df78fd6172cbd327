// One-dimensional bulk asynchronous copies (the TMA's cp.async.bulk) into shared memory, completion on an mbarrier: the
// building block of the streaming kernels' shared-memory rings (keypoint decode slabs, heatmap tiles).  One elected thread
// requests a copy; everybody waits on the barrier's phase parity.  A kernel fed this way keeps tens of KB per SM in flight
// without spending a register on it, which is what a stream against HBM latency needs.
#pragma once
#include <cuda_runtime.h>

namespace mpn {

__device__ __forceinline__ unsigned smem_addr(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// once, by one thread, before anybody waits (arrival count 1: the requesting thread's expect_tx arrival)
__device__ __forceinline__ void bulk_barrier_init(unsigned long long *bars, int n)
{
    for (int i = 0; i < n; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

// Bounded wait: a protocol bug is reported as a CUDA error instead of hanging the GPU.
__device__ __forceinline__ void bulk_wait(unsigned long long *bar, unsigned parity)
{
    const unsigned addr = smem_addr(bar);
    for (unsigned spin = 0; spin < (1u << 26); ++spin) {
        unsigned done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}

// one thread: request `bytes` (multiple of 16, 16-byte aligned source and destination) of global memory into shared
// memory, completion on `bar`.  Earlier generic-proxy reads of `dst` must be ordered before the call by a barrier.
__device__ __forceinline__ void bulk_fetch(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

}  // namespace mpn
