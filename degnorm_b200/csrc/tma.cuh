// degnorm_b200 -- TMA 1-D bulk copies (cp.async.bulk, SASS UBLKCP), mbarrier and proxy-fence helpers shared by the
// streamed kernels (nmfoa_mid.cuh, nmfoa_wide.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

namespace {

// ---- TMA 1-D bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers --------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long *bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Waits for the phase of `bar` with this parity.  A wait that lasts seconds can only be a protocol error (a chunk that
// was never requested, a lost arrival): the kernel then stops with a message instead of hanging the device.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    if (mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    unsigned spins = 0;
    while (!mbar_try(bar, parity)) {
        if ((++spins & 1023u) == 0u && clock64() - t0 > 8000000000ll) {
            printf("degnorm_b200: mbarrier wait timed out (block %d thread %d barrier +%u parity %u)\n", (int)blockIdx.x,
                   (int)threadIdx.x, smem_u32(bar) & 0xffffu, parity);
            __trap();
        }
    }
}
// generic-proxy accesses (ordinary loads / stores, already ordered by a barrier) before async-proxy (TMA) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
// the same for shared memory only (SASS: FENCE.VIEW.ASYNC.S without the MEMBAR.GPU of the full fence)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// shared -> global bulk copy (TMA store), tracked by the issuing thread's bulk async-groups
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }
// global load that bypasses L1 (the slab's M is written by the async proxy when MID_TMA_STORE is on: an L1 line
// cached by an earlier ordinary load would be stale)
__device__ __forceinline__ void ld12cg(const double *p, double (&x)[12]) {
    const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
    for (int i = 0; i < 6; ++i) { const double2 t = __ldcg(q + i); x[2 * i] = t.x; x[2 * i + 1] = t.y; }
}

}  // namespace
