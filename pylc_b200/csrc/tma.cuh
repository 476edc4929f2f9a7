// Tensor Memory Accelerator plumbing for the sm_100a kernels: tensor-map encoding on the host (through the
// driver entry point the runtime hands out, so the library does not link libcuda) and the PTX wrappers for
// mbarrier-signalled bulk tensor loads / bulk-group tensor stores on the device.
//
// Pattern used by the staged kernels (gather_tma.cu, encode.cu, confusion.cu):
//   one elected thread arms an mbarrier with the byte count of a box (expect_tx) and issues
//   cp.async.bulk.tensor.Nd.shared::cluster.global ... -- the copy engine walks the global rows and
//   lands the box densely in shared memory; every thread waits on the barrier's phase parity and then
//   works from shared memory only.  Results go back through shared memory as well: threads write their
//   16-byte pieces, fence.proxy.async + a barrier, and one thread issues cp.async.bulk.tensor stores of
//   the whole box to each destination (bulk groups; wait_group.read frees the buffer).
#pragma once
#include <cuda.h>            // CUtensorMap and its enums (types only; no libcuda symbols are linked)
#include <cuda_runtime.h>
#include <stdint.h>

namespace pylc {

// ---- host ---------------------------------------------------------------------------------------
// Encodes a tiled tensor map of `rank` (<= 3) dimensions over 32-bit elements.  dims / box in elements
// (innermost first), strides in bytes for dimensions 1..rank-1 (multiples of 16), base 16-byte aligned.
// Returns false when the driver entry point is unavailable or the encode is rejected.
bool tma_encode_u32(CUtensorMap *map, const void *base, int rank, const uint64_t *dims, const uint64_t *strides_bytes,
                    const uint32_t *box);

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (the copy engine) before the first use
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// global -> shared box load; completion is signalled on `bar` as transaction bytes.  The box must start on a
// 16-byte boundary of global memory: c0 (in 32-bit elements) a multiple of 4 -- an unaligned start faults with
// "illegal instruction" (measured on B200); its bytes beyond the tensor are zero-filled and still counted.
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap *map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst_smem),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
                 : "memory");
}
// shared -> global box store (bulk async-group completion)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, uint32_t src_smem) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(src_smem)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap *map, int c0, int c1, int c2, uint32_t src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(src_smem)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until at most N of this thread's bulk groups still READ their shared-memory source
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (before a TMA store reads them)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ uint32_t lds32(uint32_t smem_addr) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(smem_addr) : "memory");
    return r;
}
__device__ __forceinline__ void sts32(uint32_t smem_addr, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(smem_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t smem_addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(smem_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

#endif  // __CUDACC__

}  // namespace pylc
