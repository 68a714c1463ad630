// Inline-PTX wrappers for the CTA-pair (cta_group::2) forms of the tcgen05 / TMA / mbarrier primitives:
// two CTAs of a cluster (the two SMs of a TPC) execute one M = 256 MMA, each holding 128 rows of A and D
// and half of the N rows of B.  The even-ranked CTA (the leader) issues the MMAs and commits.
#pragma once
#include <cuda.h>
#include "tc_ptx.cuh"

namespace nplda {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {      // every thread of both CTAs
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// tensor memory: one warp of EACH CTA of the pair executes these (same warp index, same smem slot offset)
__device__ __forceinline__ void tmem_alloc2(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster (own rank allowed)
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_addr(bar)),
        "r"(cta)
        : "memory");
}

// The same with the default semantics (.release at CTA scope): no cluster-wide fence in front of the arrive.  For
// producers whose data is ordered by other means -- tensor-memory stores completed by tcgen05.wait::st +
// tcgen05.fence::before_thread_sync, shared memory made visible to the async proxy by fence.proxy.async and read by this
// CTA's own tensor core.  Measured in score_tcp.cu: the .release.cluster form costs a converter warp ~1700 cycles per
// arrive (it waits for the warp's outstanding memory traffic at cluster scope), this one ~100.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t *bar, uint32_t cta) {
    asm volatile(
        "{\n.reg .b32 ra;\nmapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n}\n" ::"r"(smem_addr(bar)),
        "r"(cta)
        : "memory");
}

// 2-D tiled TMA load into THIS CTA's shared memory whose completion bytes are posted on the LEADER's barrier
// (bit 24 of a shared::cluster address selects the odd CTA of the pair)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_addr(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar) & 0xFEFFFFFFu)
        : "memory");
}

// MMA issued by one thread of the leader CTA for the pair; D/A addresses and descriptors are CTA-relative
// (the same offsets are used in both CTAs)
__device__ __forceinline__ void mma2_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives (once all MMAs issued so far by this thread have completed) on the barrier at this offset in every
// CTA of cta_mask
__device__ __forceinline__ void mma2_commit_mc(uint64_t *bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_addr(bar)),
                 "h"(cta_mask)
                 : "memory");
}

// Shared-memory matrix descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row atoms of 1024 bytes (sbo),
// 16-byte chunk c of row r stored at chunk c ^ (r & 7) -- what TMA writes with CU_TENSOR_MAP_SWIZZLE_128B when the
// tile base is 1024-byte aligned.  A K step of 16 bf16 (32 bytes) inside the row is taken by adding 2 to the
// descriptor (start address is in 16-byte units): the hardware applies the XOR to the absolute address bits.
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                  // lbo: unused for swizzled K-major operands
    d |= (uint64_t)(1024 >> 4) << 32;        // sbo: 8 rows x 128 bytes
    d |= (uint64_t)1 << 46;                  // descriptor version 1 (Blackwell)
    d |= (uint64_t)2 << 61;                  // layout type 2 = SWIZZLE_128B
    return d;
}

// TMA row gather (tile::gather4): four rows r0..r3 of a 2-D tensor, `box` columns starting at column c0, land as four
// consecutive rows at dst (measured on B200, tools/gather4_probe.cu: the tensor map's box must be {cols, 1}; the
// swizzle follows the destination address).  .cta_group::2: the bytes are posted on the LEADER CTA's mbarrier
// (tools/gather4_pair_probe.cu).
__device__ __forceinline__ void tma_gather4_pair(void *dst, const CUtensorMap *map, int c0, int r0, int r1, int r2, int r3,
                                                 uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_addr(dst)),
        "l"(map), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_addr(bar) & 0xFEFFFFFFu)
        : "memory");
}

}  // namespace tc
}  // namespace nplda
