// TMA (cp.async.bulk.tensor) for the identity-indexed K-blocks of the pipelined kernels: one instruction of
// one producer lane moves a whole 128-row x 64-column bf16 tile from its [batch][rows][ld] tensor into the
// 128-byte-swizzled shared-memory tile tcgen05.mma reads, completing on the stage's mbarrier (complete_tx);
// rows / columns beyond the tensor are zero-filled by the hardware.  Gathered K-blocks (sender / receiver
// rows) keep their cp.async path: their rows are not a box of the tensor.
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace p4c {

// dims {inner columns, rows, batch}, box {64, 128, 1}, bf16, SWIZZLE_128B, zero fill.  Returns false when the
// driver entry point is missing or rejects the shape (the caller keeps the cp.async path for that block).
bool tma_rows_map(CUtensorMap* out, const void* base, long long inner, long long rows, long long ld, long long batch,
                  long long batch_stride);
// P4C_TMA=0 disables the TMA path (A/B testing)
bool tma_enabled();

namespace tc {
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;"
               :
               : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))), "r"(bytes)
               : "memory");
}
// box at (column c0, row c1, sample c2) of the tensor -> shared memory, bytes counted on `bar`
__device__ __forceinline__ void tma_load_3d(uint32_t dst_smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      :
      : "r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(static_cast<uint32_t>(__cvta_generic_to_shared(bar))),
        "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 128-byte-swizzled shared-memory tile -> box at (column c0, row c1, sample c2) of the tensor (rows / columns beyond
// the tensor are clipped by the hardware).  Bulk async-group of the ISSUING thread: commit, then wait for the source
// reads (the tile may be overwritten) or for full completion (before the CTA exits).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
}  // namespace tc
}  // namespace p4c
