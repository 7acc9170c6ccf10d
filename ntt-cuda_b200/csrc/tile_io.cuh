// tile_io.cuh -- moving [rows][16] u64 tiles (128-byte rows) between HBM and shared memory.
//
// Product path: TMA (cp.async.bulk.tensor) + mbarrier; one elected thread issues the copy, nobody spends
// issue slots on global address arithmetic, and for the contiguous pass the 128-byte hardware swizzle
// makes both access patterns of the kernel (one 16-byte chunk per lane down a column / one whole row per
// lane) bank-conflict free.  A plain LDG.128/STS.128 path with the identical shared-memory image is kept
// for debugging (NTTB200_NO_TMA=1) and is what the CPU emulator executes.
#pragma once
#include "compat.cuh"

namespace nttb200 {

// Element offset (u64 units) of (row, col) in a [rows][16] tile.  SWZ = TMA SWIZZLE_128B image: the 16-byte
// chunk index is XORed with (row & 7)  (address bits [4,7) ^= bits [7,10); tile base is 1024-byte aligned).
template <bool SWZ>
__host__ __device__ __forceinline__ u32 tile_off(u32 row, u32 col)
{
    u32 chunk = col >> 1;
    if (SWZ) chunk ^= (row & 7u);
    return row * 16u + chunk * 2u + (col & 1u);
}

#ifndef NTTB200_EMU
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    u32 addr = smem_u32(bar), ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const void *tmap, u64 *bar, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const void *tmap, u64 *bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(void *dst, const void *tmap, u64 *bar, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                     smem_u32(dst)),
                 "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void *tmap, const void *src, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const void *tmap, const void *src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void *tmap, const void *src, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap), "r"(smem_u32(src)),
                 "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
// L2 prefetch of a tile (no shared-memory destination, no barrier): issued for the tile a LATER CTA will load, so that its
// cp.async.bulk.tensor finds the lines in the 126 MB L2 instead of waiting on HBM.
__device__ __forceinline__ void tma_prefetch_3d(const void *tmap, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_4d(const void *tmap, int c0, int c1, int c2, int c3)
{
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global [%0, {%1, %2, %3, %4}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void *tmap)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
#endif  // !NTTB200_EMU

// Cooperative 16-byte-chunk copy of `rows` tile rows; global row r starts at g + r * gstride (u64 units).
// Used by the non-TMA path and by the emulator.  All threads of the CTA must call it.
template <bool SWZ, bool TO_SMEM>
__device__ __forceinline__ void tile_copy_coop(u64 *tile, u64 *g, size_t gstride, u32 rows, u32 tid, u32 nthreads)
{
    for (u32 idx = tid; idx < rows * 8u; idx += nthreads) {
        u32 row = idx >> 3, ch = idx & 7u;
        ulonglong2 *gp = reinterpret_cast<ulonglong2 *>(g + (size_t)row * gstride + ch * 2u);
        ulonglong2 *sp = reinterpret_cast<ulonglong2 *>(tile + tile_off<SWZ>(row, ch * 2u));
        if (TO_SMEM) *sp = *gp; else *gp = *sp;
    }
}

}  // namespace nttb200
