// ntt_cluster.cuh -- ONE transform spread over a thread-block CLUSTER: the latency path of BASELINE config 1 (a single N = 4096
// transform; the reference's forwardNTT is one <<<1,1024>>> launch, ntt_60bit.cuh:314-348, its inverseNTT two, :350-386).
//
// A single small polynomial on one SM is bound by that SM's issue rate and by the latency of its dependent butterflies (ntt_single_pass:
// 32 warps on four schedulers, five CTA barriers; 17 k active cycles).  Here a cluster of CS = n / 1024 CTAs (4 at n = 4096, 2 at n = 2048)
// of 256 threads shares it; every thread owns FOUR coefficients per round = two stages in registers, with ONE exchange through
// distributed shared memory:
//
//   forward   cross round  : stages 0 .. XS-1 (strides n/2 .. 1024) on the columns {i + 1024 m}: CTA c takes columns [c GPC, (c+1) GPC)
//                            straight from global memory (coalesced), then stores element m of every column into CTA m's tile
//                            (st.shared::cluster) -- the transpose that makes the remaining ten stages local to a 1024-block
//             local rounds : five rounds of two stages inside block c; canonicalised results go to global from registers
//   inverse   the mirror image (Gentleman-Sande): five local rounds, exchange, cross round with n^-1 folded into the last stage
//
// A first version with eight coefficients per thread (cluster of 8 x 64 threads, 3 + 3 + 3 + 3 stages) left two of the four schedulers
// of every SM idle and one warp on each of the others: 7-9 k active cycles, all dependency stalls (profiles/r02_experiments.md).
//
// Twiddle indexing is the reference's: stage with m blocks reads table[m + block].  Policies: the per-butterfly-corrected ones
// (ShoupPolicy for contexts, BarrettPolicy for the stateless reference-contract path), exactly as ntt_single_pass.
// Not compiled into the CPU emulator (no cluster there): parity is pinned by the -m gpu tests against the oracle and the reference.
#pragma once
#include "ntt_kernels.cuh"

namespace nttb200 {

__device__ __forceinline__ u32 cluster_ctarank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
// address of the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ u32 dsmem_addr(const void *p, u32 rank)
{
    const u32 a = (u32)__cvta_generic_to_shared(p);
    u32 r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
}
__device__ __forceinline__ void dsmem_st(u32 addr, u64 v) { asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

constexpr unsigned kClusterNttMaxPolys = 32;       // 32 polynomials x 4 CTAs: one wave of the 148 SMs
constexpr int kClusterBlock = 1024;                // coefficients per CTA (local block)
constexpr int kClusterThreads = kClusterBlock / 4;

template <class P, int LOGN, bool INV>
__global__ void __cluster_dims__(1 << (LOGN - 10), 1, 1) __launch_bounds__(kClusterThreads) ntt_cluster_pass(NttArgs A)
{
    static_assert(!P::kLazyGS, "the latency kernels use the per-butterfly-corrected policies");
    static_assert(LOGN >= 11 && LOGN <= 12, "cluster of 2 or 4 CTAs");
    constexpr u32 n = 1u << LOGN, BL = kClusterBlock, TH = kClusterThreads;
    constexpr int XS = LOGN - 10, CS = 1 << XS;      // cross stages, cluster size
    constexpr u32 GPC = BL / CS;                     // columns per CTA in the cross round
    constexpr int GT = 4 / CS;                       // columns per thread (TH threads x GT = GPC)
    constexpr int LR = 5;                            // local rounds of two stages
    __shared__ __align__(16) u64 tile[BL];           // block c, natural order (local rounds)
    __shared__ __align__(16) u64 xt[BL];             // inverse: cross-round input, [m][column of this CTA]
    const u32 t = threadIdx.x, c = cluster_ctarank(), p = blockIdx.x / CS;
    const u32 grp = p / A.group_polys, idx = p - grp * A.group_polys;
    // Programmatic dependent launch (the launcher sets cudaLaunchAttributeProgrammaticStreamSerialization): the NEXT kernel of the stream
    // may be scheduled while this one runs, and this one may have been scheduled while its predecessor was still running -- nothing
    // before griddepcontrol.wait touches global memory; after it, everything the predecessor wrote is visible (stream order is kept).
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    cluster_arrive();                                // #1: after the matching wait every CTA of the cluster is resident (its shared memory exists)
    asm volatile("griddepcontrol.wait;" ::: "memory");
    P pol;
    pol.init(A, p % A.division, n);
    u64 *g = A.a + (size_t)grp * A.group_stride + ((size_t)idx << LOGN);
    u64 v[4];
    // The 1023 table entries of block c's ten local stages -- stage s, entries [2^s + c 2^(s-XS), + 2^(s-XS)) -- are staged in shared
    // memory NOW, as a heap (entry L = 2^(s-XS) + j), while the coefficients are on their way: loads issued inside the rounds would each
    // expose an L2 round trip (they cannot move above the barriers, and the acquire of the cluster barrier drops prefetched L1 lines).
    __shared__ __align__(16) typename P::Tw ltw[BL];
    NTT_UNROLL
    for (int k = 0; k < 4; k++) {
        const u32 L = t + TH * k;
        if (L != 0) {
            const u32 sl = 31u - (u32)__clz((int)L);
            ltw[L] = pol.load((1u << (sl + XS)) + (c << sl) + (L - (1u << sl)));
        }
    }
    if (!INV) {
        NTT_UNROLL
        for (int gt = 0; gt < GT; gt++)
            NTT_UNROLL
            for (int m = 0; m < CS; m++) v[gt * CS + m] = g[c * GPC + t + TH * gt + BL * m];
        NTT_UNROLL
        for (int s = 0; s < XS; s++) {
            const int half = CS >> (s + 1);
            NTT_UNROLL
            for (int b = 0; b < (1 << s); b++) {
                const typename P::Tw w = pol.load((1u << s) + b);
                NTT_UNROLL
                for (int gt = 0; gt < GT; gt++)
                    NTT_UNROLL
                    for (int j = 0; j < half; j++) pol.ct(v[gt * CS + b * 2 * half + j], v[gt * CS + b * 2 * half + j + half], w);
            }
        }
        cluster_wait();                              // #1
        NTT_UNROLL
        for (int gt = 0; gt < GT; gt++)
            NTT_UNROLL
            for (int m = 0; m < CS; m++) dsmem_st(dsmem_addr(&tile[c * GPC + t + TH * gt], (u32)m), v[gt * CS + m]);
        cluster_arrive();                            // #2: the stores above are visible to their owners after the wait
        cluster_wait();
        NTT_UNROLL
        for (int r = 0; r < LR; r++) {
            const u32 d = 256u >> (2 * r), hi = t / d, base = hi * 4 * d + (t % d);
            NTT_UNROLL
            for (int e = 0; e < 4; e++) v[e] = tile[base + d * e];
            const u32 L0 = (1u << (2 * r)) + hi, L1 = (2u << (2 * r)) + 2 * hi;
            const typename P::Tw w0 = ltw[L0], w1a = ltw[L1], w1b = ltw[L1 + 1];
            pol.ct(v[0], v[2], w0); pol.ct(v[1], v[3], w0);
            pol.ct(v[0], v[1], w1a); pol.ct(v[2], v[3], w1b);
            if (r < LR - 1) {
                NTT_UNROLL
                for (int e = 0; e < 4; e++) tile[base + d * e] = v[e];
                __syncthreads();
            }
        }
        NTT_UNROLL
        for (int e = 0; e < 4; e++) v[e] = pol.fwd_final(v[e]);
        ulonglong2 *o = reinterpret_cast<ulonglong2 *>(g + c * BL + 4 * t);
        o[0] = make_ulonglong2(v[0], v[1]); o[1] = make_ulonglong2(v[2], v[3]);
    } else {
        {
            const ulonglong2 *in = reinterpret_cast<const ulonglong2 *>(g + c * BL + 4 * t);
            const ulonglong2 x0 = in[0], x1 = in[1];
            v[0] = x0.x; v[1] = x0.y; v[2] = x1.x; v[3] = x1.y;
        }
        __syncthreads();                             // ltw
        NTT_UNROLL
        for (int r = LR - 1; r >= 0; r--) {
            const u32 d = 256u >> (2 * r), hi = t / d, base = hi * 4 * d + (t % d);
            if (r < LR - 1) {
                NTT_UNROLL
                for (int e = 0; e < 4; e++) v[e] = tile[base + d * e];
            }
            const u32 L0 = (1u << (2 * r)) + hi, L1 = (2u << (2 * r)) + 2 * hi;
            const typename P::Tw w0 = ltw[L0], w1a = ltw[L1], w1b = ltw[L1 + 1];
            pol.gs(v[0], v[1], w1a); pol.gs(v[2], v[3], w1b);
            pol.gs(v[0], v[2], w0); pol.gs(v[1], v[3], w0);
            if (r > 0) {
                NTT_UNROLL
                for (int e = 0; e < 4; e++) tile[base + d * e] = v[e];
                __syncthreads();
            }
        }
        // v[e] is local position t + 256 e of block c = column t + 256 e, element m = c: to the column's owner
        cluster_wait();                              // #1
        NTT_UNROLL
        for (int e = 0; e < 4; e++) {
            const u32 col = t + TH * e;
            dsmem_st(dsmem_addr(&xt[c * GPC + col % GPC], col / GPC), v[e]);
        }
        cluster_arrive();                            // #2
        cluster_wait();
        NTT_UNROLL
        for (int gt = 0; gt < GT; gt++)
            NTT_UNROLL
            for (int m = 0; m < CS; m++) v[gt * CS + m] = xt[m * GPC + t + TH * gt];
        NTT_UNROLL
        for (int s = XS - 1; s >= 0; s--) {
            const int half = CS >> (s + 1);
            NTT_UNROLL
            for (int b = 0; b < (1 << s); b++) {
                if (s == 0) {
                    NTT_UNROLL
                    for (int gt = 0; gt < GT; gt++)
                        NTT_UNROLL
                        for (int j = 0; j < half; j++) pol.gs_last(v[gt * CS + j], v[gt * CS + j + half]);
                } else {
                    const typename P::Tw w = pol.load((1u << s) + b);
                    NTT_UNROLL
                    for (int gt = 0; gt < GT; gt++)
                        NTT_UNROLL
                        for (int j = 0; j < half; j++) pol.gs(v[gt * CS + b * 2 * half + j], v[gt * CS + b * 2 * half + j + half], w);
                }
            }
        }
        NTT_UNROLL
        for (int gt = 0; gt < GT; gt++)
            NTT_UNROLL
            for (int m = 0; m < CS; m++) g[c * GPC + t + TH * gt + BL * m] = v[gt * CS + m];
    }
}

}  // namespace nttb200
