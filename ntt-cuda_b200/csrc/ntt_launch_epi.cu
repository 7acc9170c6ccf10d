// ntt_launch_epi.cu -- the strided inverse kernel (last kernel of the inverse transform) with a BFV encryption epilogue fused
// into its store (epi.cuh).  Lazy-policy moduli only (q < 2^57: every reference parameter set); other rings keep the separate
// epilogue kernels of bfv_kernels.cuh.
#include "internal.h"

#include <atomic>
#include "ntt_kernels.cuh"
#include "launch_util.h"

#include <cstring>

namespace nttb200 {

template <int LOGN, class EPI>
static int launch_epi_one(const NttArgs &A, const EpiArgs &E, const CUtensorMap &ms, cudaStream_t st)
{
    using SC = Sched<LOGN>;
    using P = ShoupLazyInvPolicy;
    constexpr int R = 1 << SC::K1;
    constexpr unsigned tiles_s1 = (((1u << LOGN) >> SC::K1) >> 4) / SC::NT;
    constexpr int tpc_s = tiles_per_cta(tiles_s1);
    constexpr size_t smem_s = (size_t)tpc_s * SC::NT * R * 128 + 1024 + 64;
    static std::atomic<bool> attr_done[64];      // idempotent per-(instantiation, device) set-up: racing threads both do it
    static std::atomic<int> occ[64];
    int dev = 0;
    NTTB200_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        NTTB200_CHECK(cudaFuncSetAttribute(ntt_strided_pass<P, LOGN, true, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        if (dev >= 0 && dev < 64) {
            int o_ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o_, ntt_strided_pass<P, LOGN, true, EPI>, R * SC::NT, smem_s);
            occ[dev] = o_;
            attr_done[dev] = true;
        }
    }
    NttArgs As = A;
    As.pf_dist = (dev >= 0 && dev < 64) ? pf_dist_for(dev, occ[dev]) : 0;
    const unsigned tiles_s = tiles_s1 / tpc_s;
    if ((size_t)A.num * tiles_s >= (1ull << 31)) return NTTB200_EINVAL;
    NTT_LAUNCH_PDL(ntt_strided_pass<P, LOGN, true, EPI>, dim3(A.num * tiles_s), dim3(R * SC::NT), smem_s, st, ms, As, E);
    { const int e__ = (int)cudaGetLastError(); return e__ ? nttb200_trace_error(e__, __FILE__, __LINE__) : 0; }
}
template <class EPI>
static int launch_epi_logn(unsigned logn, const NttArgs &A, const EpiArgs &E, const CUtensorMap &ms, cudaStream_t st)
{
    switch (logn) {
    case 11: return launch_epi_one<11, EPI>(A, E, ms, st);
    case 12: return launch_epi_one<12, EPI>(A, E, ms, st);
    case 13: return launch_epi_one<13, EPI>(A, E, ms, st);
    case 14: return launch_epi_one<14, EPI>(A, E, ms, st);
    case 15: return launch_epi_one<15, EPI>(A, E, ms, st);
    case 16: return launch_epi_one<16, EPI>(A, E, ms, st);
    case 17: return launch_epi_one<17, EPI>(A, E, ms, st);
    default: return NTTB200_EINVAL;
    }
}

// h: inverse tables, group description with groups = (item, half) pairs.  mode: kEpiEncLast / kEpiEncLimb.
int launch_strided_inv_epi(unsigned logn, const NttArgsHost &h, int mode, const EpiArgs &E, cudaStream_t st)
{
    if (logn < 11 || logn > 17 || !h.a || !h.tw || !h.tws || h.division == 0 || !h.group_polys) return NTTB200_EINVAL;
    if (h.num == 0) return 0;
    NttArgs A;
    A.a = h.a; A.tw = h.tw; A.tws = h.tws; A.lc = h.lc;
    A.qv = nullptr; A.muv = nullptr; A.qbitv = nullptr; A.q = 0; A.mu = 0; A.qbit = 0;
    A.num = h.num; A.division = h.division; A.use_tma = (u32)h.use_tma; A.pf_dist = 0; A.gen_src = nullptr; A.gen_stride = 0;
    A.group_polys = h.group_polys; A.group_stride = h.group_stride;
    ntt_args_finish(A);
    const unsigned groups = (h.num + A.group_polys - 1) / A.group_polys;
    CUtensorMap ms;
    if (h.use_tma & 1) {
        int r = make_tmap_strided(&ms, A.a, logn, sched_k1(logn), A.group_polys, A.group_stride, groups);
        if (r) return r;
    } else {
        memset(&ms, 0, sizeof ms);
    }
    if (mode == kEpiEncLast) return launch_epi_logn<EncLastEpi>(logn, A, E, ms, st);
    if (mode == kEpiEncLimb) return launch_epi_logn<EncLimbEpi>(logn, A, E, ms, st);
    return NTTB200_EINVAL;
}

}  // namespace nttb200
