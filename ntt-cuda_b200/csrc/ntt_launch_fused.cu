// ntt_launch_fused.cu -- launchers of the fused transforms: contig-forward (.) key -> contig-inverse (BFV with loaded keys)
// and the fused polynomial product (full_poly_mul_device / half_poly_mul_device, poly_arithmetic.cuh:296-310).
#include "internal.h"

#include <atomic>
#include "ntt_kernels.cuh"
#include "launch_util.h"

#include <cstring>

namespace nttb200 {

// ---- fused contig-forward (.) key -> contig-inverse -------------------------------------------------------------------
template <class PF, class PI, int LOGN, int NOUT>
static int launch_fused_one(const FusedArgs &F, const CUtensorMap &mc, cudaStream_t st)
{
    constexpr size_t smem = (size_t)kContigRows * 128 * NOUT + 1024 + 16;
    static std::atomic<bool> attr_done[64];      // idempotent per-(instantiation, device) set-up: racing threads both do it
    int dev = 0;
    NTTB200_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        NTTB200_CHECK(cudaFuncSetAttribute(ntt_contig_fused_mul<PF, PI, LOGN, NOUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const unsigned tiles = ((1u << LOGN) >> 4) / kContigRows;
    NTT_LAUNCH_PDL(ntt_contig_fused_mul<PF, PI, LOGN, NOUT>, dim3(F.items * F.r * tiles), dim3(kContigRows), smem, st, mc, F);
    { const int e__ = (int)cudaGetLastError(); return e__ ? nttb200_trace_error(e__, __FILE__, __LINE__) : 0; }
}
template <class PF, class PI, int NOUT>
static int launch_fused_logn(unsigned logn, const FusedArgs &F, const CUtensorMap &mc, cudaStream_t st)
{
    switch (logn) {
    case 11: return launch_fused_one<PF, PI, 11, NOUT>(F, mc, st);
    case 12: return launch_fused_one<PF, PI, 12, NOUT>(F, mc, st);
    case 13: return launch_fused_one<PF, PI, 13, NOUT>(F, mc, st);
    case 14: return launch_fused_one<PF, PI, 14, NOUT>(F, mc, st);
    case 15: return launch_fused_one<PF, PI, 15, NOUT>(F, mc, st);
    case 16: return launch_fused_one<PF, PI, 16, NOUT>(F, mc, st);
    case 17: return launch_fused_one<PF, PI, 17, NOUT>(F, mc, st);
    default: return NTTB200_EINVAL;
    }
}
// h: data array + forward tables + group description (group = one item); lazy: both lazy policies are valid (q < 2^57)
int launch_fused_mul(bool lazy, unsigned logn, const NttArgsHost &h, const u64 *twi, const u64 *twis, const u64 *key, const u64 *key_s,
                     size_t key_item_stride, size_t key_half_stride, unsigned r, unsigned in_off, unsigned out_off0, unsigned out_off1,
                     unsigned items, int nout, cudaStream_t st)
{
    if (logn < 11 || logn > 17 || !h.a || !key || !key_s || !items || !r || !h.group_polys) return NTTB200_EINVAL;
    FusedArgs F;
    NttArgs &A = F.A;
    A.a = h.a; A.tw = h.tw; A.tws = h.tws; A.lc = h.lc;
    A.qv = nullptr; A.muv = nullptr; A.qbitv = nullptr; A.q = 0; A.mu = 0; A.qbit = 0;
    A.num = items * r; A.division = r; A.use_tma = (u32)h.use_tma; A.pf_dist = 0; A.gen_src = nullptr; A.gen_stride = 0;
    A.group_polys = h.group_polys; A.group_stride = h.group_stride;
    ntt_args_finish(A);
    F.twi = twi; F.twis = twis; F.key = key; F.key_s = key_s;
    F.key_item_stride = key_item_stride; F.key_half_stride = key_half_stride;
    F.r = r; F.in_off = in_off; F.out_off[0] = out_off0; F.out_off[1] = out_off1; F.items = items;
    CUtensorMap mc;
    if (h.use_tma & 1) {
        int rc = make_tmap_contig(&mc, A.a, logn, A.group_polys, A.group_stride, items);
        if (rc) return rc;
    } else {
        memset(&mc, 0, sizeof mc);
    }
    if (lazy) return nout == 2 ? launch_fused_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, 2>(logn, F, mc, st)
                               : launch_fused_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, 1>(logn, F, mc, st);
    return nout == 2 ? launch_fused_logn<ShoupPolicy, ShoupPolicy, 2>(logn, F, mc, st) : launch_fused_logn<ShoupPolicy, ShoupPolicy, 1>(logn, F, mc, st);
}

// ---- fused polynomial product ----------------------------------------------------------------------------------------------
template <class PF, class PI, int LOGN, bool A_FWD, bool B_FWD>
static int launch_polymul_one(const PolymulArgs &F, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo, cudaStream_t st)
{
    constexpr size_t smem = (size_t)kContigRows * 128 * 2 + 1024 + 16;
    static std::atomic<bool> attr_done[64];      // idempotent per-(instantiation, device) set-up: racing threads both do it
    int dev = 0;
    NTTB200_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        NTTB200_CHECK(cudaFuncSetAttribute(ntt_contig_polymul<PF, PI, LOGN, A_FWD, B_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const unsigned tiles = ((1u << LOGN) >> 4) / kContigRows;
    if ((size_t)F.A.num * tiles >= (1ull << 31)) return NTTB200_EINVAL;
    ntt_contig_polymul<PF, PI, LOGN, A_FWD, B_FWD><<<F.A.num * tiles, kContigRows, smem, st>>>(ma, mb, mo, F);
    { const int e__ = (int)cudaGetLastError(); return e__ ? nttb200_trace_error(e__, __FILE__, __LINE__) : 0; }
}
template <class PF, class PI, bool A_FWD, bool B_FWD>
static int launch_polymul_logn(unsigned logn, const PolymulArgs &F, const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &mo,
                               cudaStream_t st)
{
    switch (logn) {
    case 11: return launch_polymul_one<PF, PI, 11, A_FWD, B_FWD>(F, ma, mb, mo, st);
    case 12: return launch_polymul_one<PF, PI, 12, A_FWD, B_FWD>(F, ma, mb, mo, st);
    case 13: return launch_polymul_one<PF, PI, 13, A_FWD, B_FWD>(F, ma, mb, mo, st);
    case 14: return launch_polymul_one<PF, PI, 14, A_FWD, B_FWD>(F, ma, mb, mo, st);
    case 15: return launch_polymul_one<PF, PI, 15, A_FWD, B_FWD>(F, ma, mb, mo, st);
    case 16: return launch_polymul_one<PF, PI, 16, A_FWD, B_FWD>(F, ma, mb, mo, st);
    case 17: return launch_polymul_one<PF, PI, 17, A_FWD, B_FWD>(F, ma, mb, mo, st);
    default: return NTTB200_EINVAL;
    }
}
// ha: operand a (forward tables, group description); b with its own group description.  a_fwd / b_fwd: the operand still needs
// its contiguous forward pass (false: it is already in the NTT domain).  Only (true, true) and (false, false) are instantiated.
int launch_polymul(bool lazy, unsigned logn, const NttArgsHost &ha, const u64 *twi, const u64 *twis, const u64 *b, unsigned b_group_polys,
                   size_t b_group_stride, bool fwd, u64 *out, cudaStream_t st)
{
    if (logn < 11 || logn > 17 || !ha.a || !b || !ha.num || !ha.division) return NTTB200_EINVAL;
    PolymulArgs F;
    NttArgs &A = F.A;
    A.a = ha.a; A.tw = ha.tw; A.tws = ha.tws; A.lc = ha.lc;
    A.qv = nullptr; A.muv = nullptr; A.qbitv = nullptr; A.q = 0; A.mu = 0; A.qbit = 0;
    A.num = ha.num; A.division = ha.division; A.use_tma = (u32)ha.use_tma; A.pf_dist = 0; A.gen_src = nullptr; A.gen_stride = 0;
    A.group_polys = ha.group_polys ? ha.group_polys : ha.num;
    A.group_stride = ha.group_polys ? ha.group_stride : ((size_t)ha.num << logn);
    ntt_args_finish(A);
    F.b = b; F.twi = twi; F.twis = twis; F.out = out ? out : A.a;
    F.b_group_polys = b_group_polys ? b_group_polys : ha.num;
    F.b_group_stride = b_group_polys ? b_group_stride : ((size_t)ha.num << logn);
    CUtensorMap ma, mb, mo;
    if (ha.use_tma & 1) {
        int rc = make_tmap_contig(&ma, A.a, logn, A.group_polys, A.group_stride, (ha.num + A.group_polys - 1) / A.group_polys);
        if (rc) return rc;
        rc = make_tmap_contig(&mo, F.out, logn, A.group_polys, A.group_stride, (ha.num + A.group_polys - 1) / A.group_polys);
        if (rc) return rc;
        rc = make_tmap_contig(&mb, const_cast<u64 *>(b), logn, F.b_group_polys, F.b_group_stride, (ha.num + F.b_group_polys - 1) / F.b_group_polys);
        if (rc) return rc;
    } else {
        memset(&ma, 0, sizeof ma); memset(&mb, 0, sizeof mb); memset(&mo, 0, sizeof mo);
    }
    if (lazy) return fwd ? launch_polymul_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, true, true>(logn, F, ma, mb, mo, st)
                         : launch_polymul_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, false, false>(logn, F, ma, mb, mo, st);
    return fwd ? launch_polymul_logn<ShoupPolicy, ShoupPolicy, true, true>(logn, F, ma, mb, mo, st)
               : launch_polymul_logn<ShoupPolicy, ShoupPolicy, false, false>(logn, F, ma, mb, mo, st);
}

}  // namespace nttb200
