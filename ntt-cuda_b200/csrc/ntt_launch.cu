// ntt_launch.cu -- kernel instantiations, TMA tensor maps and launch schedules for the NTT / INTT.
// Replaces the host schedulers forwardNTT / inverseNTT / *_batch (ntt_60bit.cuh:267-386, 608-697): instead of
// log2(n)-11 single-stage global passes plus one shared-memory pass, every size runs exactly two kernels.
#include "internal.h"

#include <atomic>
#include "ntt_kernels.cuh"
#include "ntt_cluster.cuh"
#include "launch_util.h"

#include <cstdlib>
#include <mutex>

namespace nttb200 {

// ---- cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency) ------------
EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// L2 prefetch distance in CTAs for a kernel with `resident` CTAs per SM (NTTB200_PF_WAVES: 0 = off; default 1 wave)
unsigned pf_dist_for(int dev, int resident)
{
    static std::atomic<int> sms[64];
    static float waves = -1.f;
    if (waves < 0.f) { const char *e = getenv("NTTB200_PF_WAVES"); waves = e ? (float)atof(e) : 1.0f; }
    if (dev < 0 || dev >= 64) return 0;
    if (!sms[dev]) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); sms[dev] = v; }
    return (unsigned)(waves * (float)(sms[dev] * resident));
}

// NTTB200_SINGLE_PASS=0 keeps the two-kernel schedule for n <= 4096 (A/B)
static bool use_single_pass()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("NTTB200_SINGLE_PASS"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

int get_tma_default()
{
    const char *e = getenv("NTTB200_NO_TMA");
    int v = (e && e[0] == '1') ? 0 : 1;
    const char *d = getenv("NTTB200_DEBUG_SKIP");        // profiling only: 1 = skip butterflies, 2 = skip tile traffic
    if (d) v |= (atoi(d) & 3) << 1;
    return v;
}

// groups x [group_polys][R = 2^K1][C = n/R] u64, box [1][1][min(R,256)][16], no swizzle
int make_tmap_strided(CUtensorMap *m, u64 *a, unsigned logn, unsigned k1, unsigned group_polys, size_t group_stride, unsigned groups)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return NTTB200_ENOTMA;
    const cuuint64_t n = 1ull << logn, R = 1ull << k1, C = n >> k1;
    cuuint64_t dims[4] = {C, R, group_polys, groups};
    cuuint64_t strides[3] = {C * 8, n * 8, (cuuint64_t)group_stride * 8};
    cuuint32_t box[4] = {16, (cuuint32_t)(R > 256 ? 256 : R), 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : NTTB200_ENOTMA;
}
// groups x [rows = group_polys*n/16][16] u64, box [1][128][16], 128-byte swizzle
int make_tmap_contig(CUtensorMap *m, u64 *a, unsigned logn, unsigned group_polys, size_t group_stride, unsigned groups)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return NTTB200_ENOTMA;
    cuuint64_t dims[3] = {16, ((cuuint64_t)group_polys << logn) >> 4, groups};
    cuuint64_t strides[2] = {128, (cuuint64_t)group_stride * 8};
    cuuint32_t box[3] = {16, (cuuint32_t)kContigRows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, a, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : NTTB200_ENOTMA;
}

template <class P, int LOGN, bool INV>
static int launch_one(const NttArgs &A, int which, unsigned cnt, const CUtensorMap &ms, const CUtensorMap &mc, cudaStream_t st)
{
    using SC = Sched<LOGN>;
    constexpr int R = 1 << SC::K1;
    constexpr unsigned tiles_s1 = (((1u << LOGN) >> SC::K1) >> 4) / SC::NT, tiles_c1 = ((1u << LOGN) >> 4) / kContigRows;
    constexpr int tpc_s = tiles_per_cta(tiles_s1), tpc_c = tiles_per_cta(tiles_c1);
    constexpr size_t smem_s = (size_t)tpc_s * SC::NT * R * 128 + 1024 + 64;
    constexpr size_t smem_c = (size_t)tpc_c * kContigRows * 128 + 1024 + 64;
    // the dynamic shared-memory opt-in is a per-device function attribute: remember it per (instantiation, device)
    static std::atomic<bool> attr_done[64];      // idempotent per-(instantiation, device) set-up: racing threads both do it
    int dev = 0;
    NTTB200_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        NTTB200_CHECK(cudaFuncSetAttribute(ntt_strided_pass<P, LOGN, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
        NTTB200_CHECK(cudaFuncSetAttribute(ntt_contig_pass<P, LOGN, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_c));
        if (const char *e = getenv("NTTB200_CARVEOUT")) {      // tuning hook: shared-memory carve-out in percent (rest of the 228 KB is L1)
            cudaFuncSetAttribute(ntt_strided_pass<P, LOGN, INV>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
            cudaFuncSetAttribute(ntt_contig_pass<P, LOGN, INV>, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e));
        }
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    const unsigned tiles_s = tiles_s1 / tpc_s, tiles_c = tiles_c1 / tpc_c;       // CTAs per polynomial
    if ((size_t)cnt * tiles_c >= (1ull << 31)) return NTTB200_EINVAL;
    const dim3 gs(cnt * tiles_s);
    const dim3 gc(cnt * tiles_c);
    // which: -1 = whole transform, 0 / 1 = only the first / second kernel in execution order (profiling hook)
    const bool do_strided = which < 0 || (which == 0) == !INV;
    const bool do_contig = which < 0 || (which == 1) == !INV;
    static std::atomic<int> occ_s[64], occ_c[64];      // resident CTAs per SM of the two kernels (prefetch distance = one wave)
    if (dev >= 0 && dev < 64 && !occ_s[dev]) {
        int os_ = 0, oc_ = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&os_, ntt_strided_pass<P, LOGN, INV>, R * SC::NT, smem_s);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oc_, ntt_contig_pass<P, LOGN, INV>, kContigRows, smem_c);
        occ_s[dev] = os_; occ_c[dev] = oc_;
    }
    NttArgs As = A, Ac = A;
    As.pf_dist = (dev >= 0 && dev < 64) ? pf_dist_for(dev, occ_s[dev]) : 0;
    Ac.pf_dist = (dev >= 0 && dev < 64) ? pf_dist_for(dev, occ_c[dev]) : 0;
    if (!INV) {
        if (do_strided) NTT_LAUNCH_PDL(ntt_strided_pass<P, LOGN, INV>, gs, dim3(R * SC::NT), smem_s, st, ms, As, EpiArgs{});
        if (do_contig) NTT_LAUNCH_PDL(ntt_contig_pass<P, LOGN, INV>, gc, dim3(kContigRows), smem_c, st, mc, Ac);
    } else {
        if (do_contig) NTT_LAUNCH_PDL(ntt_contig_pass<P, LOGN, INV>, gc, dim3(kContigRows), smem_c, st, mc, Ac);
        if (do_strided) NTT_LAUNCH_PDL(ntt_strided_pass<P, LOGN, INV>, gs, dim3(R * SC::NT), smem_s, st, ms, As, EpiArgs{});
    }
    { const int e__ = (int)cudaGetLastError(); return e__ ? nttb200_trace_error(e__, __FILE__, __LINE__) : 0; }
}

// n <= 4096, few polynomials: the whole transform in one kernel, one CTA of n/4 threads per polynomial (ntt_single_pass)
template <class P, int LOGN, bool INV>
static int launch_single(const NttArgs &A, cudaStream_t st)
{
    NTT_LAUNCH_PDL(ntt_single_pass<P, LOGN, INV>, dim3(A.num), dim3(1 << (LOGN - 2)), (size_t)8 << LOGN, st, A);
    { const int e__ = (int)cudaGetLastError(); return e__ ? nttb200_trace_error(e__, __FILE__, __LINE__) : 0; }
}
// NTTB200_CLUSTER_NTT=0 keeps the one-CTA latency kernel (A/B)
static bool use_cluster_ntt()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("NTTB200_CLUSTER_NTT"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
// n <= 4096, at most kClusterNttMaxPolys polynomials: one transform per cluster of n / 1024 CTAs (ntt_cluster_pass)
template <class P, int LOGN, bool INV>
static int launch_cluster(const NttArgs &A, cudaStream_t st)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(A.num << (LOGN - 10));
    cfg.blockDim = dim3(kClusterThreads);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // see the kernel: griddepcontrol.wait precedes its first global access
    at[0].val.programmaticStreamSerializationAllowed = pdl_allow(st, (unsigned long long)A.num << (LOGN - 10)) ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    NTTB200_CHECK(cudaLaunchKernelEx(&cfg, ntt_cluster_pass<P, LOGN, INV>, A));
    { const int e__ = (int)cudaGetLastError(); return e__ ? nttb200_trace_error(e__, __FILE__, __LINE__) : 0; }
}
template <class P, bool INV>
static int launch_single_logn(unsigned logn, const NttArgs &A, cudaStream_t st)
{
    if (A.num <= kClusterNttMaxPolys && use_cluster_ntt())
        return logn == 11 ? launch_cluster<P, 11, INV>(A, st) : launch_cluster<P, 12, INV>(A, st);
    return logn == 11 ? launch_single<P, 11, INV>(A, st) : launch_single<P, 12, INV>(A, st);
}

template <class P, bool INV>
static int launch_logn(unsigned logn, const NttArgs &A, int p0, unsigned cnt, const CUtensorMap &ms, const CUtensorMap &mc,
                       cudaStream_t st)
{
    switch (logn) {
    case 11: return launch_one<P, 11, INV>(A, p0, cnt, ms, mc, st);
    case 12: return launch_one<P, 12, INV>(A, p0, cnt, ms, mc, st);
    case 13: return launch_one<P, 13, INV>(A, p0, cnt, ms, mc, st);
    case 14: return launch_one<P, 14, INV>(A, p0, cnt, ms, mc, st);
    case 15: return launch_one<P, 15, INV>(A, p0, cnt, ms, mc, st);
    case 16: return launch_one<P, 16, INV>(A, p0, cnt, ms, mc, st);
    case 17: return launch_one<P, 17, INV>(A, p0, cnt, ms, mc, st);
    default: return NTTB200_EINVAL;
    }
}

unsigned sched_k1(unsigned logn)
{
    switch (logn) {
    case 11: return Sched<11>::K1; case 12: return Sched<12>::K1; case 13: return Sched<13>::K1; case 14: return Sched<14>::K1;
    case 15: return Sched<15>::K1; case 16: return Sched<16>::K1; case 17: return Sched<17>::K1;
    default: return 0;
    }
}

int launch_ntt(bool inverse, int policy, unsigned logn, const NttArgsHost &h, cudaStream_t st) { return launch_ntt_pass(inverse, policy, logn, h, -1, st); }

int launch_ntt_pass(bool inverse, int policy, unsigned logn, const NttArgsHost &h, int which, cudaStream_t st)
{
    if (logn < 11 || logn > 17 || !h.a || !h.tw || h.division == 0) return NTTB200_EINVAL;
    if (h.num == 0) return 0;
    if (h.division > h.num && h.num) { /* fewer polynomials than limbs: classes beyond num are simply empty */ }
    NttArgs A;
    A.a = h.a;
    A.tw = h.tw; A.tws = h.tws; A.lc = h.lc;
    A.qv = h.qv; A.muv = h.muv; A.qbitv = h.qbitv;
    A.q = h.q; A.mu = h.mu; A.qbit = h.qbit;
    A.num = h.num; A.division = h.division; A.use_tma = (u32)h.use_tma; A.pf_dist = 0;
    A.gen_src = h.gen_src; A.gen_stride = h.gen_stride;
    A.group_polys = h.group_polys ? h.group_polys : h.num;
    A.group_stride = h.group_polys ? h.group_stride : ((size_t)h.num << logn);
    ntt_args_finish(A);
    const unsigned groups = (h.num + A.group_polys - 1) / A.group_polys;
    if (which < 0 && logn <= 12 && h.num <= kSmallNttMaxPolys && !h.gen_src && !(h.use_tma & 6) && use_single_pass()) {   // latency path, no tensor maps
        // (the lazy policies track bounds over 16-coefficient rounds; here every butterfly is corrected: general Shoup policy)
        if (policy != kPolicyBarrett) return inverse ? launch_single_logn<ShoupPolicy, true>(logn, A, st) : launch_single_logn<ShoupPolicy, false>(logn, A, st);
        return inverse ? launch_single_logn<BarrettPolicy, true>(logn, A, st) : launch_single_logn<BarrettPolicy, false>(logn, A, st);
    }
    CUtensorMap ms, mc;
    if (h.use_tma & 1) {
        int r = make_tmap_strided(&ms, A.a, logn, sched_k1(logn), A.group_polys, A.group_stride, groups);
        if (r) return r;
        r = make_tmap_contig(&mc, A.a, logn, A.group_polys, A.group_stride, groups);
        if (r) return r;
    } else {
        memset(&ms, 0, sizeof ms); memset(&mc, 0, sizeof mc);
    }
    const unsigned cnt = h.num;
    int r;
    if (policy == kPolicyShoupLazy && !inverse) r = launch_logn<ShoupLazyPolicy, false>(logn, A, which, cnt, ms, mc, st);
    else if (policy == kPolicyShoupLazy && inverse) r = launch_logn<ShoupLazyInvPolicy, true>(logn, A, which, cnt, ms, mc, st);
    else if (policy != kPolicyBarrett) r = inverse ? launch_logn<ShoupPolicy, true>(logn, A, which, cnt, ms, mc, st) : launch_logn<ShoupPolicy, false>(logn, A, which, cnt, ms, mc, st);
    else r = inverse ? launch_logn<BarrettPolicy, true>(logn, A, which, cnt, ms, mc, st) : launch_logn<BarrettPolicy, false>(logn, A, which, cnt, ms, mc, st);
    if (r) return r;
    return 0;
}

}  // namespace nttb200

using namespace nttb200;

static unsigned ilog2u(unsigned n) { unsigned l = 0; while ((1u << l) < n) l++; return l; }

extern "C" {

int nttb200_forward_ntt_batch(const nttb200_ctx *ctx, nttb200_u64 *a, unsigned num, unsigned division, void *stream)
{
    if (!ctx || division == 0 || division > ctx->limbs) return NTTB200_EINVAL;
    NttArgsHost h{a, ctx->psi, ctx->psi_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, ctx->use_tma};
    return launch_ntt(false, ctx->lazy_ok ? kPolicyShoupLazy : kPolicyShoup, ctx->logn, h, (cudaStream_t)stream);
}
int nttb200_inverse_ntt_batch(const nttb200_ctx *ctx, nttb200_u64 *a, unsigned num, unsigned division, void *stream)
{
    if (!ctx || division == 0 || division > ctx->limbs) return NTTB200_EINVAL;
    NttArgsHost h{a, ctx->psiinv, ctx->psiinv_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, ctx->use_tma};
    return launch_ntt(true, ctx->lazy_ok ? kPolicyShoupLazy : kPolicyShoup, ctx->logn, h, (cudaStream_t)stream);
}

int nttb200_ntt_pass(const nttb200_ctx *ctx, nttb200_u64 *a, unsigned num, unsigned division, int inverse, int which, void *stream)
{
    if (!ctx || division == 0 || division > ctx->limbs || which < 0 || which > 1) return NTTB200_EINVAL;
    NttArgsHost h{a, inverse ? ctx->psiinv : ctx->psi, inverse ? ctx->psiinv_s : ctx->psi_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0,
                  num, division, ctx->use_tma};
    return launch_ntt_pass(inverse != 0, ctx->lazy_ok ? kPolicyShoupLazy : kPolicyShoup, ctx->logn, h, which, (cudaStream_t)stream);
}

// a <- a * b in Z_q[X]/(X^n + 1), polynomial p modulo limb p % division; b is clobbered (it holds its half-transformed image).
// 4 launches: strided forward pass of a and of b, ONE fused kernel (both contiguous forward passes, coefficient-wise product,
// contiguous inverse pass), strided inverse pass.
int nttb200_poly_mul_batch(const nttb200_ctx *ctx, nttb200_u64 *a, nttb200_u64 *b, unsigned num, unsigned division, void *stream)
{
    if (!ctx || !a || !b || division == 0 || division > ctx->limbs) return NTTB200_EINVAL;
    if (num == 0) return 0;
    const int pol = ctx->lazy_ok ? kPolicyShoupLazy : kPolicyShoup;
    cudaStream_t st = (cudaStream_t)stream;
    NttArgsHost ha{a, ctx->psi, ctx->psi_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, ctx->use_tma};
    NttArgsHost hb = ha; hb.a = b;
    int r = launch_ntt_pass(false, pol, ctx->logn, ha, 0, st);
    if (!r) r = launch_ntt_pass(false, pol, ctx->logn, hb, 0, st);
    if (!r) r = launch_polymul(ctx->lazy_ok != 0, ctx->logn, ha, ctx->psiinv, ctx->psiinv_s, b, 0, 0, true, nullptr, st);
    NttArgsHost hi{a, ctx->psiinv, ctx->psiinv_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, ctx->use_tma};
    if (!r) r = launch_ntt_pass(true, pol, ctx->logn, hi, 1, st);
    return r;
}
// a <- INTT(a (.) b) for operands that are both already in the NTT domain (canonical): coefficient-wise product fused into the
// contiguous inverse pass, then the strided inverse pass.  b is only read.
int nttb200_ntt_domain_mul_inverse_batch(const nttb200_ctx *ctx, nttb200_u64 *a, const nttb200_u64 *b, unsigned num, unsigned division,
                                         void *stream)
{
    if (!ctx || !a || !b || division == 0 || division > ctx->limbs) return NTTB200_EINVAL;
    if (num == 0) return 0;
    const int pol = ctx->lazy_ok ? kPolicyShoupLazy : kPolicyShoup;
    cudaStream_t st = (cudaStream_t)stream;
    NttArgsHost ha{a, ctx->psi, ctx->psi_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, ctx->use_tma};
    int r = launch_polymul(ctx->lazy_ok != 0, ctx->logn, ha, ctx->psiinv, ctx->psiinv_s, b, 0, 0, false, nullptr, st);
    NttArgsHost hi{a, ctx->psiinv, ctx->psiinv_s, ctx->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, ctx->use_tma};
    if (!r) r = launch_ntt_pass(true, pol, ctx->logn, hi, 1, st);
    return r;
}

int nttb200_ref_forward_ntt_batch(nttb200_u64 *a, unsigned n, const nttb200_u64 *psi_powers, unsigned num, unsigned division,
                                  const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream)
{
    if (!q_dev || !mu_dev || !qbit_dev || (n & (n - 1))) return NTTB200_EINVAL;
    NttArgsHost h{a, psi_powers, nullptr, nullptr, q_dev, mu_dev, qbit_dev, 0, 0, 0, num, division, get_tma_default()};
    return launch_ntt(false, kPolicyBarrett, ilog2u(n), h, (cudaStream_t)stream);
}
int nttb200_ref_inverse_ntt_batch(nttb200_u64 *a, unsigned n, const nttb200_u64 *psiinv_powers, unsigned num, unsigned division,
                                  const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream)
{
    if (!q_dev || !mu_dev || !qbit_dev || (n & (n - 1))) return NTTB200_EINVAL;
    NttArgsHost h{a, psiinv_powers, nullptr, nullptr, q_dev, mu_dev, qbit_dev, 0, 0, 0, num, division, get_tma_default()};
    return launch_ntt(true, kPolicyBarrett, ilog2u(n), h, (cudaStream_t)stream);
}
int nttb200_ref_forward_ntt(nttb200_u64 *a, unsigned n, void *stream, nttb200_u64 q, nttb200_u64 mu, int qbit, const nttb200_u64 *psi_powers)
{
    if (n & (n - 1)) return NTTB200_EINVAL;
    NttArgsHost h{a, psi_powers, nullptr, nullptr, nullptr, nullptr, nullptr, q, mu, (unsigned)qbit, 1, 1, get_tma_default()};
    return launch_ntt(false, kPolicyBarrett, ilog2u(n), h, (cudaStream_t)stream);
}
int nttb200_ref_inverse_ntt(nttb200_u64 *a, unsigned n, void *stream, nttb200_u64 q, nttb200_u64 mu, int qbit, const nttb200_u64 *psiinv_powers)
{
    if (n & (n - 1)) return NTTB200_EINVAL;
    NttArgsHost h{a, psiinv_powers, nullptr, nullptr, nullptr, nullptr, nullptr, q, mu, (unsigned)qbit, 1, 1, get_tma_default()};
    return launch_ntt(true, kPolicyBarrett, ilog2u(n), h, (cudaStream_t)stream);
}

}  // extern "C"
