// emu.cpp -- CPU execution of the CUDA kernel sources (compiled with -DNTTB200_EMU).  TEST INFRASTRUCTURE.
#include "emu_rt.h"

#include "../ntt_kernels.cuh"

thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
thread_local unsigned char *emu_dyn_smem = nullptr;
thread_local EmuCta *emu_cta = nullptr;

void __syncthreads() { emu_cta->block_bar->arrive_and_wait(); }
void __syncwarp() { emu_cta->warp_bars[threadIdx.x / 32]->arrive_and_wait(); }

float emu_normcdfinvf(float x)
{
    // Beasley-Springer-Moro style via bisection on erfc: slow but only used on tiny emulator inputs
    double lo = -40, hi = 40;
    for (int i = 0; i < 200; i++) {
        double mid = 0.5 * (lo + hi);
        if (0.5 * std::erfc(-mid / std::sqrt(2.0)) < (double)x) lo = mid; else hi = mid;
    }
    return (float)(0.5 * (lo + hi));
}

// ---- functional TMA model ------------------------------------------------------------------------------------------
static void emu_tma(bool load, const TensorMap *m, void *smem, const int *c)
{
    const EmuTmapDesc *d = reinterpret_cast<const EmuTmapDesc *>(m->opaque);
    const size_t es = d->strides[0];
    uint32_t b0 = d->box[0], b1 = d->rank > 1 ? d->box[1] : 1, b2 = d->rank > 2 ? d->box[2] : 1, b3 = d->rank > 3 ? d->box[3] : 1;
    unsigned char *s = (unsigned char *)smem;
    if (d->swizzle128 && ((uintptr_t)s & 1023)) abort();   // hardware requirement
    size_t lin = 0;
    for (uint32_t m3 = 0; m3 < b3; m3++)
    for (uint32_t k = 0; k < b2; k++)
        for (uint32_t j = 0; j < b1; j++)
            for (uint32_t i = 0; i < b0; i++, lin += es) {
                size_t goff = (size_t)(c[0] + i) * d->strides[0];
                if (d->rank > 1) goff += (size_t)(c[1] + j) * d->strides[1];
                if (d->rank > 2) goff += (size_t)(c[2] + k) * d->strides[2];
                if (d->rank > 3) goff += (size_t)(c[3] + m3) * d->strides[3];
                for (int dd = 0; dd < d->rank; dd++) {   // out-of-bounds boxes would be a bug in the kernels
                    uint64_t cc = (uint64_t)c[dd] + (dd == 0 ? i : dd == 1 ? j : dd == 2 ? k : m3);
                    if (cc >= d->dims[dd]) abort();
                }
                size_t soff = lin;
                if (d->swizzle128) soff ^= ((soff >> 7) & 7) << 4;
                if (load) memcpy(s + soff, d->base + goff, es); else memcpy(d->base + goff, s + soff, es);
            }
}
void emu_tma_2d(bool load, const TensorMap *m, void *smem, int c0, int c1) { int c[4] = {c0, c1, 0, 0}; emu_tma(load, m, smem, c); }
void emu_tma_3d(bool load, const TensorMap *m, void *smem, int c0, int c1, int c2) { int c[4] = {c0, c1, c2, 0}; emu_tma(load, m, smem, c); }
void emu_tma_4d(bool load, const TensorMap *m, void *smem, int c0, int c1, int c2, int c3) { int c[4] = {c0, c1, c2, c3}; emu_tma(load, m, smem, c); }

using namespace nttb200;

static int g_which = -1;   // -1: whole transform, 0 / 1: only the first / second kernel in execution order
static const unsigned char *g_gen_src = nullptr;   // next forward strided pass generates its input (NttArgs::gen_src)
static size_t g_gen_stride = 0;

static int g_single_pass = 1;          // n <= 4096: whole transform in one kernel, as csrc/ntt_launch.cu dispatches it
extern "C" __attribute__((visibility("default"))) void emu_set_single_pass(int on) { g_single_pass = on; }

template <class P, int LOGN, bool INV>
static void run_one(const NttArgs &A)
{
    if constexpr (LOGN <= 12 && !P::kLazyGS) {
        if (g_which < 0 && !A.gen_src && g_single_pass && A.num <= kSmallNttMaxPolys) {
            emu_dim3 g1;
            g1.x = A.num;
            emu_launch(g1, 1 << (LOGN - 2), (size_t)8 << LOGN, [&] { ntt_single_pass<P, LOGN, INV>(A); });
            return;
        }
    }
    using SC = Sched<LOGN>;
    constexpr int R = 1 << SC::K1;
    const size_t n = (size_t)1 << LOGN;
    const unsigned groups = (A.num + A.group_polys - 1) / A.group_polys;
    TensorMap ms, mc;
    EmuTmapDesc ds{}, dc{};
    ds.base = (unsigned char *)A.a; ds.rank = 4;
    ds.dims[0] = n >> SC::K1; ds.dims[1] = R; ds.dims[2] = A.group_polys; ds.dims[3] = groups;
    ds.strides[0] = 8; ds.strides[1] = (n >> SC::K1) * 8; ds.strides[2] = n * 8; ds.strides[3] = A.group_stride * 8;
    ds.box[0] = 16; ds.box[1] = R > 256 ? 256 : R; ds.box[2] = 1; ds.box[3] = 1; ds.swizzle128 = 0;
    dc.base = (unsigned char *)A.a; dc.rank = 3;
    dc.dims[0] = 16; dc.dims[1] = ((size_t)A.group_polys << LOGN) >> 4; dc.dims[2] = groups;
    dc.strides[0] = 8; dc.strides[1] = 128; dc.strides[2] = A.group_stride * 8;
    dc.box[0] = 16; dc.box[1] = kContigRows; dc.box[2] = 1; dc.swizzle128 = 1;
    static_assert(sizeof(EmuTmapDesc) <= sizeof(TensorMap), "descriptor stub too small");
    memcpy(ms.opaque, &ds, sizeof ds);
    memcpy(mc.opaque, &dc, sizeof dc);
    const unsigned tiles_s1 = (unsigned)(((n >> SC::K1) >> 4) / SC::NT), tiles_c1 = (unsigned)((n >> 4) / kContigRows);
    const int tpc_s = tiles_per_cta(tiles_s1), tpc_c = tiles_per_cta(tiles_c1);
    emu_dim3 gs, gc;
    gs.x = A.num * (tiles_s1 / tpc_s);
    gc.x = A.num * (tiles_c1 / tpc_c);
    const size_t smem_s = (size_t)tpc_s * SC::NT * R * 128 + 1024 + 64, smem_c = (size_t)tpc_c * kContigRows * 128 + 1024 + 64;
    auto strided = [&] { emu_launch(gs, R * SC::NT, smem_s, [&] { ntt_strided_pass<P, LOGN, INV>(ms, A, EpiArgs{}); }); };
    auto contig = [&] { emu_launch(gc, kContigRows, smem_c, [&] { ntt_contig_pass<P, LOGN, INV>(mc, A); }); };
    if (!INV) { if (g_which != 1) strided(); if (g_which != 0) contig(); } else { if (g_which != 1) contig(); if (g_which != 0) strided(); }
}

template <class PF, class PI, int LOGN, int NOUT>
static void run_fused_one(const FusedArgs &F)
{
    const NttArgs &A = F.A;
    TensorMap mc;
    EmuTmapDesc dc{};
    dc.base = (unsigned char *)A.a; dc.rank = 3;
    dc.dims[0] = 16; dc.dims[1] = ((size_t)A.group_polys << LOGN) >> 4; dc.dims[2] = F.items;
    dc.strides[0] = 8; dc.strides[1] = 128; dc.strides[2] = A.group_stride * 8;
    dc.box[0] = 16; dc.box[1] = kContigRows; dc.box[2] = 1; dc.swizzle128 = 1;
    memcpy(mc.opaque, &dc, sizeof dc);
    emu_dim3 g;
    g.x = F.items * F.r * (unsigned)((((size_t)1 << LOGN) >> 4) / kContigRows);
    emu_launch(g, kContigRows, (size_t)kContigRows * 128 * NOUT + 1024 + 16, [&] { ntt_contig_fused_mul<PF, PI, LOGN, NOUT>(mc, F); });
}
template <class PF, class PI, int NOUT>
static int run_fused_logn(int logn, const FusedArgs &F)
{
    switch (logn) {
    case 11: run_fused_one<PF, PI, 11, NOUT>(F); return 0;
    case 12: run_fused_one<PF, PI, 12, NOUT>(F); return 0;
    case 13: run_fused_one<PF, PI, 13, NOUT>(F); return 0;
    case 14: run_fused_one<PF, PI, 14, NOUT>(F); return 0;
    case 15: run_fused_one<PF, PI, 15, NOUT>(F); return 0;
    case 16: run_fused_one<PF, PI, 16, NOUT>(F); return 0;
    case 17: run_fused_one<PF, PI, 17, NOUT>(F); return 0;
    }
    return 1;
}

template <class P, bool INV>
static int run_logn(int logn, const NttArgs &A)
{
    switch (logn) {
    case 11: run_one<P, 11, INV>(A); return 0;
    case 12: run_one<P, 12, INV>(A); return 0;
    case 13: run_one<P, 13, INV>(A); return 0;
    case 14: run_one<P, 14, INV>(A); return 0;
    case 15: run_one<P, 15, INV>(A); return 0;
    case 16: run_one<P, 16, INV>(A); return 0;
    case 17: run_one<P, 17, INV>(A); return 0;
    }
    return 1;
}

extern "C" __attribute__((visibility("default")))
int emu_ntt(int inverse, int barrett, int use_tma, int logn, u64 *a, const u64 *tw, const u64 *tws, const LimbConst *lc,
            const u64 *qv, const u64 *muv, const u32 *qbitv, unsigned num, unsigned division, unsigned group_polys, size_t group_stride)
{
    NttArgs A{};
    A.group_polys = group_polys ? group_polys : num;
    A.group_stride = group_polys ? group_stride : ((size_t)num << logn);
    A.a = a; A.tw = tw; A.tws = tws; A.lc = lc; A.qv = qv; A.muv = muv; A.qbitv = qbitv;
    A.num = num; A.division = division; A.use_tma = (u32)use_tma;
    A.gen_src = inverse ? nullptr : g_gen_src; A.gen_stride = g_gen_stride;
    ntt_args_finish(A);
    // the library sends small rings with few polynomials to the latency kernel under the general Shoup policy (csrc/ntt_launch.cu)
    if (barrett == 2 && logn <= 12 && num <= kSmallNttMaxPolys && g_single_pass && g_which < 0 && !A.gen_src) barrett = 0;
    if (barrett == 2 && !inverse) return run_logn<ShoupLazyPolicy, false>(logn, A);
    if (barrett == 2 && inverse) return run_logn<ShoupLazyInvPolicy, true>(logn, A);
    if (barrett != 1) return inverse ? run_logn<ShoupPolicy, true>(logn, A) : run_logn<ShoupPolicy, false>(logn, A);
    return inverse ? run_logn<BarrettPolicy, true>(logn, A) : run_logn<BarrettPolicy, false>(logn, A);
}

template <class PF, class PI, int LOGN, bool FWD>
static void run_polymul_one(const PolymulArgs &F)
{
    const NttArgs &A = F.A;
    TensorMap ma, mb, mo;
    EmuTmapDesc da{}, db{}, dout{};
    da.base = (unsigned char *)A.a; da.rank = 3;
    da.dims[0] = 16; da.dims[1] = ((size_t)A.group_polys << LOGN) >> 4; da.dims[2] = (A.num + A.group_polys - 1) / A.group_polys;
    da.strides[0] = 8; da.strides[1] = 128; da.strides[2] = A.group_stride * 8;
    da.box[0] = 16; da.box[1] = kContigRows; da.box[2] = 1; da.swizzle128 = 1;
    db = da;
    db.base = (unsigned char *)F.b; db.dims[1] = ((size_t)F.b_group_polys << LOGN) >> 4; db.dims[2] = (A.num + F.b_group_polys - 1) / F.b_group_polys;
    db.strides[2] = F.b_group_stride * 8;
    dout = da; dout.base = (unsigned char *)F.out;
    memcpy(ma.opaque, &da, sizeof da);
    memcpy(mb.opaque, &db, sizeof db);
    memcpy(mo.opaque, &dout, sizeof dout);
    emu_dim3 g;
    g.x = A.num * (unsigned)((((size_t)1 << LOGN) >> 4) / kContigRows);
    emu_launch(g, kContigRows, (size_t)kContigRows * 128 * 2 + 1024 + 16, [&] { ntt_contig_polymul<PF, PI, LOGN, FWD, FWD>(ma, mb, mo, F); });
}
template <class PF, class PI, bool FWD>
static int run_polymul_logn(int logn, const PolymulArgs &F)
{
    switch (logn) {
    case 11: run_polymul_one<PF, PI, 11, FWD>(F); return 0;
    case 12: run_polymul_one<PF, PI, 12, FWD>(F); return 0;
    case 13: run_polymul_one<PF, PI, 13, FWD>(F); return 0;
    case 14: run_polymul_one<PF, PI, 14, FWD>(F); return 0;
    case 15: run_polymul_one<PF, PI, 15, FWD>(F); return 0;
    case 16: run_polymul_one<PF, PI, 16, FWD>(F); return 0;
    case 17: run_polymul_one<PF, PI, 17, FWD>(F); return 0;
    }
    return 1;
}
// nttb200_poly_mul_batch (fwd = 1) / nttb200_ntt_domain_mul_inverse_batch (fwd = 0) on the emulator, same launch order
extern "C" __attribute__((visibility("default")))
int emu_polymul(int fwd, int lazy, int logn, u64 *a, u64 *b, const u64 *psi, const u64 *psi_s, const u64 *psiinv, const u64 *psiinv_s,
                const LimbConst *lc, unsigned num, unsigned division)
{
    const int pol = lazy ? 2 : 0;
    if (fwd) {
        g_which = 0;
        emu_ntt(0, pol, 1, logn, a, psi, psi_s, lc, nullptr, nullptr, nullptr, num, division, 0, 0);
        emu_ntt(0, pol, 1, logn, b, psi, psi_s, lc, nullptr, nullptr, nullptr, num, division, 0, 0);
        g_which = -1;
    }
    PolymulArgs F{};
    F.A.a = a; F.A.tw = psi; F.A.tws = psi_s; F.A.lc = lc; F.A.num = num; F.A.division = division; F.A.use_tma = 1;
    F.A.group_polys = num; F.A.group_stride = (size_t)num << logn;
    F.out = a; F.b = b; F.twi = psiinv; F.twis = psiinv_s; F.b_group_polys = num; F.b_group_stride = (size_t)num << logn;
    int r;
    if (lazy) r = fwd ? run_polymul_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, true>(logn, F) : run_polymul_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, false>(logn, F);
    else r = fwd ? run_polymul_logn<ShoupPolicy, ShoupPolicy, true>(logn, F) : run_polymul_logn<ShoupPolicy, ShoupPolicy, false>(logn, F);
    if (r) return r;
    g_which = 1;
    r = emu_ntt(1, pol, 1, logn, a, psiinv, psiinv_s, lc, nullptr, nullptr, nullptr, num, division, 0, 0);
    g_which = -1;
    return r;
}

extern "C" __attribute__((visibility("default"))) unsigned emu_sizeof_limbconst() { return (unsigned)sizeof(LimbConst); }

// ---- BFV pipelines on the emulator: same kernels, same order as csrc/bfv.cu ---------------------------------------------
#include "../bfv_kernels.cuh"

namespace {
struct EmuRing {
    unsigned n, logn, r;
    const u64 *q, *mu; const u32 *qbit;
    const u64 *psi, *psiinv, *psi_s, *psiinv_s; const LimbConst *lc;
    int barrett;    // 1: stateless Barrett NTT, 0: Shoup (lazy forward)
};
template <class F> void ew(F &&f) { emu_dim3 g; g.x = 3; emu_launch(g, 64, 0, f); }
template <class F> void ew3(unsigned y, unsigned z, F &&f) { emu_dim3 g; g.x = 2; g.y = y; g.z = z; emu_launch(g, 64, 0, f); }
int ring_ntt(const EmuRing &R, bool inv, u64 *a, unsigned num, unsigned division, unsigned gp, size_t gs, int which = -1)
{
    struct Restore { ~Restore() { g_which = -1; } } restore;
    g_which = which;
    return emu_ntt(inv, R.barrett < 0 ? 0 : R.barrett ? 1 : 2, 1, (int)R.logn, a, inv ? R.psiinv : R.psi, inv ? R.psiinv_s : R.psi_s, R.lc, R.q, R.mu, R.qbit,
                   num, division, gp, gs);
}
// fused contig-forward (.) key -> contig-inverse, as launch_fused_mul in csrc/ntt_launch.cu; lazy = 0 uses ShoupPolicy both ways
int ring_fused(const EmuRing &R, bool lazy, u64 *a, unsigned group_polys, size_t group_stride, const u64 *key, const u64 *key_s,
               size_t key_half_stride, unsigned r, unsigned in_off, unsigned out0, unsigned out1, unsigned items, int nout)
{
    FusedArgs F{};
    F.A.a = a; F.A.tw = R.psi; F.A.tws = R.psi_s; F.A.lc = R.lc; F.A.num = items * r; F.A.division = r; F.A.use_tma = 1;
    F.A.group_polys = group_polys; F.A.group_stride = group_stride;
    F.twi = R.psiinv; F.twis = R.psiinv_s; F.key = key; F.key_s = key_s; F.key_item_stride = 0; F.key_half_stride = key_half_stride;
    F.r = r; F.in_off = in_off; F.out_off[0] = out0; F.out_off[1] = out1; F.items = items;
    if (lazy) return nout == 2 ? run_fused_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, 2>((int)R.logn, F)
                               : run_fused_logn<ShoupLazyPolicy, ShoupLazyInvPolicy, 1>((int)R.logn, F);
    return nout == 2 ? run_fused_logn<ShoupPolicy, ShoupPolicy, 2>((int)R.logn, F) : run_fused_logn<ShoupPolicy, ShoupPolicy, 1>((int)R.logn, F);
}
}  // namespace

#define EXPORT extern "C" __attribute__((visibility("default")))
#include "../table_kernels.cuh"

// epilogue flavour the next emu_bfv calls run: 1 = the ALL_LAZY / ALL_FAST instantiations the host picks for eligible parameter sets
static int g_epi_fast = 0;
EXPORT void emu_set_epilogue_fast(int on) { g_epi_fast = on; }

EXPORT int emu_bfv(int op, unsigned n, unsigned r, const u64 *q, const u64 *mu, const u32 *qbit, const u64 *psi, const u64 *psiinv,
                   const u64 *psi_s, const u64 *psiinv_s, const LimbConst *lc, int barrett,
                   // buffers
                   unsigned char *in, int *es, u64 *sk, u64 *pk, u64 *c, const u64 *m, u64 *out, unsigned batch, u64 nonce0,
                   // constants
                   const u64 *inv_q_last, const u64 *qi_div_t, const u64 *ptg, const u64 *ipq, const u64 *bcm, u64 t, u64 gamma, u64 mu_gamma,
                   int gamma_bits, u64 neg_inv_t, u64 neg_inv_gamma, int per_item_keys)
{
    EmuRing R{n, 0, r, q, mu, qbit, psi, psiinv, psi_s, psiinv_s, lc, barrett};
    while ((1u << R.logn) < n) R.logn++;
    LimbArrays L{q, mu, qbit, inv_q_last, ipq, ptg};
    const size_t rn = (size_t)r * n;
    SalsaKey key; for (int i = 0; i < 8; i++) key.k[i] = 0x01010101u;
    if (op == 0) {          // keygen
        const size_t stride = 9 * rn + 4 * (size_t)n; const u64 nblk = stride / 64;
        ew([&] { k_salsa20_keystream(in, nblk, (u64)batch, stride, key, nonce0); });
        if (R.barrett == 1) {
            ew3(batch, 1, [&] { k_keygen_sample(in, stride, sk, pk, es, n, r, batch, q); });
            ring_ntt(R, false, sk, batch * r, r, 0, 0);
        } else {           // context path: the ternary secret is generated inside the first strided pass
            ew3(batch, 1, [&] { k_keygen_sample(in, stride, (u64 *)nullptr, pk, es, n, r, batch, q); });
            g_gen_src = in; g_gen_stride = stride;
            ring_ntt(R, false, sk, batch * r, r, r, rn, 0);
            g_gen_src = nullptr;
            ring_ntt(R, false, sk, batch * r, r, r, rn, 1);
        }
        ew3(r, batch, [&] { k_keygen_mul(pk, sk, n, r, batch, L); });
        ring_ntt(R, true, pk, batch * r, r, r, 2 * rn);
        ew3(r, batch, [&] { k_keygen_add_negate(pk, es, n, r, batch, L); });
        ring_ntt(R, false, pk, batch * r, r, r, 2 * rn);
    } else if (op == 1) {   // encrypt
        const size_t stride = 9 * (size_t)n; const u64 nblk = stride / 64;
        ew([&] { k_salsa20_keystream(in, nblk, (u64)batch, stride, key, nonce0); });
        if (R.barrett == 1) {
            ew3(batch, 1, [&] { k_encrypt_sample(in, stride, c, es, n, r, batch, q); });
            ring_ntt(R, false, c, batch * r, r, r, 2 * rn);
        } else {           // context path: u generated inside the first strided pass
            ew3(batch, 1, [&] { k_encrypt_gauss(in, stride, es, n, batch, (size_t)n); });
            g_gen_src = in; g_gen_stride = stride;
            ring_ntt(R, false, c, batch * r, r, r, 2 * rn, 0);
            g_gen_src = nullptr;
            ring_ntt(R, false, c, batch * r, r, r, 2 * rn, 1);
        }
        ew3(r, batch, [&] { k_encrypt_mul(c, pk, per_item_keys ? 2 * rn : 0, n, r, batch, L); });
        ring_ntt(R, true, c, batch * 2 * r, r, 0, 0);
        if (r > 1) ew3(2 * ((r - 1 + kEncChunk - 1) / kEncChunk), batch, [&] { if (g_epi_fast) k_encrypt_epilogue<true, int, false>(c, 2 * rn, rn, es, m, (size_t)n, n, r, 0, r - 1, c + (size_t)(r - 1) * n, 2 * rn, rn, t, qi_div_t, L); else k_encrypt_epilogue<false, int, false>(c, 2 * rn, rn, es, m, (size_t)n, n, r, 0, r - 1, c + (size_t)(r - 1) * n, 2 * rn, rn, t, qi_div_t, L); });
        ew3(2, batch, [&] { k_encrypt_last_limb(c, es, n, r, batch, L); });
    } else if (op == 3 || op == 5) {   // encrypt through the fused kernel, key loaded (op 5: non-lazy policies); companions built here
        const size_t stride = 9 * (size_t)n; const u64 nblk = stride / 64;
        u64 *pk_s = new u64[2 * rn];
        ew([&] { k_build_companions(pk, pk_s, q, R.logn, r, 2 * r); });
        ew([&] { k_salsa20_keystream(in, nblk, (u64)batch, stride, key, nonce0); });
        ew3(batch, 1, [&] { k_encrypt_gauss(in, stride, es, n, batch, (size_t)n); });
        if (op == 5) R.barrett = -1;
        g_gen_src = in; g_gen_stride = stride;
        ring_ntt(R, false, c, batch * r, r, r, 2 * rn, 0);
        g_gen_src = nullptr;
        ring_fused(R, op == 3, c, 2 * r, 2 * rn, pk, pk_s, rn, r, 0, 0, r, batch, 2);
        ring_ntt(R, true, c, batch * 2 * r, r, 0, 0, 1);
        if (r > 1) ew3(2 * ((r - 1 + kEncChunk - 1) / kEncChunk), batch, [&] { if (g_epi_fast) k_encrypt_epilogue<true, int, false>(c, 2 * rn, rn, es, m, (size_t)n, n, r, 0, r - 1, c + (size_t)(r - 1) * n, 2 * rn, rn, t, qi_div_t, L); else k_encrypt_epilogue<false, int, false>(c, 2 * rn, rn, es, m, (size_t)n, n, r, 0, r - 1, c + (size_t)(r - 1) * n, 2 * rn, rn, t, qi_div_t, L); });
        ew3(2, batch, [&] { k_encrypt_last_limb(c, es, n, r, batch, L); });
        delete[] pk_s;
    } else if (op == 4 || op == 6) {   // decrypt through the fused kernel
        const unsigned rp = r - 1;
        const size_t item = 2 * rn, c1_off = rn;
        u64 *sk_s = new u64[rn];
        ew([&] { k_build_companions(sk, sk_s, q, R.logn, r, r); });
        DecryptConsts D{t, gamma, mu_gamma, gamma >> 1, neg_inv_t, neg_inv_gamma, gamma_bits, rp, bcm};
        if (op == 6) R.barrett = -1;
        ring_ntt(R, false, c + c1_off, batch * rp, rp, rp, item, 0);
        ring_fused(R, op == 4, c, 2 * r, item, sk, sk_s, 0, rp, r, r, r, batch, 1);
        ring_ntt(R, true, c + c1_off, batch * rp, rp, rp, item, 1);
        ew3(batch, 1, [&] { if (g_epi_fast) k_decrypt_epilogue<true>(c, item, c1_off, out, (size_t)n, n, batch, D, L); else k_decrypt_epilogue<false>(c, item, c1_off, out, (size_t)n, n, batch, D, L); });
        delete[] sk_s;
    } else {                // decrypt (r = all limbs)
        const unsigned rp = r - 1;
        const size_t item = 2 * rn, c1_off = rn;
        DecryptConsts D{t, gamma, mu_gamma, gamma >> 1, neg_inv_t, neg_inv_gamma, gamma_bits, rp, bcm};
        ring_ntt(R, false, c + c1_off, batch * rp, rp, rp, item);
        ew3(rp, batch, [&] { k_decrypt_mul(c, item, c1_off, sk, per_item_keys ? rn : 0, n, rp, batch, L); });
        ring_ntt(R, true, c + c1_off, batch * rp, rp, rp, item);
        ew3(batch, 1, [&] { if (g_epi_fast) k_decrypt_epilogue<true>(c, item, c1_off, out, (size_t)n, n, batch, D, L); else k_decrypt_epilogue<false>(c, item, c1_off, out, (size_t)n, n, batch, D, L); });
    }
    return 0;
}

// stand-alone kernels: op codes in tests/emu.py
EXPORT int emu_pointwise(int op, u64 *a, const u64 *b, u64 *c, size_t n, u64 s0, u64 s1, u64 s2, int i0, unsigned u0, unsigned u1,
                         const u64 *qv, const u64 *muv, const u32 *qbitv, const u64 *aux)
{
    LimbArrays L{qv, muv, qbitv, aux, aux, aux};
    switch (op) {
    case 0: ew([&] { k_barrett(c, a, b, n, s0, s1, i0); }); break;
    case 1: ew([&] { k_barrett_int(a, s2, n, s0, s1, i0); }); break;
    case 2: ew([&] { k_mod_t(a, s2, n, s0); }); break;
    case 3: ew([&] { k_poly_add(a, b, n, s0); }); break;
    case 4: ew([&] { k_poly_add_integer(a, s2, n, s0); }); break;
    case 5: ew([&] { k_poly_sub(a, b, n, s0); }); break;
    case 6: ew([&] { k_poly_negate(a, n, s0); }); break;
    case 7: ew([&] { k_barrett_batch(c, a, b, u0, n, u1, L); }); break;
    case 8: ew([&] { k_fast_convert_t(a, c, s0, b, u0, n); }); ew([&] { k_fast_convert_gamma(a, c, s1, b, u0, i0, s2, n); }); break;
    case 9: ew([&] { k_dec_round(a, c, s0, s1, s2, n); }); break;
    case 10: ew([&] { k_divide_and_round_q_last_inplace_loop(a, b, n, s0, s2, aux[0], s1, i0); }); break;
    case 11: ew([&] { k_poly_add_negate_xq(a, b, u0, n, L); }); break;
    default: return 1;
    }
    return 0;
}
EXPORT int emu_sampling(int op, const unsigned char *in, u64 *out, u64 *out2, size_t n, unsigned u0, const u64 *q, u64 nonce, size_t stride)
{
    SalsaKey key; for (int i = 0; i < 8; i++) key.k[i] = 0x01010101u;
    switch (op) {
    case 0: ew([&] { k_salsa20_keystream((unsigned char *)out, (u64)n, (u64)u0, stride, key, nonce); }); break;
    case 1: ew([&] { k_ternary_dist_xq(in, out, u0, n, q); }); break;
    case 2: ew([&] { k_uniform_dist_xq(in, out, u0, n, q); }); break;
    case 3: ew([&] { k_convert_ternary(in, out, n, q[0]); }); break;
    case 4: ew([&] { k_convert_range((const u64 *)in, out, n, q[0]); }); break;
    default: return 1;
    }
    return 0;
}

// limb-sharded decryption on the emulator (same split as nttb200_bfv_decrypt_partial / _finish in csrc/bfv.cu)
EXPORT int emu_bfv_sharded(int op, unsigned n, unsigned r, const u64 *q, const u64 *mu, const u32 *qbit, const u64 *psi, const u64 *psiinv,
                           const u64 *psi_s, const u64 *psiinv_s, const LimbConst *lc, u64 *c_shard, const u64 *sk_shard, u64 *part, u64 *out,
                           unsigned batch, unsigned first, unsigned count, const u64 *ptg, const u64 *ipq, const u64 *bcm, u64 t, u64 gamma,
                           u64 mu_gamma, int gamma_bits, u64 neg_inv_t, u64 neg_inv_gamma, unsigned shard_half_limbs)
{
    const unsigned rp = r - 1;
    DecryptConsts D{t, gamma, mu_gamma, gamma >> 1, neg_inv_t, neg_inv_gamma, gamma_bits, rp, bcm};
    if (op == 0) {
        const size_t toff = (size_t)first * n;
        EmuRing R{n, 0, count, q + first, mu + first, qbit + first, psi + toff, psiinv + toff, psi_s + toff, psiinv_s + toff, lc + first, 0};
        while ((1u << R.logn) < n) R.logn++;
        LimbArrays Lloc{q + first, mu + first, qbit + first, nullptr, nullptr, nullptr};
        LimbArrays Lglob{q, mu, qbit, nullptr, ipq, ptg};
        if (shard_half_limbs == 0) shard_half_limbs = count;
        const size_t item = (size_t)2 * shard_half_limbs * n, c1_off = (size_t)shard_half_limbs * n;
        ring_ntt(R, false, c_shard + c1_off, batch * count, count, count, item);
        ew3(count, batch, [&] { k_decrypt_mul(c_shard, item, c1_off, sk_shard, 0, n, count, batch, Lloc); });
        ring_ntt(R, true, c_shard + c1_off, batch * count, count, count, item);
        ew3(batch, 1, [&] { k_decrypt_partial<false>(c_shard, item, c1_off, part, n, batch, first, count, D, Lglob); });
    } else {
        ew3(batch, 1, [&] { k_decrypt_finish<false, false>(part, out, (size_t)n, n, batch, D, 1, 0); });
    }
    return 0;
}


// ---- fused-epilogue encryption / fused sharded decryption building blocks on the emulator (mirror csrc/bfv.cu: enc_sample, enc_front,
// enc_finish_last, enc_finish_limbs, dec_partial, dec_finish, dec_expand16) -----------------------------------------------------------
template <int LOGN, class EPI>
static void run_strided_inv_epi_one(const NttArgs &A, const EpiArgs &E)
{
    using SC = Sched<LOGN>;
    using P = ShoupLazyInvPolicy;
    constexpr int R = 1 << SC::K1;
    const size_t n = (size_t)1 << LOGN;
    const unsigned groups = (A.num + A.group_polys - 1) / A.group_polys;
    TensorMap ms;
    EmuTmapDesc ds{};
    ds.base = (unsigned char *)A.a; ds.rank = 4;
    ds.dims[0] = n >> SC::K1; ds.dims[1] = R; ds.dims[2] = A.group_polys; ds.dims[3] = groups;
    ds.strides[0] = 8; ds.strides[1] = (n >> SC::K1) * 8; ds.strides[2] = n * 8; ds.strides[3] = A.group_stride * 8;
    ds.box[0] = 16; ds.box[1] = R > 256 ? 256 : R; ds.box[2] = 1; ds.box[3] = 1; ds.swizzle128 = 0;
    memcpy(ms.opaque, &ds, sizeof ds);
    const unsigned tiles_s1 = (unsigned)(((n >> SC::K1) >> 4) / SC::NT);
    const int tpc_s = tiles_per_cta(tiles_s1);
    emu_dim3 gs;
    gs.x = A.num * (tiles_s1 / tpc_s);
    emu_launch(gs, R * SC::NT, (size_t)tpc_s * SC::NT * R * 128 + 1024 + 64, [&] { ntt_strided_pass<P, LOGN, true, EPI>(ms, A, E); });
}
template <class EPI>
static int run_strided_inv_epi(int logn, const NttArgs &A, const EpiArgs &E)
{
    switch (logn) {
    case 11: run_strided_inv_epi_one<11, EPI>(A, E); return 0;
    case 12: run_strided_inv_epi_one<12, EPI>(A, E); return 0;
    case 13: run_strided_inv_epi_one<13, EPI>(A, E); return 0;
    case 14: run_strided_inv_epi_one<14, EPI>(A, E); return 0;
    case 15: run_strided_inv_epi_one<15, EPI>(A, E); return 0;
    case 16: run_strided_inv_epi_one<16, EPI>(A, E); return 0;
    case 17: run_strided_inv_epi_one<17, EPI>(A, E); return 0;
    }
    return 1;
}
namespace {
EmuRing ring_window(const EmuRing &R, unsigned first, unsigned count)
{
    EmuRing W = R;
    const size_t toff = (size_t)first * R.n;
    W.r = count; W.q = R.q + first; W.mu = R.mu + first; W.qbit = R.qbit + first;
    W.psi = R.psi + toff; W.psiinv = R.psiinv + toff; W.psi_s = R.psi_s + toff; W.psiinv_s = R.psiinv_s + toff; W.lc = R.lc + first;
    return W;
}
}  // namespace

static const u64 *g_inv_q_last = nullptr, *g_qi_div_t = nullptr;     // per-limb arrays of the separate epilogue kernel (set by emu_set_enc_arrays)
EXPORT void emu_set_enc_arrays(const u64 *inv_q_last, const u64 *qi_div_t) { g_inv_q_last = inv_q_last; g_qi_div_t = qi_div_t; }
// op 0: sample (want_u = i0, want_e = i1; items, nonce0) | 1: front | 2: finish_last | 3: finish_limbs (i0: fused epilogue, i1: ALL_LAZY)
EXPORT int emu_enc_blocks(int op, unsigned n, unsigned r, const u64 *q, const u64 *mu, const u32 *qbit, const u64 *psi, const u64 *psiinv,
                          const u64 *psi_s, const u64 *psiinv_s, const LimbConst *lc, unsigned first, unsigned count, unsigned slots, unsigned items,
                          u64 *c, unsigned char *ub, signed char *es8, const u64 *pk, const u64 *pk_s, u64 *cl, size_t cl_item_stride,
                          size_t cl_half_stride, const u64 *m, const EncEpiLimb *K, u64 t, unsigned tsh, u64 nonce0, int i0, int i1)
{
    EmuRing R{n, 0, r, q, mu, qbit, psi, psiinv, psi_s, psiinv_s, lc, 0};
    while ((1u << R.logn) < n) R.logn++;
    const size_t rn = (size_t)r * n;
    SalsaKey key; for (int i = 0; i < 8; i++) key.k[i] = 0x01010101u;
    EpiArgs E{};
    E.es = es8; E.K = K; E.last = q[r - 1]; E.half_last = E.last >> 1; E.t = t; E.tfix = (t + 1) >> 1; E.tsh = tsh;
    if (op == 0) {
        ew([&] { k_encrypt_sample_fused(ub, es8, n, (u64)items, key, nonce0, i0, i1); });
    } else if (op == 1) {
        EmuRing W = ring_window(R, first, count);
        const size_t item = (size_t)2 * slots * n;
        g_gen_src = ub; g_gen_stride = n;
        ring_ntt(W, false, c, items * count, count, count, item, 0);
        g_gen_src = nullptr;
        ring_fused(W, true, c, 2 * slots, item, pk + (size_t)first * n, pk_s + (size_t)first * n, rn, count, 0, 0, slots, items, 2);
    } else if (op == 2) {
        EmuRing W = ring_window(R, r - 1, 1);
        NttArgs A{};
        A.a = cl; A.tw = W.psiinv; A.tws = W.psiinv_s; A.lc = W.lc; A.num = items * 2; A.division = 1; A.use_tma = 1;
        A.group_polys = 1; A.group_stride = cl_half_stride;
        if (cl_item_stride != 2 * cl_half_stride) return 2;
        return run_strided_inv_epi<EncLastEpi>((int)R.logn, A, E);
    } else if (op == 3) {
        EmuRing W = ring_window(R, first, count);
        NttArgs A{};
        A.a = c; A.tw = W.psiinv; A.tws = W.psiinv_s; A.lc = W.lc; A.num = items * 2 * count; A.division = count; A.use_tma = 1;
        A.group_polys = count; A.group_stride = (size_t)slots * n;
        E.cl = cl; E.cl_item_stride = cl_item_stride; E.cl_half_stride = cl_half_stride; E.m = m; E.m_stride = n; E.first_limb = first;
        if (i0) return run_strided_inv_epi<EncLimbEpi>((int)R.logn, A, E);          // A/B variant: epilogue in the kernel's store
        ring_ntt(W, true, c, items * 2 * count, count, count, (size_t)slots * n, 1);  // default: plain strided inverse pass + epilogue kernel
        LimbArrays L{q, mu, qbit, g_inv_q_last, nullptr, nullptr};
        const size_t item = (size_t)2 * slots * n, half = (size_t)slots * n;
        const unsigned rows = 2 * ((count + kEncChunk - 1) / kEncChunk);
        if (i1) ew3(rows, items, [&] { k_encrypt_epilogue<true, signed char, true>(c, item, half, es8, m, (size_t)n, n, r, first, count, cl, cl_item_stride, cl_half_stride, t, g_qi_div_t, L); });
        else ew3(rows, items, [&] { k_encrypt_epilogue<false, signed char, true>(c, item, half, es8, m, (size_t)n, n, r, first, count, cl, cl_item_stride, cl_half_stride, t, g_qi_div_t, L); });
    } else {
        return 1;
    }
    return 0;
}

// op 0: partial sums of limbs [first, first+count) through the fused kernel (loaded key) | 1: finish | 2: expand 16-bit plaintexts
EXPORT int emu_dec_blocks(int op, unsigned n, unsigned r, const u64 *q, const u64 *mu, const u32 *qbit, const u64 *psi, const u64 *psiinv,
                          const u64 *psi_s, const u64 *psiinv_s, const LimbConst *lc, unsigned first, unsigned count, unsigned slots, unsigned items,
                          u64 *c_shard, const u64 *sk, const u64 *sk_s, u64 *part, void *out, int packed, int out16, const u64 *ptg, const u64 *ipq,
                          const u64 *bcm, u64 t, u64 gamma, u64 mu_gamma, int gamma_bits, u64 neg_inv_t, u64 neg_inv_gamma)
{
    EmuRing R{n, 0, r, q, mu, qbit, psi, psiinv, psi_s, psiinv_s, lc, 0};
    while ((1u << R.logn) < n) R.logn++;
    DecryptConsts D{t, gamma, mu_gamma, gamma >> 1, neg_inv_t, neg_inv_gamma, gamma_bits, r - 1, bcm};
    LimbArrays Lglob{q, mu, qbit, nullptr, ipq, ptg};
    if (op == 0) {
        EmuRing W = ring_window(R, first, count);
        const size_t item = (size_t)2 * slots * n, c1_off = (size_t)slots * n;
        ring_ntt(W, false, c_shard + c1_off, items * count, count, count, item, 0);
        ring_fused(W, true, c_shard, 2 * slots, item, sk + (size_t)first * n, sk_s + (size_t)first * n, 0, count, slots, slots, slots, items, 1);
        ring_ntt(W, true, c_shard + c1_off, items * count, count, count, item, 1);
        if (packed) ew3(items, 1, [&] { k_decrypt_partial<true>(c_shard, item, c1_off, part, n, items, first, count, D, Lglob); });
        else ew3(items, 1, [&] { k_decrypt_partial<false>(c_shard, item, c1_off, part, n, items, first, count, D, Lglob); });
    } else if (op == 1) {
        if (packed && out16) ew3(items, 1, [&] { k_decrypt_finish<true, true>(part, out, (size_t)n, n, items, D, 1, 0); });
        else if (packed) ew3(items, 1, [&] { k_decrypt_finish<true, false>(part, out, (size_t)n, n, items, D, 1, 0); });
        else if (out16) ew3(items, 1, [&] { k_decrypt_finish<false, true>(part, out, (size_t)n, n, items, D, 1, 0); });
        else ew3(items, 1, [&] { k_decrypt_finish<false, false>(part, out, (size_t)n, n, items, D, 1, 0); });
    } else if (op == 2) {
        ew([&] { k_expand16((const unsigned short *)part, (u64 *)out, (size_t)items * n, 0); });
    } else {
        return 1;
    }
    return 0;
}
EXPORT unsigned emu_sizeof_encepilimb() { return (unsigned)sizeof(EncEpiLimb); }

// wire format + homomorphic helper kernels on the emulator (op 0: pack, 1: unpack, 2: ct add, 3: plaintext lift)
EXPORT int emu_ct_ops(int op, u64 *c, u64 *other, unsigned n, unsigned r, unsigned batch, const u64 *q, const u32 *qbit, const u32 *word_off,
                      u32 half_words, u64 t)
{
    emu_dim3 g; g.x = 2; g.y = r - 1; g.z = 2 * batch;
    switch (op) {
    case 0: emu_launch(g, 64, 0, [&] { k_ct_pack(c, other, n, r, batch, qbit, word_off, half_words); }); break;
    case 1: emu_launch(g, 64, 0, [&] { k_ct_unpack(other, c, n, r, batch, qbit, word_off, half_words); }); break;
    case 2: emu_launch(g, 64, 0, [&] { k_ct_add(c, other, n, r, batch, q); }); break;
    case 3: { emu_dim3 g2; g2.x = 2; g2.y = batch; emu_launch(g2, 64, 0, [&] { k_plain_lift(other, (size_t)n, c, n, r - 1, batch, t, q); }); break; }
    default: return 1;
    }
    return 0;
}

// both ternary formulas for one byte (exhaustive equivalence test)
EXPORT int emu_ternary(int which, unsigned byte)
{
    if (which == 0) return ternary_float_formula((unsigned char)byte);
    const u64 v = ternary_value((unsigned char)byte, 1000);
    return v == 999 ? -1 : (int)v;
}

// device-side table generation on the emulator
EXPORT int emu_build_tables(u64 *psi, u64 *psi_s, u64 *psiinv, u64 *psiinv_s, const u64 *q, const u64 *roots, const u64 *roots_inv,
                            unsigned logn, unsigned limbs)
{
    emu_dim3 g; g.x = limbs; g.y = 2;
    emu_launch(g, 128, 0, [&] { k_build_tables(psi, psi_s, psiinv, psiinv_s, q, roots, roots_inv, logn); });
    return 0;
}

// ---- BFV multiplication kernels (csrc/mul_kernels.cuh) on the emulator: plain constants in, the device structures are built here -----
#include "../mul_kernels.cuh"
namespace {
ModC emu_modc(u64 q)
{
    ModC m; m.q = q; m.ratio = ~0ull / q; m.qbit = 0; while ((q >> m.qbit) != 0) m.qbit++;
    m.mu = (u64)(((unsigned __int128)1 << (2 * m.qbit)) / q); m.pad = 0;
    return m;
}
ShoupC emu_shoupc(u64 c, u64 q) { ShoupC s; s.c = c; s.cs = (u64)(((unsigned __int128)c << 64) / q); return s; }
SplitC emu_splitc(u64 c, unsigned h) { SplitC r; r.c0 = (u32)(c & ((1ull << h) - 1)); r.c1 = (u32)(c >> h); return r; }
}  // namespace

// x[items][lin][n] -> out[items][lout][n]; pre[lin] = (B/b_i)^-1 mod b_i, M[lout][lin] = B/b_i mod o_j, corr[lout] = B mod o_j
EXPORT int emu_mul_bconv(const u64 *x, u64 *out, unsigned items, unsigned lin, unsigned lout, unsigned n, unsigned h, const u64 *bin,
                         const u64 *bout, const u64 *pre, const double *binv, const u64 *M, const u64 *corr)
{
    std::vector<ShoupC> preS(lin), r64(lout);
    std::vector<ModC> min_(lin), mout(lout);
    std::vector<SplitC> Ms((size_t)lin * lout);
    for (unsigned i = 0; i < lin; i++) { min_[i] = emu_modc(bin[i]); preS[i] = emu_shoupc(pre[i], bin[i]); }
    for (unsigned j = 0; j < lout; j++) {
        mout[j] = emu_modc(bout[j]);
        r64[j] = emu_shoupc((u64)(((unsigned __int128)1 << 64) % bout[j]), bout[j]);
        for (unsigned i = 0; i < lin; i++) Ms[(size_t)j * lin + i] = emu_splitc(M[(size_t)j * lin + i], h);
    }
    BconvArgs A{x, (size_t)lin * n, out, (size_t)lout * n, preS.data(), min_.data(), binv, Ms.data(), corr, r64.data(), mout.data(), lin, lout, n, h};
    emu_dim3 g; g.x = (n + 127) / 128; g.y = items;
    emu_launch(g, 128, ((size_t)lin * lout + 5 * lout) * 8, [&] { if (lin <= 16) k_bconv<16>(A); else k_bconv<kBaseMax>(A); });
    return 0;
}
// d[kc][rp + k][n] -> y[kc][k][n]
EXPORT int emu_mul_scale(const u64 *d, u64 *y, unsigned kc, unsigned rp, unsigned k, unsigned n, unsigned h, const u64 *qQ, const u64 *qP,
                         const u64 *preQ, const double *theta, const u64 *W, const u64 *lam)
{
    std::vector<ShoupC> preS(rp), r64(k);
    std::vector<ModC> mQ(rp), mP(k);
    std::vector<SplitC> Ws((size_t)rp * k), lams(k);
    for (unsigned i = 0; i < rp; i++) { mQ[i] = emu_modc(qQ[i]); preS[i] = emu_shoupc(preQ[i], qQ[i]); }
    for (unsigned j = 0; j < k; j++) {
        mP[j] = emu_modc(qP[j]);
        r64[j] = emu_shoupc((u64)(((unsigned __int128)1 << 64) % qP[j]), qP[j]);
        lams[j] = emu_splitc(lam[j], h);
        for (unsigned i = 0; i < rp; i++) Ws[(size_t)j * rp + i] = emu_splitc(W[(size_t)j * rp + i], h);
    }
    ScaleArgs S{d, y, preS.data(), mQ.data(), mP.data(), theta, Ws.data(), lams.data(), r64.data(), rp, k, n, h};
    emu_dim3 g; g.x = (n + 127) / 128; g.y = kc;
    emu_launch(g, 128, ((size_t)rp * k + 5 * k) * 8, [&] { if (rp <= 16) k_scale<16>(S); else k_scale<kBaseMax>(S); });
    return 0;
}
// D[items][rp][rp][n], evk[rp][2][rp][n] -> acc[items][2][rp][n]
EXPORT int emu_mul_relin_accum(const u64 *D, const u64 *evk, u64 *acc, unsigned n, unsigned rp, unsigned items, unsigned h, const u64 *qQ)
{
    std::vector<ModC> mQ(rp);
    std::vector<ShoupC> r64(rp);
    for (unsigned i = 0; i < rp; i++) { mQ[i] = emu_modc(qQ[i]); r64[i] = emu_shoupc((u64)(((unsigned __int128)1 << 64) % qQ[i]), qQ[i]); }
    emu_dim3 g; g.x = (items + kAccumItems - 1) / kAccumItems; g.y = (n + 255) / 256; g.z = rp;
    emu_launch(g, 128, 0, [&] { k_relin_accum(D, evk, acc, n, rp, items, h, mQ.data(), r64.data()); });
    return 0;
}
