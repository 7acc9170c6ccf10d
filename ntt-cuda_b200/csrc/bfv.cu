// bfv.cu -- BFV keygen / encrypt / decrypt orchestration (replaces bfv_keygen.cuh:95-151 keygen_rns,
// bfv_encryption.cuh:223-290 encryption_rns, bfv_decryption.cuh:76-138 decryption_rns and the constant derivation
// every reference driver repeats, demo.cu:62-264).
//
// One implementation serves two front ends:
//   * nttb200_bfv_*      : context-based and BATCHED (item k = Salsa20 nonce nonce0 + k), Shoup NTT on context tables;
//   * nttb200_ref_*_rns  : the reference's single-item calls with the reference's own tables / constant arrays,
//                          stateless Barrett NTT (what include/dropin/bfv_*.cuh forwards to).
// Launch counts (any batch size): keygen 10, encrypt 8, decrypt 6 kernels on ONE stream -- the reference issues
// 13 / 16 / 18 per item, creates r streams and mallocs inside decryption_rns, and races dec_round against mod_t.
#include "internal.h"
#include "bfv_kernels.cuh"
#include "epi.cuh"
#include "table_kernels.cuh"
#include "bfv_internal.h"
#include "launch_util.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

using namespace nttb200;
typedef unsigned __int128 u128;

namespace nttb200 {
dim3 grid_for(size_t total, int threads);
SalsaKey default_key();
}
// grid of the fused kernels: x strides over coefficient pairs (256 threads x 2 coefficients per CTA step), y / z carry
// limb or half / batch item; x is trimmed so the whole grid stays around 8 CTAs per SM
// threads per CTA of the fused kernels: 256, or 64 when the whole problem would not even give two 256-thread CTAs per SM
// (single-item calls are latency-bound: many small CTAs spread over all SMs beat few large ones)
static unsigned pair_block(unsigned n, unsigned y, unsigned z)
{
    const unsigned cap = grid_for((size_t)1 << 40, 256).x;          // = 8 * SM count
    const unsigned long long ctas256 = (unsigned long long)((n + 511) / 512) * y * z;
    return ctas256 * 4 < cap ? 64u : 256u;
}
static dim3 pair_grid(unsigned n, unsigned y, unsigned z)
{
    const unsigned cap = grid_for((size_t)1 << 40, 256).x;
    const unsigned per_cta = 2 * pair_block(n, y, z);
    unsigned x = (n + per_cta - 1) / per_cta;
    const unsigned long long yz = (unsigned long long)y * z;
    unsigned lim = (unsigned)(yz >= cap ? 1 : cap / yz);
    if (x > lim) x = lim;
    return dim3(x ? x : 1, y, z);
}

static u64 h_modpow(u64 a, u64 e, u64 m)
{
    u64 r = 1 % m; a %= m;
    while (e) { if (e & 1) r = (u64)((u128)r * a % m); a = (u64)((u128)a * a % m); e >>= 1; }
    return r;
}
// deterministic Miller-Rabin for 64-bit integers
static bool h_is_prime(u64 m)
{
    if (m < 2) return false;
    for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (m % p == 0) return m == p;
    }
    u64 d = m - 1; int sft = 0;
    while (!(d & 1)) { d >>= 1; sft++; }
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        u64 x = h_modpow(a, d, m);
        if (x == 1 || x == m - 1) continue;
        bool comp = true;
        for (int i = 1; i < sft && comp; i++) { x = (u64)((u128)x * x % m); if (x == m - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}
// the reference's modinv128(a, m) = a^(m-2) mod m (helper.h:52-56), also applied to the non-prime t (demo.cu:109)
static u64 h_modinv_fermat(u64 a, u64 m) { return h_modpow(a, m - 2, m); }

static NttArgsHost pipe_args(const Pipe &P, bool inverse, u64 *a, unsigned num, unsigned division, unsigned group_polys, size_t group_stride)
{
    NttArgsHost h{a, inverse ? P.psiinv : P.psi, inverse ? P.psiinv_s : P.psi_s, P.lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division, P.use_tma,
                  group_polys, group_stride};
    return h;
}
// only the first / second kernel of a transform (execution order), context path
static int pipe_ntt_pass(const Pipe &P, bool inverse, int which, u64 *a, unsigned num, unsigned division, unsigned group_polys, size_t group_stride)
{
    return launch_ntt_pass(inverse, inverse ? P.policy_inv : P.policy_fwd, P.logn, pipe_args(P, inverse, a, num, division, group_polys, group_stride),
                           which, P.st);
}
// forward strided pass whose input u is generated from the keystream inside the kernel (context path)
static int pipe_ntt_gen_pass(const Pipe &P, u64 *a, unsigned num, unsigned division, unsigned group_polys, size_t group_stride,
                             const unsigned char *in, size_t in_stride)
{
    NttArgsHost h = pipe_args(P, false, a, num, division, group_polys, group_stride);
    h.gen_src = in; h.gen_stride = in_stride;
    return launch_ntt_pass(false, P.policy_fwd, P.logn, h, 0, P.st);
}
static int pipe_ntt(const Pipe &P, bool inverse, u64 *a, unsigned num, unsigned division, unsigned group_polys, size_t group_stride)
{
    NttArgsHost h{a, inverse ? P.psiinv : P.psi, inverse ? P.psiinv_s : P.psi_s, P.lc, P.L.q, P.L.mu, P.L.qbit, 0, 0, 0, num, division, P.use_tma,
                  group_polys, group_stride};
    const int pol = inverse ? P.policy_inv : P.policy_fwd;
    if (pol != kPolicyBarrett) { h.qv = nullptr; h.muv = nullptr; h.qbitv = nullptr; }
    return launch_ntt(inverse, pol, P.logn, h, P.st);
}


// keygen: in = keystream scratch ([batch] streams of in_stride bytes), es = n ints per item
static int run_keygen(const Pipe &P, unsigned char *in, size_t in_stride, int *es, u64 *sk, u64 *pk, unsigned batch, u64 nonce0)
{
    const unsigned n = P.n, r = P.r;
    const size_t rn = (size_t)r * n;
    const u64 nblk = (9 * rn + 4 * (size_t)n) / 64;                                   // bfv_keygen.cuh:99
    NTT_LAUNCH_PDL(k_salsa20_keystream, dim3(grid_for(nblk * batch, 256)), dim3(256), 0, P.st, in, nblk, (u64)batch, in_stride, P.key, nonce0, (u64)0);
    if (P.policy_fwd != kPolicyBarrett) {      // context path: the ternary secret is generated inside the first strided pass
        NTT_LAUNCH_PDL(k_keygen_sample, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, in, in_stride, nullptr, pk, es, n, r, batch, P.L.q);   // :121-122
        KCHECK();
        NTTB200_TRY(pipe_ntt_gen_pass(P, sk, batch * r, r, r, rn, in, in_stride));         // :120 + :129, first kernel
        NTTB200_TRY(pipe_ntt_pass(P, false, 1, sk, batch * r, r, r, rn));
    } else {
        NTT_LAUNCH_PDL(k_keygen_sample, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, in, in_stride, sk, pk, es, n, r, batch, P.L.q);   // :120-122
        KCHECK();
        NTTB200_TRY(pipe_ntt(P, false, sk, batch * r, r, 0, 0));                           // :129
    }
    if (P.all_exact && P.policy_inv != kPolicyBarrett) {
        // pk0 = INTT(a (.) NTT(s)): the product rides in the first inverse kernel (both operands are canonical NTT-domain values)
        NTTB200_TRY(launch_polymul(P.policy_inv == kPolicyShoupLazy, P.logn, pipe_args(P, false, pk + rn, batch * r, r, r, 2 * rn), P.psiinv,
                                   P.psiinv_s, sk, 0, 0, false, pk, P.st));                // :132
        NTTB200_TRY(pipe_ntt_pass(P, true, 1, pk, batch * r, r, r, 2 * rn));               // :133 (second kernel)
    } else {
        NTT_LAUNCH_PDL(k_keygen_mul, dim3(pair_grid(n, r, batch)), dim3(pair_block(n, r, batch)), 0, P.st, pk, sk, n, r, batch, P.L);                  // :132
        KCHECK();
        NTTB200_TRY(pipe_ntt(P, true, pk, batch * r, r, r, 2 * rn));                       // :133
    }
    NTT_LAUNCH_PDL(k_keygen_add_negate, dim3(pair_grid(n, r, batch)), dim3(pair_block(n, r, batch)), 0, P.st, pk, es, n, r, batch, P.L);               // :144
    KCHECK();
    NTTB200_TRY(pipe_ntt(P, false, pk, batch * r, r, r, 2 * rn));                          // :145
    return 0;
}

// mod-switch + Delta*m on the limbs below the dropped one (limb-chunked rows), then the dropped limb's padding value
static int run_encrypt_epilogue(const Pipe &P, u64 *c, const int *es, const u64 *m, size_t m_stride, u64 t, unsigned batch)
{
    const unsigned n = P.n, r = P.r;
    if (r > 1) {
        const unsigned rows = 2 * ((r - 1 + kEncChunk - 1) / kEncChunk);
        const size_t rn = (size_t)r * n;
        const u64 *cl = c + (size_t)(r - 1) * n;      // the padding slot still holds the RAW inverse-transform output here
        if (P.enc_lazy) NTT_LAUNCH_PDL(k_encrypt_epilogue<true, int, false>, dim3(pair_grid(n, rows, batch)), dim3(pair_block(n, rows, batch)), 0, P.st, c, 2 * rn, rn, es, m, m_stride, n, r, 0, r - 1, cl, 2 * rn, rn, t, P.qi_div_t, P.L);
        else NTT_LAUNCH_PDL(k_encrypt_epilogue<false, int, false>, dim3(pair_grid(n, rows, batch)), dim3(pair_block(n, rows, batch)), 0, P.st, c, 2 * rn, rn, es, m, m_stride, n, r, 0, r - 1, cl, 2 * rn, rn, t, P.qi_div_t, P.L);
    }
    NTT_LAUNCH_PDL(k_encrypt_last_limb, dim3(pair_grid(n, 2, batch)), dim3(pair_block(n, 2, batch)), 0, P.st, c, es, n, r, batch, P.L);
    KCHECK();
    return 0;
}

static int run_encrypt(const Pipe &P, unsigned char *in, size_t in_stride, int *es, u64 *c, const u64 *pk, size_t pk_stride, const u64 *m,
                       size_t m_stride, u64 t, unsigned batch, u64 nonce0)
{
    const unsigned n = P.n, r = P.r;
    const size_t rn = (size_t)r * n;
    const u64 nblk = (9 * (size_t)n) / 64;                                                // bfv_encryption.cuh:228
    NTT_LAUNCH_PDL(k_salsa20_keystream, dim3(grid_for(nblk * batch, 256)), dim3(256), 0, P.st, in, nblk, (u64)batch, in_stride, P.key, nonce0, (u64)0);
    if (P.policy_fwd != kPolicyBarrett) {
        NTT_LAUNCH_PDL(k_encrypt_gauss, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, in, in_stride, es, n, batch, (size_t)n);               // :247 (e0, e1)
        KCHECK();
        NTTB200_TRY(pipe_ntt_gen_pass(P, c, batch * r, r, r, 2 * rn, in, in_stride));      // :247 (u) + :268 (once, not twice), first kernel
        NTTB200_TRY(pipe_ntt_pass(P, false, 1, c, batch * r, r, r, 2 * rn));
    } else {
        NTT_LAUNCH_PDL(k_encrypt_sample, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, in, in_stride, c, es, n, r, batch, P.L.q);   // :247
        KCHECK();
        NTTB200_TRY(pipe_ntt(P, false, c, batch * r, r, r, 2 * rn));                       // :268 (once, not twice)
    }
    NTT_LAUNCH_PDL(k_encrypt_mul, dim3(pair_grid(n, r, batch)), dim3(pair_block(n, r, batch)), 0, P.st, c, pk, pk_stride, n, r, batch, P.L);              // :270
    KCHECK();
    NTTB200_TRY(pipe_ntt(P, true, c, batch * 2 * r, r, 0, 0));                             // :271
    NTTB200_TRY(run_encrypt_epilogue(P, c, es, m, m_stride, t, batch));                    // :280-289
    KCHECK();
    return 0;
}

// encryption with a key loaded into the context: strided forward pass, ONE fused kernel (contig forward pass, (.) pk0 / pk1 by
// Shoup companions, contig inverse pass for both halves), strided inverse pass, epilogue.  7 launches.
static int run_encrypt_fused(const Pipe &P, bool lazy, unsigned char *in, size_t in_stride, int *es, u64 *c, const u64 *pk, const u64 *pk_s,
                             const u64 *m, size_t m_stride, u64 t, unsigned batch, u64 nonce0)
{
    const unsigned n = P.n, r = P.r;
    const size_t rn = (size_t)r * n;
    const u64 nblk = (9 * (size_t)n) / 64;
    NTT_LAUNCH_PDL(k_salsa20_keystream, dim3(grid_for(nblk * batch, 256)), dim3(256), 0, P.st, in, nblk, (u64)batch, in_stride, P.key, nonce0, (u64)0);
    NTT_LAUNCH_PDL(k_encrypt_gauss, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, in, in_stride, es, n, batch, (size_t)n);
    KCHECK();
    NTTB200_TRY(pipe_ntt_gen_pass(P, c, batch * r, r, r, 2 * rn, in, in_stride));          // strided forward pass, u generated in the kernel
    NTTB200_TRY(launch_fused_mul(lazy, P.logn, pipe_args(P, false, c, batch * 2 * r, r, 2 * r, 2 * rn), P.psiinv, P.psiinv_s, pk, pk_s, 0, rn, r,
                                 0, 0, r, batch, 2, P.st));
    NTTB200_TRY(pipe_ntt_pass(P, true, 1, c, batch * 2 * r, r, 0, 0));                     // strided inverse pass on both halves
    NTTB200_TRY(run_encrypt_epilogue(P, c, es, m, m_stride, t, batch));
    KCHECK();
    return 0;
}
// encryption with a loaded key on a lazy-policy ring, 5 launches: sampling (u bytes + signed-byte gaussian draws), strided forward
// pass generating u, fused contig forward (.) pk0 | pk1 -> contig inverse, strided inverse pass of the dropped limb (+ e, rounding
// offset in its store), strided inverse pass of the other limbs with mod-switch + Delta*m in its store.  c is written once.
static int run_encrypt_v2(nttb200_bfv *b, const Pipe &P, u64 *c, const u64 *m, unsigned batch, u64 nonce0)
{
    const unsigned n = P.n, r = P.r;
    const size_t rn = (size_t)r * n;
    NTTB200_TRY(ensure_enc_scratch(b, (size_t)batch * n, (size_t)batch * 2 * n));
    const u64 per = 9 * (u64)n / 64;
    NTT_LAUNCH_PDL(k_encrypt_sample_fused, dim3(grid_for(per * batch, 128)), dim3(128), 0, P.st, b->ub, b->es8, n, (u64)batch, P.key, nonce0, 1, 1);
    KCHECK();
    NTTB200_TRY(enc_front(b, P, c, r, 0, r, batch, b->ub));
    u64 *cl = c + (size_t)(r - 1) * n;
    NTTB200_TRY(enc_finish_last(b, P, cl, 2 * rn, rn, b->es8, batch));
    NTTB200_TRY(enc_finish_limbs(b, P, c, r, 0, r - 1, batch, cl, 2 * rn, rn, b->es8, m, (size_t)n));
    return 0;
}
// decryption with a loaded secret key: 4 launches
static int run_decrypt_fused(const Pipe &P, bool lazy, u64 *c, const u64 *sk, const u64 *sk_s, u64 *out, size_t out_stride, const DecryptConsts &D,
                             unsigned batch)
{
    const unsigned n = P.n, rp = D.rp;
    const size_t item = (size_t)2 * (rp + 1) * n, c1_off = (size_t)(rp + 1) * n;
    NTTB200_TRY(pipe_ntt_pass(P, false, 0, c + c1_off, batch * rp, rp, rp, item));
    NTTB200_TRY(launch_fused_mul(lazy, P.logn, pipe_args(P, false, c, batch * 2 * (rp + 1), rp, 2 * (rp + 1), item), P.psiinv, P.psiinv_s, sk, sk_s,
                                 0, 0, rp, rp + 1, rp + 1, rp + 1, batch, 1, P.st));
    NTTB200_TRY(pipe_ntt_pass(P, true, 1, c + c1_off, batch * rp, rp, rp, item));
    if (P.dec_fast) NTT_LAUNCH_PDL(k_decrypt_epilogue<true>, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, c, item, c1_off, out, out_stride, n, batch, D, P.L);
    else NTT_LAUNCH_PDL(k_decrypt_epilogue<false>, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, c, item, c1_off, out, out_stride, n, batch, D, P.L);   // :103-137
    KCHECK();
    return 0;
}

// c: items of [2][rp+1][n]; the plaintext of item k goes to out + k*out_stride
static int run_decrypt(const Pipe &P, u64 *c, const u64 *sk, size_t sk_stride, u64 *out, size_t out_stride, const DecryptConsts &D, unsigned batch)
{
    const unsigned n = P.n, rp = D.rp;
    const size_t item = (size_t)2 * (rp + 1) * n, c1_off = (size_t)(rp + 1) * n;
    NTTB200_TRY(pipe_ntt(P, false, c + c1_off, batch * rp, rp, rp, item));                 // bfv_decryption.cuh:98
    NTT_LAUNCH_PDL(k_decrypt_mul, dim3(pair_grid(n, rp, batch)), dim3(pair_block(n, rp, batch)), 0, P.st, c, item, c1_off, sk, sk_stride, n, rp, batch, P.L);   // :100
    KCHECK();
    NTTB200_TRY(pipe_ntt(P, true, c + c1_off, batch * rp, rp, rp, item));                  // :101
    if (P.dec_fast) NTT_LAUNCH_PDL(k_decrypt_epilogue<true>, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, c, item, c1_off, out, out_stride, n, batch, D, P.L);
    else NTT_LAUNCH_PDL(k_decrypt_epilogue<false>, dim3(pair_grid(n, batch, 1)), dim3(pair_block(n, batch, 1)), 0, P.st, c, item, c1_off, out, out_stride, n, batch, D, P.L);   // :103-137
    KCHECK();
    return 0;
}

static unsigned ilog2u(unsigned n) { unsigned l = 0; while ((1u << l) < n) l++; return l; }

namespace nttb200 {
Pipe pipe_from_bfv(const nttb200_bfv *b, cudaStream_t st)
{
    const nttb200_ctx *c = b->ctx;
    Pipe P;
    P.n = b->n; P.logn = c->logn; P.r = b->r;
    P.L = LimbArrays{c->q_dev, c->mu_dev, c->qbit_dev, b->inv_q_last_mod_q, b->inv_punctured_q, b->prod_t_gamma_mod_q};
    P.qi_div_t = b->qi_div_t;
    P.enc_lazy = b->enc_lazy; P.dec_fast = b->dec_fast; P.all_exact = b->all_exact;
    P.policy_fwd = P.policy_inv = c->lazy_ok ? kPolicyShoupLazy : kPolicyShoup;
    P.psi = c->psi; P.psiinv = c->psiinv; P.psi_s = c->psi_s; P.psiinv_s = c->psiinv_s; P.lc = c->lc;
    P.use_tma = c->use_tma; P.st = st;
    P.key = bfv_salsa_key(b);
    return P;
}
Pipe pipe_limb_window(const Pipe &P0, const nttb200_ctx *c, unsigned first)
{
    Pipe P = P0;
    const size_t toff = (size_t)first * P.n;
    P.psi = c->psi + toff; P.psiinv = c->psiinv + toff; P.psi_s = c->psi_s + toff; P.psiinv_s = c->psiinv_s + toff; P.lc = c->lc + first;
    return P;
}
SalsaKey key_from_bytes(const unsigned char *b);
SalsaKey bfv_salsa_key(const nttb200_bfv *b) { return key_from_bytes(b->salsa_key); }

int ensure_enc_scratch(nttb200_bfv *b, size_t ub_bytes, size_t es8_bytes)
{
    if (b->ub_bytes < ub_bytes) {
        if (b->ub) cudaFree(b->ub);
        b->ub = nullptr; b->ub_bytes = 0;
        NTTB200_CHECK(cudaMalloc(&b->ub, ub_bytes));
        b->ub_bytes = ub_bytes;
    }
    if (b->es8_bytes < es8_bytes) {
        if (b->es8) cudaFree(b->es8);
        b->es8 = nullptr; b->es8_bytes = 0;
        NTTB200_CHECK(cudaMalloc(&b->es8, es8_bytes));
        b->es8_bytes = es8_bytes;
    }
    return 0;
}

// ---- fused-epilogue encryption building blocks (see bfv_internal.h) ---------------------------------------------------------------
int enc_front(const nttb200_bfv *b, const Pipe &P0, u64 *c, unsigned slots, unsigned first, unsigned count, unsigned items, const unsigned char *ub)
{
    const unsigned n = P0.n;
    const size_t rn = (size_t)b->r * n, item = (size_t)2 * slots * n;
    Pipe P = pipe_limb_window(P0, b->ctx, first);
    NTTB200_TRY(pipe_ntt_gen_pass(P, c, items * count, count, count, item, ub, (size_t)n));      // half 0, slots [0, count)
    NTTB200_TRY(launch_fused_mul(true, P.logn, pipe_args(P, false, c, items * 2 * slots, count, 2 * slots, item), P.psiinv, P.psiinv_s,
                                 b->pk_l + (size_t)first * n, b->pk_ls + (size_t)first * n, 0, rn, count, 0, 0, slots, items, 2, P.st));
    return 0;
}
static EpiArgs epi_args(const nttb200_bfv *b, const signed char *es8)
{
    EpiArgs E{};
    E.es = es8; E.K = b->enc_epi;
    E.last = b->ctx->q[b->r - 1]; E.half_last = E.last >> 1;
    E.t = b->t; E.tfix = (b->t + 1) >> 1; E.tsh = b->tsh;
    return E;
}
int enc_finish_last(const nttb200_bfv *b, const Pipe &P0, u64 *cl, size_t cl_item_stride, size_t cl_half_stride, const signed char *es8, unsigned items)
{
    if (cl_item_stride != 2 * cl_half_stride) return NTTB200_EINVAL;
    Pipe P = pipe_limb_window(P0, b->ctx, b->r - 1);
    EpiArgs E = epi_args(b, es8);
    return launch_strided_inv_epi(P.logn, pipe_args(P, true, cl, items * 2, 1, 1, cl_half_stride), kEpiEncLast, E, P.st);
}
int enc_finish_limbs(const nttb200_bfv *b, const Pipe &P0, u64 *c, unsigned slots, unsigned first, unsigned count, unsigned items, const u64 *cl,
                     size_t cl_item_stride, size_t cl_half_stride, const signed char *es8, const u64 *m, size_t m_stride)
{
    Pipe P = pipe_limb_window(P0, b->ctx, first);
    const unsigned n = P.n;
    if (!b->no_fused_epilogue) {          // A/B variant: mod-switch + Delta*m in the store of the last inverse kernel
        EpiArgs E = epi_args(b, es8);
        E.cl = cl; E.cl_item_stride = cl_item_stride; E.cl_half_stride = cl_half_stride;
        E.m = m; E.m_stride = m_stride; E.first_limb = first;
        return launch_strided_inv_epi(P.logn, pipe_args(P, true, c, items * 2 * count, count, count, (size_t)slots * n), kEpiEncLimb, E, P.st);
    }
    // default: plain strided inverse pass, then the HBM-bound epilogue pass (its arithmetic hides under its memory time, whereas inside
    // the issue-bound NTT kernel the same arithmetic costs more than the pass it saves: profiles/r02_experiments.md)
    NTTB200_TRY(launch_ntt_pass(true, P.policy_inv, P.logn, pipe_args(P, true, c, items * 2 * count, count, count, (size_t)slots * n), 1, P.st));
    const unsigned rows = 2 * ((count + kEncChunk - 1) / kEncChunk);
    const dim3 g = pair_grid(n, rows, items);
    const unsigned tb = pair_block(n, rows, items);
    const size_t item = (size_t)2 * slots * n, half = (size_t)slots * n;
    if (b->enc_lazy) k_encrypt_epilogue<true, signed char, true><<<g, tb, 0, P.st>>>(c, item, half, es8, m, m_stride, n, b->r, first, count, cl, cl_item_stride, cl_half_stride, b->t, b->qi_div_t, P0.L);
    else k_encrypt_epilogue<false, signed char, true><<<g, tb, 0, P.st>>>(c, item, half, es8, m, m_stride, n, b->r, first, count, cl, cl_item_stride, cl_half_stride, b->t, b->qi_div_t, P0.L);
    KCHECK();
    return 0;
}
// decryption of a limb window, step 1: NTT(c1) (.) sk -> INTT on `items` ciphertext tiles c_shard[item][2][slots][n] (fused key kernel)
int dec_transforms(const nttb200_bfv *b, const Pipe &P0, u64 *c_shard, unsigned slots, unsigned first, unsigned count, unsigned items)
{
    const unsigned n = P0.n;
    const size_t item = (size_t)2 * slots * n, c1_off = (size_t)slots * n;
    Pipe P = pipe_limb_window(P0, b->ctx, first);
    NTTB200_TRY(pipe_ntt_pass(P, false, 0, c_shard + c1_off, items * count, count, count, item));
    NTTB200_TRY(launch_fused_mul(b->ctx->lazy_ok != 0, P.logn, pipe_args(P, false, c_shard, items * 2 * slots, count, 2 * slots, item), P.psiinv,
                                 P.psiinv_s, b->sk_l + (size_t)first * n, b->sk_ls + (size_t)first * n, 0, 0, count, slots, slots, slots, items, 1, P.st));
    NTTB200_TRY(pipe_ntt_pass(P, true, 1, c_shard + c1_off, items * count, count, count, item));
    return 0;
}
// step 2: the window's partial base-conversion sums of `items` tiles (packed: 1.25 n words per item, else 2 n)
int dec_partial_sums(const nttb200_bfv *b, const Pipe &P0, u64 *partial, int packed, const u64 *c_shard, unsigned slots, unsigned first, unsigned count,
                     unsigned items)
{
    const unsigned n = P0.n;
    const size_t item = (size_t)2 * slots * n, c1_off = (size_t)slots * n;
    DecryptConsts D{b->t, b->gamma, b->mu_gamma, b->gamma_div_2, b->neg_inv_t, b->neg_inv_gamma, b->gamma_bits, b->r - 1, b->bcm};
    if (packed) k_decrypt_partial<true><<<pair_grid(n, items, 1), pair_block(n, items, 1), 0, P0.st>>>(c_shard, item, c1_off, partial, n, items, first, count, D, P0.L);
    else k_decrypt_partial<false><<<pair_grid(n, items, 1), pair_block(n, items, 1), 0, P0.st>>>(c_shard, item, c1_off, partial, n, items, first, count, D, P0.L);
    KCHECK();
    return 0;
}
int dec_partial(const nttb200_bfv *b, const Pipe &P0, u64 *partial, int packed, u64 *c_shard, unsigned slots, unsigned first, unsigned count,
                unsigned items)
{
    NTTB200_TRY(dec_transforms(b, P0, c_shard, slots, first, count, items));
    return dec_partial_sums(b, P0, partial, packed, c_shard, slots, first, count, items);
}
// partial sums -> plaintext (16-bit words or u64 coefficients); expansion of gathered 16-bit plaintexts
int dec_finish(const nttb200_bfv *b, void *out, int out16, const u64 *partial_sum, int packed, unsigned items, cudaStream_t st, unsigned slots,
               size_t slot_stride)
{
    DecryptConsts D{b->t, b->gamma, b->mu_gamma, b->gamma_div_2, b->neg_inv_t, b->neg_inv_gamma, b->gamma_bits, b->r - 1, b->bcm};
    const unsigned n = b->n;
    const dim3 g = pair_grid(n, items, 1);
    const unsigned tb = pair_block(n, items, 1);
    if (packed && out16) k_decrypt_finish<true, true><<<g, tb, 0, st>>>(partial_sum, out, (size_t)n, n, items, D, slots, slot_stride);
    else if (packed) k_decrypt_finish<true, false><<<g, tb, 0, st>>>(partial_sum, out, (size_t)n, n, items, D, slots, slot_stride);
    else if (out16) k_decrypt_finish<false, true><<<g, tb, 0, st>>>(partial_sum, out, (size_t)n, n, items, D, slots, slot_stride);
    else k_decrypt_finish<false, false><<<g, tb, 0, st>>>(partial_sum, out, (size_t)n, n, items, D, slots, slot_stride);
    KCHECK();
    return 0;
}
int dec_expand16(const unsigned short *in, u64 *out, size_t total, cudaStream_t st, unsigned blocks, size_t out_block_stride)
{
    dim3 g = grid_for(total / 2, 256);
    if (blocks > 1) { g.x = std::max(1u, g.x / blocks); g.y = blocks; }
    k_expand16<<<g, 256, 0, st>>>(in, out, total, out_block_stride);
    KCHECK();
    return 0;
}
// u bytes for `items` items (want_u) and / or signed-byte gaussian draws (want_e), nonce = nonce0 + item
int enc_sample(const nttb200_bfv *b, unsigned char *ub, signed char *es8, unsigned items, u64 nonce0, int want_u, int want_e, cudaStream_t st)
{
    const u64 per = (want_u ? (u64)b->n / 64 : 0) + (want_e ? 8 * (u64)b->n / 64 : 0);
    if (!per || !items) return 0;
    k_encrypt_sample_fused<<<grid_for(per * items, 128), 128, 0, st>>>(ub, es8, b->n, (u64)items, bfv_salsa_key(b), nonce0, want_u, want_e);
    KCHECK();
    return 0;
}
}  // namespace nttb200

static int ensure_scratch(nttb200_bfv *b, size_t ks_bytes, size_t es_count)
{
    if (b->ks_bytes < ks_bytes) {
        if (b->ks) cudaFree(b->ks);
        b->ks = nullptr; b->ks_bytes = 0;
        NTTB200_CHECK(cudaMalloc(&b->ks, ks_bytes));
        b->ks_bytes = ks_bytes;
    }
    if (b->es_count < es_count) {
        if (b->es) cudaFree(b->es);
        b->es = nullptr; b->es_count = 0;
        NTTB200_CHECK(cudaMalloc(&b->es, es_count * sizeof(int)));
        b->es_count = es_count;
    }
    return 0;
}

// Runs part(stream, first item, count) for the two halves of a batch on two streams when the batch is large enough to fill the GPU
// twice: every pipeline alternates issue-bound transforms with HBM-bound element-wise kernels, and with two independent halves in
// flight one half's memory-bound kernels run under the other half's transforms (measured on the sharded calls, where tiles alternate
// streams: 12 % on encryption).  Fork / join with events, so the caller's stream semantics are unchanged (and it is graph-capturable).
template <class F>
static int run_split(nttb200_bfv *b, unsigned batch, cudaStream_t st, F part)
{
    const bool worth = b->split && batch >= 2 && (size_t)batch * b->r * b->n >= ((size_t)1 << 24);
    if (!worth) return part(st, 0u, batch);
    if (!b->st2) {
        NTTB200_CHECK(cudaStreamCreateWithFlags(&b->st2, cudaStreamNonBlocking));
        NTTB200_CHECK(cudaEventCreateWithFlags(&b->ev_fork, cudaEventDisableTiming));
        NTTB200_CHECK(cudaEventCreateWithFlags(&b->ev_join, cudaEventDisableTiming));
    }
    const unsigned h = batch / 2;
    NTTB200_CHECK(cudaEventRecord(b->ev_fork, st));
    NTTB200_CHECK(cudaStreamWaitEvent(b->st2, b->ev_fork, 0));
    NTTB200_TRY(part(st, 0u, h));
    NTTB200_TRY(part(b->st2, h, batch - h));
    NTTB200_CHECK(cudaEventRecord(b->ev_join, b->st2));
    NTTB200_CHECK(cudaStreamWaitEvent(st, b->ev_join, 0));
    return 0;
}

extern "C" {

int nttb200_bfv_create(nttb200_bfv **out, unsigned n, unsigned limbs, const nttb200_u64 *q, const nttb200_u64 *psi_roots, nttb200_u64 t,
                       nttb200_u64 gamma)
{
    if (!out || !q || limbs < 2 || t < 2 || (t & (t - 1)) || gamma < 3) return NTTB200_EINVAL;
    // The scheme as the reference implements it needs: t a power of two below 2^32 (32-bit masks in mod_t / fast_convert, poly_arithmetic.cuh:139,222);
    // every q_i = 1 (mod t) (floor(q_i / t) stands in for floor(q / t) mod q_i, bfv_encryption.cuh:193-212, and the Fermat "inverse" mod t of
    // demo.cu:109 is only an inverse then); gamma an odd prime below 2^62 coprime to every q_i (Fermat inverse demo.cu:110, 2 * gamma_bits <= 124).
    // gamma = 1 (mod t): dec_round (poly_arithmetic.cuh:253-263) omits BEHZ's multiplication by gamma^-1 mod t -- the reference's
    // gamma = 2^61 - 10239 is 1 mod 2^11, so its own scheme is only correct for t <= 2048.
    if (t > (1ull << 32) || gamma >= (1ull << 62) || !(gamma & 1) || !h_is_prime(gamma) || gamma % t != 1) return NTTB200_EINVAL;
    for (unsigned i = 0; i < limbs; i++)
        if (q[i] % t != 1 || q[i] == gamma) return NTTB200_EINVAL;
    nttb200_ctx *ctx = nullptr;
    int rc = nttb200_ctx_create(&ctx, n, limbs, q, psi_roots);
    if (rc) return rc;
    nttb200_bfv *b = new nttb200_bfv();
    b->ctx = ctx; b->n = n; b->r = limbs; b->t = t; b->gamma = gamma;
    const unsigned r = limbs, rp = r - 1;
    b->gamma_bits = (int)(log2((double)gamma) + 1);                                     // demo.cu:100 uses {10, 61}
    b->gamma_div_2 = gamma >> 1;
    b->mu_gamma = (u64)(((u128)1 << (2 * b->gamma_bits)) / gamma);                      // demo.cu:221-226
    std::vector<u64> iql(rp), qdt(r), ptg(rp), ipq(rp), bcm(2 * rp);
    for (unsigned i = 0; i < r; i++) qdt[i] = q[i] / t;                                 // :84-88
    for (unsigned i = 0; i < rp; i++) iql[i] = h_modinv_fermat(q[r - 1] % q[i], q[i]);  // :75-79
    u64 mult_t = 1, mult_g = 1;
    for (unsigned i = 0; i < rp; i++) { mult_t = (u64)((u128)mult_t * q[i] % t); mult_g = (u64)((u128)mult_g * q[i] % gamma); }   // :103-108
    b->neg_inv_t = t - h_modinv_fermat(mult_t, t);                                      // :109 (Fermat on t = 2^k: exact only because q_i = 1 mod t)
    b->neg_inv_gamma = gamma - h_modinv_fermat(mult_g, gamma);                          // :110
    for (unsigned i = 0; i < rp; i++) ptg[i] = (u64)((u128)t * gamma % q[i]);           // :118-123
    for (unsigned i = 0; i < rp; i++) {                                                 // :229-243
        u64 tmp = 1;
        for (unsigned j = 0; j < rp; j++) if (j != i) tmp = (u64)((u128)tmp * q[j] % q[i]);
        ipq[i] = h_modinv_fermat(tmp, q[i]);
    }
    const u64 base[2] = {t, gamma};
    for (unsigned k = 0; k < 2; k++)                                                    // :248-264
        for (unsigned j = 0; j < rp; j++) {
            u64 tmp = 1;
            for (unsigned i = 0; i < rp; i++) if (i != j) tmp = (u64)((u128)tmp * q[i] % base[k]);
            bcm[k * rp + j] = tmp;
        }
    auto up = [](u64 **d, const std::vector<u64> &h) -> int {
        NTTB200_CHECK(cudaMalloc(d, h.size() * 8));
        NTTB200_CHECK(cudaMemcpy(*d, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
        return 0;
    };
    // epilogue flavours (bfv_kernels.cuh): the same predicates the kernels evaluate per limb, checked here for ALL limbs
    b->enc_lazy = true; b->epi_ok = true; b->dec_fast = barrett_is_exact(gamma, b->mu_gamma, b->gamma_bits);
    b->all_exact = true;
    for (unsigned i = 0; i < r; i++) b->all_exact = b->all_exact && barrett_is_exact(q[i], ctx->mu[i], (int)ctx->qbit[i]);
    for (unsigned i = 0; i < rp; i++) {
        const bool exact = barrett_is_exact(q[i], ctx->mu[i], (int)ctx->qbit[i]);
        b->enc_lazy = b->enc_lazy && exact && iql[i] < q[i] && q[r - 1] <= 2 * q[i] && q[i] < (1ull << 60);
        b->epi_ok = b->epi_ok && exact && iql[i] < q[i] && q[i] < (1ull << 60);       // the fused epilogue reduces c_last itself when q_last > 2 q_i
        b->dec_fast = b->dec_fast && exact && ptg[i] < q[i] && ipq[i] < q[i] && bcm[rp + i] < gamma;
    }
    {   // wire format offsets (bfv_kernels.cuh: k_ct_pack)
        std::vector<unsigned> woff(rp);
        unsigned acc = 0;
        for (unsigned i = 0; i < rp; i++) { woff[i] = acc; acc += n / 64 * ctx->qbit[i]; }
        b->half_words = acc;
        std::vector<unsigned> koff(r);
        acc = 0;
        for (unsigned i = 0; i < r; i++) { koff[i] = acc; acc += n / 64 * ctx->qbit[i]; }
        b->key_half_words = acc;
        if (cudaMalloc(&b->word_off, rp * sizeof(unsigned)) != cudaSuccess ||
            cudaMemcpy(b->word_off, woff.data(), rp * sizeof(unsigned), cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMalloc(&b->key_word_off, r * sizeof(unsigned)) != cudaSuccess ||
            cudaMemcpy(b->key_word_off, koff.data(), r * sizeof(unsigned), cudaMemcpyHostToDevice) != cudaSuccess) {
            nttb200_bfv_destroy(b);
            return (int)cudaErrorMemoryAllocation;
        }
    }
    memset(b->salsa_key, 1, 32);                                                        // distributions.cuh:249 generate_random_default
    if (const char *e = getenv("NTTB200_BFV_SPLIT")) b->split = atoi(e) != 0;
    while ((1ull << b->tsh) < t) b->tsh++;
    {   // per-limb constants of the epilogue fused into the last inverse kernel (epi.cuh)
        std::vector<EncEpiLimb> ek(rp);
        for (unsigned i = 0; i < rp; i++) {
            EncEpiLimb &e = ek[i];
            e.q = q[i]; e.twoq = 2 * q[i]; e.inv_q_last = iql[i]; e.inv_q_last_s = shoup_companion(iql[i], q[i]);
            e.qdt = qdt[i]; e.bias = (q[r - 1] >> 1) % q[i] + 3 * q[i]; e.ratio = ratio_of(q[i]); e.reduce_cl = q[r - 1] > 2 * q[i];
        }
        if (cudaMalloc(&b->enc_epi, rp * sizeof(EncEpiLimb)) != cudaSuccess ||
            cudaMemcpy(b->enc_epi, ek.data(), rp * sizeof(EncEpiLimb), cudaMemcpyHostToDevice) != cudaSuccess) {
            nttb200_bfv_destroy(b);
            return (int)cudaErrorMemoryAllocation;
        }
    }
    rc = up(&b->inv_q_last_mod_q, iql); if (!rc) rc = up(&b->qi_div_t, qdt); if (!rc) rc = up(&b->prod_t_gamma_mod_q, ptg);
    if (!rc) rc = up(&b->inv_punctured_q, ipq); if (!rc) rc = up(&b->bcm, bcm);
    if (rc) { nttb200_bfv_destroy(b); return rc; }
    *out = b;
    return 0;
}

void nttb200_bfv_destroy(nttb200_bfv *b)
{
    if (!b) return;
    cudaFree(b->inv_q_last_mod_q); cudaFree(b->qi_div_t); cudaFree(b->prod_t_gamma_mod_q); cudaFree(b->inv_punctured_q); cudaFree(b->bcm);
    if (b->ks) cudaFree(b->ks);
    if (b->es) cudaFree(b->es);
    if (b->pt) cudaFree(b->pt);
    if (b->word_off) cudaFree(b->word_off);
    if (b->key_word_off) cudaFree(b->key_word_off);
    nttb200_host_state_destroy(b->host);
    nttb200_mul_state_destroy(b->mul);
    if (b->st2) cudaStreamDestroy(b->st2);
    if (b->ev_fork) cudaEventDestroy(b->ev_fork);
    if (b->ev_join) cudaEventDestroy(b->ev_join);
    if (b->ub) cudaFree(b->ub);
    if (b->es8) cudaFree(b->es8);
    if (b->enc_epi) cudaFree(b->enc_epi);
    nttb200_shard_state_destroy(b->shard);
    cudaFree(b->sk_l); cudaFree(b->sk_ls); cudaFree(b->pk_l); cudaFree(b->pk_ls);
    nttb200_ctx_destroy(b->ctx);
    delete b;
}

nttb200_ctx *nttb200_bfv_ctx(nttb200_bfv *b) { return b ? b->ctx : nullptr; }

int nttb200_bfv_set_sampling_key(nttb200_bfv *b, const unsigned char key[32])
{
    if (!b || !key) return NTTB200_EINVAL;
    memcpy(b->salsa_key, key, 32);
    return 0;
}
int nttb200_bfv_set_fused_epilogue(nttb200_bfv *b, int enable)
{
    if (!b) return NTTB200_EINVAL;
    b->no_fused_epilogue = !enable;
    return 0;
}

int nttb200_bfv_reserve(nttb200_bfv *b, unsigned batch)
{
    if (!b || !batch) return NTTB200_EINVAL;
    const size_t rn = (size_t)b->r * b->n;
    return ensure_scratch(b, (9 * rn + 4 * (size_t)b->n) * batch, (size_t)2 * b->n * batch);
}

// Copies the keys into the context and builds their Shoup companions (one division per coefficient, once per key).
int nttb200_bfv_load_keys(nttb200_bfv *b, const nttb200_u64 *sk, const nttb200_u64 *pk, void *stream)
{
    if (!b || (!sk && !pk)) return NTTB200_EINVAL;
    const size_t rn = (size_t)b->r * b->n;
    const nttb200_ctx *c = b->ctx;
    cudaStream_t st = (cudaStream_t)stream;
    if (sk) {
        if (!b->sk_l) { NTTB200_CHECK(cudaMalloc(&b->sk_l, rn * 8)); NTTB200_CHECK(cudaMalloc(&b->sk_ls, rn * 8)); }
        NTTB200_CHECK(cudaMemcpyAsync(b->sk_l, sk, rn * 8, cudaMemcpyDeviceToDevice, st));
        k_build_companions<<<grid_for(rn, 256), 256, 0, st>>>(b->sk_l, b->sk_ls, c->q_dev, c->logn, b->r, b->r);
    }
    if (pk) {
        if (!b->pk_l) { NTTB200_CHECK(cudaMalloc(&b->pk_l, 2 * rn * 8)); NTTB200_CHECK(cudaMalloc(&b->pk_ls, 2 * rn * 8)); }
        NTTB200_CHECK(cudaMemcpyAsync(b->pk_l, pk, 2 * rn * 8, cudaMemcpyDeviceToDevice, st));
        k_build_companions<<<grid_for(2 * rn, 256), 256, 0, st>>>(b->pk_l, b->pk_ls, c->q_dev, c->logn, b->r, 2 * b->r);
    }
    KCHECK();
    return 0;
}

int nttb200_bfv_keygen(nttb200_bfv *b, nttb200_u64 *sk, nttb200_u64 *pk, unsigned batch, nttb200_u64 nonce0, void *stream)
{
    if (!b || !sk || !pk || !batch || batch > 65535) return NTTB200_EINVAL;
    NTTB200_TRY(nttb200_bfv_reserve(b, batch));
    const size_t rn = (size_t)b->r * b->n, ks_stride = 9 * rn + 4 * (size_t)b->n;
    return run_split(b, batch, (cudaStream_t)stream, [&](cudaStream_t s, unsigned first, unsigned cnt) {
        Pipe P = pipe_from_bfv(b, s);
        return run_keygen(P, b->ks + (size_t)first * ks_stride, ks_stride, b->es + (size_t)first * b->n, sk + (size_t)first * rn, pk + (size_t)first * 2 * rn, cnt,
                          nonce0 + first);
    });
}

int nttb200_bfv_encrypt(nttb200_bfv *b, nttb200_u64 *c, const nttb200_u64 *pk, int pk_per_item, const nttb200_u64 *m, unsigned batch,
                        nttb200_u64 nonce0, void *stream)
{
    if (!b || !c || !m || !batch || batch > 65535 || (!pk && !b->pk_l)) return NTTB200_EINVAL;
    const bool v2 = !pk && b->ctx->lazy_ok && b->epi_ok && !b->no_fused_epilogue;
    if (!v2) NTTB200_TRY(ensure_scratch(b, 9 * (size_t)b->n * batch, (size_t)2 * b->n * batch));      // encryption draws 9n bytes per item
    const size_t rn = (size_t)b->r * b->n, n = b->n;
    if (v2) { Pipe P = pipe_from_bfv(b, (cudaStream_t)stream); return run_encrypt_v2(b, P, c, m, batch, nonce0); }
    return run_split(b, batch, (cudaStream_t)stream, [&](cudaStream_t s, unsigned first, unsigned cnt) {
        Pipe P = pipe_from_bfv(b, s);
        unsigned char *ks = b->ks + (size_t)first * 9 * n;
        int *es = b->es + (size_t)first * 2 * n;
        if (!pk) return run_encrypt_fused(P, b->ctx->lazy_ok != 0, ks, 9 * n, es, c + (size_t)first * 2 * rn, b->pk_l, b->pk_ls, m + (size_t)first * n, n, b->t, cnt,
                                          nonce0 + first);
        return run_encrypt(P, ks, 9 * n, es, c + (size_t)first * 2 * rn, pk + (pk_per_item ? (size_t)first * 2 * rn : 0), pk_per_item ? 2 * rn : 0,
                           m + (size_t)first * n, n, b->t, cnt, nonce0 + first);
    });
}

int nttb200_bfv_decrypt(nttb200_bfv *b, nttb200_u64 *m_out, nttb200_u64 *c, const nttb200_u64 *sk, int sk_per_item, unsigned batch, void *stream)
{
    if (!b || !c || !m_out || !batch || batch > 65535 || (!sk && !b->sk_l)) return NTTB200_EINVAL;
    const size_t rn = (size_t)b->r * b->n, n = b->n;
    DecryptConsts D{b->t, b->gamma, b->mu_gamma, b->gamma_div_2, b->neg_inv_t, b->neg_inv_gamma, b->gamma_bits, b->r - 1, b->bcm};
    return run_split(b, batch, (cudaStream_t)stream, [&](cudaStream_t s, unsigned first, unsigned cnt) {
        Pipe P = pipe_from_bfv(b, s);
        if (!sk) return run_decrypt_fused(P, b->ctx->lazy_ok != 0, c + (size_t)first * 2 * rn, b->sk_l, b->sk_ls, m_out + (size_t)first * n, n, D, cnt);
        return run_decrypt(P, c + (size_t)first * 2 * rn, sk + (sk_per_item ? (size_t)first * rn : 0), sk_per_item ? rn : 0, m_out + (size_t)first * n, n, D, cnt);
    });
}

// ---- homomorphic add and plaintext multiply (the reference stops at decryption) ------------------------------------
// c_a <- c_a + c_b: Dec(result) = m_a + m_b mod t.  Ciphertexts in the reference layout c[batch][2][r][n]; the padding limb is left alone.
int nttb200_bfv_add(nttb200_bfv *b, nttb200_u64 *c_a, const nttb200_u64 *c_b, unsigned batch, void *stream)
{
    if (!b || !c_a || !c_b || !batch || 2 * (size_t)batch > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r;
    k_ct_add<<<pair_grid(n, r - 1, 2 * batch), pair_block(n, r - 1, 2 * batch), 0, (cudaStream_t)stream>>>(c_a, c_b, n, r, batch, b->ctx->q_dev);
    KCHECK();
    return 0;
}
// c <- c + m for a plaintext polynomial m (one per item, or one for the whole batch): Dec(result) = m_c + m mod t.
int nttb200_bfv_add_plain(nttb200_bfv *b, nttb200_u64 *c, const nttb200_u64 *m_poly, int plain_per_item, unsigned batch, void *stream)
{
    if (!b || !c || !m_poly || !batch || batch > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r;
    k_ct_add_plain<<<pair_grid(n, r - 1, batch), pair_block(n, r - 1, batch), 0, (cudaStream_t)stream>>>(c, m_poly, plain_per_item ? (size_t)n : 0, n, r,
                                                                                                 batch, b->t, b->ctx->q_dev, b->qi_div_t);
    KCHECK();
    return 0;
}
// c <- c * p for a plaintext polynomial p (coefficients taken mod t, centred lift): Dec(result) = m * p mod (X^n + 1, t).
// One plaintext for the whole batch (plain_per_item = 0) or one per item.  Per call: lift, NTT of the plaintext limbs, NTT of both
// ciphertext halves, the coefficient-wise product fused into the first inverse kernel, strided inverse pass.
int nttb200_bfv_mul_plain(nttb200_bfv *b, nttb200_u64 *c, const nttb200_u64 *p_poly, int plain_per_item, unsigned batch, void *stream)
{
    if (!b || !c || !p_poly || !batch || batch > 32767) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r, rp = r - 1;
    const size_t rn = (size_t)r * n;
    const unsigned items = plain_per_item ? batch : 1;
    const size_t need = (size_t)items * rp * n;
    if (b->pt_count < need) {
        if (b->pt) cudaFree(b->pt);
        b->pt = nullptr; b->pt_count = 0;
        NTTB200_CHECK(cudaMalloc(&b->pt, need * 8));
        b->pt_count = need;
    }
    Pipe P = pipe_from_bfv(b, (cudaStream_t)stream);
    k_plain_lift<<<pair_grid(n, items, 1), pair_block(n, items, 1), 0, P.st>>>(p_poly, (size_t)n, b->pt, n, rp, items, b->t, b->ctx->q_dev);
    KCHECK();
    NTTB200_TRY(pipe_ntt(P, false, b->pt, items * rp, rp, 0, 0));
    NTTB200_TRY(pipe_ntt(P, false, c, batch * 2 * rp, rp, rp, rn));              // groups = (item, half): rp limbs every r*n
    const bool lazy = P.policy_inv == kPolicyShoupLazy;
    if (!plain_per_item) {
        NTTB200_TRY(launch_polymul(lazy, P.logn, pipe_args(P, false, c, batch * 2 * rp, rp, rp, rn), P.psiinv, P.psiinv_s, b->pt, rp, 0, false,
                                   nullptr, P.st));
    } else {
        for (unsigned h = 0; h < 2; h++)
            NTTB200_TRY(launch_polymul(lazy, P.logn, pipe_args(P, false, c + h * rn, batch * rp, rp, rp, 2 * rn), P.psiinv, P.psiinv_s, b->pt, rp,
                                       (size_t)rp * n, false, nullptr, P.st));
    }
    NTTB200_TRY(pipe_ntt_pass(P, true, 1, c, batch * 2 * rp, rp, rp, rn));
    return 0;
}

// ---- compact wire format --------------------------------------------------------------------------------------------
size_t nttb200_bfv_packed_words(const nttb200_bfv *b) { return b ? 2 * (size_t)b->half_words : 0; }
int nttb200_bfv_pack(nttb200_bfv *b, nttb200_u64 *packed, const nttb200_u64 *c, unsigned batch, void *stream)
{
    if (!b || !packed || !c || !batch || 2 * (size_t)batch > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r;
    const unsigned x = (n / 64 + 127) / 128;
    k_ct_pack<<<dim3(x, r - 1, 2 * batch), 128, 0, (cudaStream_t)stream>>>(c, packed, n, r, batch, b->ctx->qbit_dev, b->word_off, b->half_words);
    KCHECK();
    return 0;
}
int nttb200_bfv_unpack(nttb200_bfv *b, nttb200_u64 *c, const nttb200_u64 *packed, unsigned batch, void *stream)
{
    if (!b || !packed || !c || !batch || 2 * (size_t)batch > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r;
    const size_t rn = (size_t)r * n;
    for (unsigned h = 0; h < 2; h++)      // the padding limb of the reference layout is not stored: zero it
        NTTB200_CHECK(cudaMemset2DAsync(c + h * rn + (size_t)(r - 1) * n, 2 * rn * 8, 0, (size_t)n * 8, batch, (cudaStream_t)stream));
    const unsigned x = (n / 64 + 127) / 128;
    k_ct_unpack<<<dim3(x, r - 1, 2 * batch), 128, 0, (cudaStream_t)stream>>>(packed, c, n, r, batch, b->ctx->qbit_dev, b->word_off, b->half_words);
    KCHECK();
    return 0;
}

// keys: sk[batch][r][n] (polys = r) or pk[batch][2][r][n] (polys = 2r), all r limbs packed
int nttb200_bfv_pack_key(nttb200_bfv *b, nttb200_u64 *packed, const nttb200_u64 *key, unsigned polys, unsigned batch, void *stream)
{
    if (!b || !packed || !key || !batch || (polys != b->r && polys != 2 * b->r) || (size_t)batch * 2 > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r, groups = polys / r * batch;
    const unsigned x = (n / 64 + 127) / 128;
    // the kernel packs limbs [0, gridDim.y) of groups of `r + 1` polynomials when told r + 1: here every limb of a group of r is stored
    k_ct_pack<<<dim3(x, r, groups), 128, 0, (cudaStream_t)stream>>>(key, packed, n, r, batch, b->ctx->qbit_dev, b->key_word_off, b->key_half_words);
    KCHECK();
    return 0;
}
int nttb200_bfv_unpack_key(nttb200_bfv *b, nttb200_u64 *key, const nttb200_u64 *packed, unsigned polys, unsigned batch, void *stream)
{
    if (!b || !packed || !key || !batch || (polys != b->r && polys != 2 * b->r) || (size_t)batch * 2 > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r, groups = polys / r * batch;
    const unsigned x = (n / 64 + 127) / 128;
    k_ct_unpack<<<dim3(x, r, groups), 128, 0, (cudaStream_t)stream>>>(packed, key, n, r, batch, b->ctx->qbit_dev, b->key_word_off, b->key_half_words);
    KCHECK();
    return 0;
}
// device ciphertexts <-> packed HOST bytes (synchronous; a staging buffer is allocated per call: not a hot-path entry)
int nttb200_bfv_pack_host(nttb200_bfv *b, nttb200_u64 *packed_host, const nttb200_u64 *c, unsigned batch, void *stream)
{
    if (!b || !packed_host || !c || !batch) return NTTB200_EINVAL;
    const size_t words = (size_t)batch * 2 * b->half_words;
    u64 *tmp = nullptr;
    NTTB200_CHECK(cudaMalloc(&tmp, words * 8));
    int rc = nttb200_bfv_pack(b, tmp, c, batch, stream);
    if (!rc) rc = (int)cudaMemcpyAsync(packed_host, tmp, words * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (!rc) rc = (int)cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(tmp);
    return rc;
}
int nttb200_bfv_unpack_host(nttb200_bfv *b, nttb200_u64 *c, const nttb200_u64 *packed_host, unsigned batch, void *stream)
{
    if (!b || !packed_host || !c || !batch) return NTTB200_EINVAL;
    const size_t words = (size_t)batch * 2 * b->half_words;
    u64 *tmp = nullptr;
    NTTB200_CHECK(cudaMalloc(&tmp, words * 8));
    int rc = (int)cudaMemcpyAsync(tmp, packed_host, words * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (!rc) rc = nttb200_bfv_unpack(b, c, tmp, batch, stream);
    if (!rc) rc = (int)cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(tmp);
    return rc;
}

// ---- limb-sharded decryption: this GPU's share up to the cross-limb reduction, and the step after the all-reduce -------
// c_shard: items of [2][count][n] holding limbs [first, first+count) of c0 then of c1; sk_shard[count][n] (or per item).
int nttb200_bfv_decrypt_partial(nttb200_bfv *b, nttb200_u64 *partial, nttb200_u64 *c_shard, const nttb200_u64 *sk_shard, int sk_per_item,
                                unsigned first_limb, unsigned limb_count, unsigned shard_half_limbs, unsigned batch, void *stream)
{
    if (!b || !partial || !c_shard || !sk_shard || !batch || !limb_count || first_limb + limb_count > b->r - 1) return NTTB200_EINVAL;
    if (shard_half_limbs == 0) shard_half_limbs = limb_count;
    if (shard_half_limbs < limb_count || batch > 65535) return NTTB200_EINVAL;
    const unsigned n = b->n;
    const nttb200_ctx *c = b->ctx;
    Pipe P = pipe_from_bfv(b, (cudaStream_t)stream);
    // the transform sees only this shard: tables / constants start at the first owned limb, local limb = poly % count
    const size_t toff = (size_t)first_limb * n;
    P.psi = c->psi + toff; P.psiinv = c->psiinv + toff; P.psi_s = c->psi_s + toff; P.psiinv_s = c->psiinv_s + toff; P.lc = c->lc + first_limb;
    LimbArrays Lloc{c->q_dev + first_limb, c->mu_dev + first_limb, c->qbit_dev + first_limb, nullptr, nullptr, nullptr};
    const LimbArrays Lglob = P.L;
    P.L = Lloc;
    const size_t item = (size_t)2 * shard_half_limbs * n, c1_off = (size_t)shard_half_limbs * n;
    NTTB200_TRY(pipe_ntt(P, false, c_shard + c1_off, batch * limb_count, limb_count, limb_count, item));
    k_decrypt_mul<<<pair_grid(n, limb_count, batch), pair_block(n, limb_count, batch), 0, P.st>>>(c_shard, item, c1_off, sk_shard,
                                                                                 sk_per_item ? (size_t)limb_count * n : 0, n, limb_count, batch, Lloc);
    KCHECK();
    NTTB200_TRY(pipe_ntt(P, true, c_shard + c1_off, batch * limb_count, limb_count, limb_count, item));
    DecryptConsts D{b->t, b->gamma, b->mu_gamma, b->gamma_div_2, b->neg_inv_t, b->neg_inv_gamma, b->gamma_bits, b->r - 1, b->bcm};
    k_decrypt_partial<false><<<pair_grid(n, batch, 1), pair_block(n, batch, 1), 0, P.st>>>(c_shard, item, c1_off, partial, n, batch, first_limb, limb_count, D, Lglob);
    KCHECK();
    return 0;
}
int nttb200_bfv_decrypt_finish(nttb200_bfv *b, nttb200_u64 *m_out, const nttb200_u64 *partial_sum, unsigned batch, void *stream)
{
    if (!b || !m_out || !partial_sum || !batch) return NTTB200_EINVAL;
    DecryptConsts D{b->t, b->gamma, b->mu_gamma, b->gamma_div_2, b->neg_inv_t, b->neg_inv_gamma, b->gamma_bits, b->r - 1, b->bcm};
    k_decrypt_finish<false, false><<<pair_grid(b->n, batch, 1), pair_block(b->n, batch, 1), 0, (cudaStream_t)stream>>>(partial_sum, m_out, b->n, b->n, batch, D, 1, 0);
    KCHECK();
    return 0;
}

// ---- the reference's single-item calls (stateless; constants come from the caller's device arrays) -------------------
static Pipe pipe_ref(unsigned n, unsigned r, const nttb200_u64 *psi, const nttb200_u64 *psiinv, const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev,
                     const unsigned *qbit_dev, const nttb200_u64 *inv_q_last, const nttb200_u64 *inv_punct, const nttb200_u64 *ptg,
                     const nttb200_u64 *qi_div_t, cudaStream_t st)
{
    Pipe P;
    P.n = n; P.logn = ilog2u(n); P.r = r;
    P.L = LimbArrays{q_dev, mu_dev, qbit_dev, inv_q_last, inv_punct, ptg};
    P.qi_div_t = qi_div_t;
    P.policy_fwd = P.policy_inv = kPolicyBarrett;
    P.psi = psi; P.psiinv = psiinv; P.psi_s = P.psiinv_s = nullptr; P.lc = nullptr;
    P.use_tma = get_tma_default(); P.st = st;
    P.key = default_key();
    return P;
}

// keygen_rns bfv_keygen.cuh:95.  `in`: 9*r*n + 4*n bytes; `temp`: r*n u64 (its first n*4 bytes receive the gaussian draws).
int nttb200_ref_keygen_rns(unsigned char *in, unsigned q_amount, unsigned n, nttb200_u64 *secret_key, nttb200_u64 *public_key, nttb200_u64 *temp,
                           const nttb200_u64 *psi_table, const nttb200_u64 *psiinv_table, const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev,
                           const unsigned *qbit_dev, void *stream)
{
    if (!in || !secret_key || !public_key || !temp || !q_amount || (n & (n - 1))) return NTTB200_EINVAL;
    Pipe P = pipe_ref(n, q_amount, psi_table, psiinv_table, q_dev, mu_dev, qbit_dev, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream);
    return run_keygen(P, in, 0, reinterpret_cast<int *>(temp), secret_key, public_key, 1, 0);
}
// encryption_rns bfv_encryption.cuh:223.  `in`: 9*n bytes; `e`: 2*r*n u64 (its first 2*n*4 bytes receive the draws).
int nttb200_ref_encryption_rns(nttb200_u64 *c, const nttb200_u64 *public_key, unsigned char *in, nttb200_u64 *e, unsigned n,
                               const nttb200_u64 *psi_table, const nttb200_u64 *psiinv_table, const nttb200_u64 *m_poly,
                               const nttb200_u64 *qi_div_t_dev, nttb200_u64 t, unsigned q_amount, const nttb200_u64 *q_dev,
                               const nttb200_u64 *mu_dev, const unsigned *qbit_dev, const nttb200_u64 *inv_q_last_mod_q_dev, void *stream)
{
    if (!c || !public_key || !in || !e || !m_poly || q_amount < 2 || (n & (n - 1))) return NTTB200_EINVAL;
    Pipe P = pipe_ref(n, q_amount, psi_table, psiinv_table, q_dev, mu_dev, qbit_dev, inv_q_last_mod_q_dev, nullptr, nullptr, qi_div_t_dev,
                      (cudaStream_t)stream);
    return run_encrypt(P, in, 0, reinterpret_cast<int *>(e), c, public_key, 0, m_poly, 0, t, 1, 0);
}
// decryption_rns bfv_decryption.cuh:76.  q_amount = limbs after the drop; plaintext lands at c + n*(q_amount-1) (demo.cu:299).
int nttb200_ref_decryption_rns(nttb200_u64 *c, const nttb200_u64 *secret_key, const nttb200_u64 *psi_table, const nttb200_u64 *psiinv_table,
                               unsigned n, unsigned q_amount, const nttb200_u64 *base_change_matrix_dev, nttb200_u64 t, nttb200_u64 gamma,
                               nttb200_u64 mu_gamma, int gamma_bits, nttb200_u64 neg_inv_t, nttb200_u64 neg_inv_gamma, nttb200_u64 gamma_div_2,
                               const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev,
                               const nttb200_u64 *inv_punctured_q_dev, const nttb200_u64 *prod_t_gamma_mod_q_dev, void *stream)
{
    if (!c || !secret_key || !q_amount || (n & (n - 1))) return NTTB200_EINVAL;
    Pipe P = pipe_ref(n, q_amount + 1, psi_table, psiinv_table, q_dev, mu_dev, qbit_dev, nullptr, inv_punctured_q_dev, prod_t_gamma_mod_q_dev,
                      nullptr, (cudaStream_t)stream);
    DecryptConsts D{t, gamma, mu_gamma, gamma_div_2, neg_inv_t, neg_inv_gamma, gamma_bits, q_amount, base_change_matrix_dev};
    return run_decrypt(P, c, secret_key, 0, c + (size_t)n * (q_amount - 1), 0, D, 1);
}

}  // extern "C"
