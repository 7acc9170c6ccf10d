// bfv_mul.cu -- BFV ciphertext x ciphertext multiplication and relinearisation on the batched NTT core, plus the
// NTT-prime search as a C entry point.  The reference stops at decryption (the paper names homomorphic multiplication as future
// work, Article.pdf p.29); the algorithm is the RNS variant of Halevi-Polyakov-Shoup (see mul_kernels.cuh), every constant derived
// here with exact multi-word integer arithmetic.
//
//   tensor:  ct_a, ct_b (base Q = the rp limbs below the dropped one) --base extension--> Q u P (P: rp + 1 fresh NTT primes)
//            --NTT, d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1, INTT--> d (integers below n Q^2, exact in Q u P)
//            --round(t/Q d) in base P--> --base conversion P -> Q--> y[3][rp][n]
//   relin:   y2 = sum_i [y2]_{q_i} g_i  (g_i = 1 mod q_i, 0 mod q_j):  c = (y0, y1) + sum_i [y2]_{q_i} * evk_i,
//            evk_i = (-(a_i s + e_i) + g_i s^2, a_i)
#include "bfv_internal.h"
#include "mul_kernels.cuh"
#include "table_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

using namespace nttb200;
typedef unsigned __int128 u128;

namespace {
// ---- exact host arithmetic ----------------------------------------------------------------------------------------------------------
u64 mulmod(u64 a, u64 b, u64 m) { return (u64)((u128)a * b % m); }
u64 powmod(u64 a, u64 e, u64 m)
{
    u64 r = 1 % m; a %= m;
    while (e) { if (e & 1) r = mulmod(r, a, m); a = mulmod(a, a, m); e >>= 1; }
    return r;
}
u64 invmod_prime(u64 a, u64 p) { return powmod(a % p, p - 2, p); }
bool is_prime64(u64 m)
{
    if (m < 2) return false;
    for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull})
        if (m % p == 0) return m == p;
    u64 d = m - 1; int s = 0;
    while (!(d & 1)) { d >>= 1; s++; }
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        u64 x = powmod(a, d, m);
        if (x == 1 || x == m - 1) continue;
        bool comp = true;
        for (int i = 1; i < s && comp; i++) { x = mulmod(x, x, m); if (x == m - 1) comp = false; }
        if (comp) return false;
    }
    return true;
}
// little-endian multi-word integers
typedef std::vector<u64> Big;
Big big_mul_small(const Big &a, u64 b)
{
    Big r(a.size() + 1, 0);
    u128 carry = 0;
    for (size_t i = 0; i < a.size(); i++) { u128 t = (u128)a[i] * b + carry; r[i] = (u64)t; carry = t >> 64; }
    r[a.size()] = (u64)carry;
    while (r.size() > 1 && r.back() == 0) r.pop_back();
    return r;
}
u64 big_divmod_small(const Big &a, u64 b, Big *quot)
{
    Big q(a.size(), 0);
    u128 rem = 0;
    for (size_t i = a.size(); i-- > 0;) { u128 cur = (rem << 64) | a[i]; q[i] = (u64)(cur / b); rem = cur % b; }
    while (q.size() > 1 && q.back() == 0) q.pop_back();
    if (quot) *quot = q;
    return (u64)rem;
}
u64 big_mod_small(const Big &a, u64 m) { return big_divmod_small(a, m, nullptr); }
ShoupC shoupc(u64 c, u64 q) { ShoupC s; s.c = c; s.cs = (u64)(((u128)c << 64) / q); return s; }
ModC modc(u64 q)
{
    ModC m;
    m.q = q; m.ratio = ~0ull / q;
    m.qbit = (u32)(log2((double)q) + 1);
    m.mu = (u64)(((u128)1 << (2 * m.qbit)) / q);
    m.pad = 0;
    return m;
}
template <class T> int upload(T **d, const std::vector<T> &h)
{
    NTTB200_CHECK(cudaMalloc(d, h.size() * sizeof(T)));
    NTTB200_CHECK(cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}
}  // namespace

// state of the multiplication entry points, owned by the BFV context
struct nttb200_mul_state {
    unsigned rp = 0, k = 0;
    nttb200_ctx *ctxP = nullptr;            // base P alone (its tables are the source of ctxQP's second part)
    nttb200_ctx *ctxQP = nullptr;           // tables of Q's r - 1 limbs followed by P's: one launch transforms a whole [..][rp + k][n] array
    std::vector<u64> p, psi_p;
    // device constants
    ModC *modQ = nullptr, *modP = nullptr, *modQP = nullptr;
    ShoupC *preQ = nullptr, *preP = nullptr, *preQs = nullptr, *r64Q = nullptr, *r64P = nullptr;
    SplitC *M_QP = nullptr, *M_PQ = nullptr, *W = nullptr, *lam = nullptr;
    u64 *corr_QP = nullptr, *corr_PQ = nullptr;
    unsigned h = 0;                          // split position of the lazy sums (mul_kernels.cuh)
    double *binvQ = nullptr, *binvP = nullptr, *theta = nullptr;
    // relinearisation key evk[rp][2][rp][n] (NTT domain, canonical; no companions: k_relin_accum multiplies by split words)
    u64 *evk = nullptr;
    // grow-only work buffers
    u64 *buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
};
enum { kWA = 0, kWB /* unused since the operands share kWA */, kD, kYP, kY, kDd };

void nttb200_mul_state_destroy(nttb200_mul_state *s)
{
    if (!s) return;
    cudaFree(s->modQ); cudaFree(s->modP); cudaFree(s->modQP); cudaFree(s->preQ); cudaFree(s->M_QP); cudaFree(s->preP); cudaFree(s->M_PQ);
    cudaFree(s->preQs); cudaFree(s->W); cudaFree(s->lam); cudaFree(s->corr_QP); cudaFree(s->corr_PQ); cudaFree(s->binvQ); cudaFree(s->binvP);
    cudaFree(s->theta); cudaFree(s->evk); cudaFree(s->r64Q); cudaFree(s->r64P);
    for (auto b : s->buf) cudaFree(b);
    if (s->ctxP) nttb200_ctx_destroy(s->ctxP);
    if (s->ctxQP) nttb200_ctx_destroy(s->ctxQP);
    delete s;
}

// a context whose limbs are the first `la` limbs of `a` followed by all limbs of `b` (device-to-device copies of the tables)
static int ctx_concat(nttb200_ctx **out, const nttb200_ctx *a, unsigned la, const nttb200_ctx *b)
{
    nttb200_ctx *c = new nttb200_ctx();
    *out = c;
    c->n = a->n; c->logn = a->logn; c->limbs = la + b->limbs; c->device = a->device; c->use_tma = a->use_tma;
    c->lazy_ok = a->lazy_ok && b->lazy_ok;
    const size_t ta = (size_t)la * a->n * 8, tb = (size_t)b->limbs * b->n * 8;
    u64 **dst[4] = {&c->psi, &c->psiinv, &c->psi_s, &c->psiinv_s};
    u64 *const srca[4] = {a->psi, a->psiinv, a->psi_s, a->psiinv_s}, *const srcb[4] = {b->psi, b->psiinv, b->psi_s, b->psiinv_s};
    for (int i = 0; i < 4; i++) {
        NTTB200_CHECK(cudaMalloc(dst[i], ta + tb));
        NTTB200_CHECK(cudaMemcpy(*dst[i], srca[i], ta, cudaMemcpyDeviceToDevice));
        NTTB200_CHECK(cudaMemcpy(reinterpret_cast<char *>(*dst[i]) + ta, srcb[i], tb, cudaMemcpyDeviceToDevice));
    }
    NTTB200_CHECK(cudaMalloc(&c->lc, c->limbs * sizeof(LimbConst)));
    NTTB200_CHECK(cudaMemcpy(c->lc, a->lc, la * sizeof(LimbConst), cudaMemcpyDeviceToDevice));
    NTTB200_CHECK(cudaMemcpy(c->lc + la, b->lc, b->limbs * sizeof(LimbConst), cudaMemcpyDeviceToDevice));
    c->q.assign(a->q.begin(), a->q.begin() + la); c->q.insert(c->q.end(), b->q.begin(), b->q.end());
    return 0;
}

static int find_primes(unsigned bits, unsigned n, unsigned count, const u64 *exclude, unsigned nexclude, u64 *q_out, u64 *psi_out)
{
    if (bits < 20 || bits > 61 || (n & (n - 1)) || !count || !q_out) return NTTB200_EINVAL;
    const u64 step = 2ull * n;
    u64 q = ((1ull << bits) - 1) / step * step + 1;
    unsigned found = 0;
    while (found < count) {
        if (q < step + 1) return NTTB200_EINVAL;
        bool skip = false;
        for (unsigned i = 0; i < nexclude; i++) skip = skip || exclude[i] == q;
        if (!skip && is_prime64(q)) {
            u64 g = 2, psi;
            for (;; g++) {                                   // smallest g whose (q-1)/2n-th power is a primitive 2n-th root: psi^n = -1
                psi = powmod(g, (q - 1) / step, q);
                if (powmod(psi, n, q) == q - 1) break;
            }
            q_out[found] = q;
            if (psi_out) psi_out[found] = psi;
            found++;
        }
        q -= step;
    }
    return 0;
}

static int mul_state(nttb200_bfv *b, nttb200_mul_state **out)
{
    if (b->mul) { *out = b->mul; return 0; }
    const nttb200_ctx *c = b->ctx;
    const unsigned r = b->r, rp = r - 1, k = rp + 1, n = b->n;
    if (k > (unsigned)kBaseMax || k > NTTB200_MAX_LIMBS) return NTTB200_EINVAL;
    nttb200_mul_state *s = new nttb200_mul_state();
    s->rp = rp; s->k = k;
    unsigned bits = 0;
    for (unsigned i = 0; i < r; i++) bits = std::max(bits, c->qbit[i]);
    // the three partial sums of k_bconv / k_scale (<= k + 1 terms, two products of < 2^(2h) each in the cross sum) must not carry,
    // and neither may relinearisation's sum of r - 1 lazy Shoup products
    s->h = (bits + 1) / 2;
    if ((u128)(2 * (k + 1)) * ((u128)1 << (2 * s->h)) >= ((u128)1 << 64) || (u128)(2 * k) * ((u128)1 << bits) >= ((u128)1 << 64)) {
        delete s; return NTTB200_EINVAL;
    }
    const unsigned h = s->h;
    auto splitc = [h](u64 c) { SplitC r; r.c0 = (u32)(c & ((1ull << h) - 1)); r.c1 = (u32)(c >> h); return r; };
    s->p.resize(k); s->psi_p.resize(k);
    std::vector<u64> excl(c->q.begin(), c->q.end());
    excl.push_back(b->gamma);
    int rc = find_primes(bits, n, k, excl.data(), (unsigned)excl.size(), s->p.data(), s->psi_p.data());
    if (!rc) rc = nttb200_ctx_create(&s->ctxP, n, k, s->p.data(), s->psi_p.data());
    if (!rc) rc = ctx_concat(&s->ctxQP, c, rp, s->ctxP);
    if (rc) { nttb200_mul_state_destroy(s); return rc; }
    const std::vector<u64> &q = c->q;      // first rp entries = base Q
    const std::vector<u64> &p = s->p;
    Big Qb(1, 1), Pb(1, 1);
    for (unsigned i = 0; i < rp; i++) Qb = big_mul_small(Qb, q[i]);
    for (unsigned j = 0; j < k; j++) Pb = big_mul_small(Pb, p[j]);
    std::vector<ModC> modQ(rp), modP(k), modQP(rp + k);
    for (unsigned i = 0; i < rp; i++) modQP[i] = modQ[i] = modc(q[i]);
    for (unsigned j = 0; j < k; j++) modQP[rp + j] = modP[j] = modc(p[j]);
    std::vector<ShoupC> preQ(rp), preP(k), preQs(rp), r64Q(rp), r64P(k);
    std::vector<SplitC> M_QP((size_t)k * rp), M_PQ((size_t)rp * k), W((size_t)k * rp), lam(k);
    std::vector<u64> corr_QP(k), corr_PQ(rp);
    std::vector<double> binvQ(rp), binvP(k), theta(rp);
    std::vector<Big> Qi(rp), Pj(k);        // Q / q_i, P / p_j
    for (unsigned i = 0; i < rp; i++) { big_divmod_small(Qb, q[i], &Qi[i]); binvQ[i] = 1.0 / (double)q[i]; }
    for (unsigned j = 0; j < k; j++) { big_divmod_small(Pb, p[j], &Pj[j]); binvP[j] = 1.0 / (double)p[j]; }
    const Big tP = big_mul_small(Pb, b->t);
    for (unsigned i = 0; i < rp; i++) {
        const u64 qi_inv = invmod_prime(big_mod_small(Qi[i], q[i]), q[i]);                    // (Q/q_i)^-1 mod q_i
        preQ[i] = shoupc(qi_inv, q[i]);
        preQs[i] = shoupc(mulmod(qi_inv, invmod_prime(big_mod_small(Pb, q[i]), q[i]), q[i]), q[i]);   // (QP/q_i)^-1 mod q_i
        corr_PQ[i] = big_mod_small(Pb, q[i]);
        r64Q[i] = shoupc((u64)(((u128)1 << 64) % q[i]), q[i]);
        Big omega;
        const u64 rem = big_divmod_small(tP, q[i], &omega);                                  // t P / q_i = omega + rem / q_i
        theta[i] = (double)rem / (double)q[i];
        for (unsigned j = 0; j < k; j++) {
            M_QP[(size_t)j * rp + i] = splitc(big_mod_small(Qi[i], p[j]));
            W[(size_t)j * rp + i] = splitc(big_mod_small(omega, p[j]));
            M_PQ[(size_t)i * k + j] = splitc(big_mod_small(Pj[j], q[i]));
        }
    }
    for (unsigned j = 0; j < k; j++) {
        preP[j] = shoupc(invmod_prime(big_mod_small(Pj[j], p[j]), p[j]), p[j]);
        corr_QP[j] = big_mod_small(Qb, p[j]);
        lam[j] = splitc(mulmod(b->t % p[j], invmod_prime(big_mod_small(Qb, p[j]), p[j]), p[j]));         // t Q^-1 mod p_j
        r64P[j] = shoupc((u64)(((u128)1 << 64) % p[j]), p[j]);
    }
    rc = upload(&s->modQ, modQ); if (!rc) rc = upload(&s->modP, modP); if (!rc) rc = upload(&s->modQP, modQP);
    if (!rc) rc = upload(&s->preQ, preQ); if (!rc) rc = upload(&s->M_QP, M_QP); if (!rc) rc = upload(&s->preP, preP);
    if (!rc) rc = upload(&s->M_PQ, M_PQ); if (!rc) rc = upload(&s->preQs, preQs); if (!rc) rc = upload(&s->W, W); if (!rc) rc = upload(&s->lam, lam);
    if (!rc) rc = upload(&s->corr_QP, corr_QP); if (!rc) rc = upload(&s->corr_PQ, corr_PQ);
    if (!rc) rc = upload(&s->r64Q, r64Q); if (!rc) rc = upload(&s->r64P, r64P);
    if (!rc) rc = upload(&s->binvQ, binvQ); if (!rc) rc = upload(&s->binvP, binvP); if (!rc) rc = upload(&s->theta, theta);
    if (rc) { nttb200_mul_state_destroy(s); return rc; }
    b->mul = s;
    *out = s;
    return 0;
}
static int mul_buf(nttb200_mul_state *s, int which, size_t words, u64 **p)
{
    if (s->cap[which] < words) {
        if (s->buf[which]) NTTB200_CHECK(cudaFree(s->buf[which]));
        s->buf[which] = nullptr; s->cap[which] = 0;
        NTTB200_CHECK(cudaMalloc(&s->buf[which], words * 8));
        s->cap[which] = words;
    }
    *p = s->buf[which];
    return 0;
}
// transform of `num` polynomials grouped as [group][group_polys][n] every group_stride words, polynomial p modulo limb p % division
static int ntt_call(const nttb200_ctx *c, bool inverse, u64 *a, unsigned num, unsigned division, unsigned group_polys, size_t group_stride,
                    cudaStream_t st)
{
    NttArgsHost h{a, inverse ? c->psiinv : c->psi, inverse ? c->psiinv_s : c->psi_s, c->lc, nullptr, nullptr, nullptr, 0, 0, 0, num, division,
                  c->use_tma, group_polys, group_stride};
    return launch_ntt(inverse, c->lazy_ok ? kPolicyShoupLazy : kPolicyShoup, c->logn, h, st);
}
static dim3 grid1(unsigned n, unsigned y) { return dim3((n + 127) / 128, y); }    // one coefficient per thread
static dim3 pair_grid3(unsigned n, unsigned y, unsigned z)
{
    unsigned x = (n + 511) / 512;
    const unsigned long long yz = (unsigned long long)y * z;
    const unsigned cap = 148 * 8;
    if (x * yz > cap) x = (unsigned)std::max<unsigned long long>(1, cap / yz);
    return dim3(x, y, z);
}

// work areas of one (part of a) batch, carved from the state's grow-only buffers
struct MulWork { u64 *WA, *D, *YP, *Y, *Dd; };
static int mul_reserve(nttb200_bfv *b, nttb200_mul_state *s, unsigned batch, bool square, bool want_y, bool want_relin)
{
    const unsigned n = b->n, rp = s->rp, k = s->k, L = rp + k;
    u64 *p;
    NTTB200_TRY(mul_buf(s, kWA, (size_t)batch * 2 * L * n * (square ? 1 : 2), &p));
    NTTB200_TRY(mul_buf(s, kD, (size_t)batch * 3 * L * n, &p));
    NTTB200_TRY(mul_buf(s, kYP, (size_t)batch * 3 * k * n, &p));
    if (want_y) NTTB200_TRY(mul_buf(s, kY, (size_t)batch * 3 * rp * n, &p));
    if (want_relin) NTTB200_TRY(mul_buf(s, kDd, (size_t)batch * ((size_t)rp * rp + 2 * rp) * n, &p));
    return 0;
}
static MulWork mul_carve(nttb200_bfv *b, nttb200_mul_state *s, unsigned first, bool square)
{
    const size_t n = b->n, rp = s->rp, k = s->k, L = rp + k;
    MulWork w;
    w.WA = s->buf[kWA] ? s->buf[kWA] + (size_t)first * 2 * L * n * (square ? 1 : 2) : nullptr;
    w.D = s->buf[kD] ? s->buf[kD] + (size_t)first * 3 * L * n : nullptr;
    w.YP = s->buf[kYP] ? s->buf[kYP] + (size_t)first * 3 * k * n : nullptr;
    w.Y = s->buf[kY] ? s->buf[kY] + (size_t)first * 3 * rp * n : nullptr;
    w.Dd = s->buf[kDd] ? s->buf[kDd] + (size_t)first * (rp * rp + 2 * rp) * n : nullptr;
    return w;
}
// Two halves of a batch on two streams (as run_split in bfv.cu): the memory-bound kernels of one half (tensor product, digit lift,
// key accumulation) run under the issue-bound transforms and base conversions of the other.
template <class F>
static int mul_split(nttb200_bfv *b, unsigned batch, cudaStream_t st, F part)
{
    if (!b->split || batch < 2) return part(st, 0u, batch);
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

static int run_tensor(nttb200_bfv *b, nttb200_mul_state *s, const MulWork &w, u64 *y, const u64 *ca, const u64 *cb, unsigned batch, cudaStream_t st)
{
    const unsigned n = b->n, r = b->r, rp = s->rp, k = s->k, L = rp + k;
    const size_t Ln = (size_t)L * n;
    const bool square = ca == cb;
    u64 *WA = w.WA, *WB = WA + (size_t)batch * 2 * Ln, *D = w.D, *YP = w.YP;
    for (int op = 0; op < (square ? 1 : 2); op++) {
        u64 *Wx = op == 0 ? WA : WB;
        k_gather_q<<<pair_grid3(n, rp, 2 * batch), 256, 0, st>>>(op == 0 ? ca : cb, Wx, n, r, L);
        BconvArgs A{Wx, Ln, Wx + (size_t)rp * n, Ln, s->preQ, s->modQ, s->binvQ, s->M_QP, s->corr_QP, s->r64P, s->modP, rp, k, n, s->h};
        if (rp <= 16) k_bconv<16><<<grid1(n, 2 * batch), 128, ((size_t)rp * k + 5 * k) * 8, st>>>(A);
        else k_bconv<kBaseMax><<<grid1(n, 2 * batch), 128, ((size_t)rp * k + 5 * k) * 8, st>>>(A);
        KCHECK();
    }
    NTTB200_TRY(ntt_call(s->ctxQP, false, WA, (square ? 2 : 4) * batch * L, L, 0, 0, st));
    k_tensor<<<pair_grid3(n, L, batch), 256, 0, st>>>(WA, square ? WA : WB, D, n, L, s->modQP);
    KCHECK();
    NTTB200_TRY(ntt_call(s->ctxQP, true, D, 3 * batch * L, L, 0, 0, st));
    ScaleArgs S{D, YP, s->preQs, s->modQ, s->modP, s->theta, s->W, s->lam, s->r64P, rp, k, n, s->h};
    if (rp <= 16) k_scale<16><<<grid1(n, 3 * batch), 128, ((size_t)rp * k + 5 * k) * 8, st>>>(S);
    else k_scale<kBaseMax><<<grid1(n, 3 * batch), 128, ((size_t)rp * k + 5 * k) * 8, st>>>(S);
    BconvArgs Bk{YP, (size_t)k * n, y, (size_t)rp * n, s->preP, s->modP, s->binvP, s->M_PQ, s->corr_PQ, s->r64Q, s->modQ, k, rp, n, s->h};
    if (k <= 16) k_bconv<16><<<grid1(n, 3 * batch), 128, ((size_t)rp * k + 5 * rp) * 8, st>>>(Bk);
    else k_bconv<kBaseMax><<<grid1(n, 3 * batch), 128, ((size_t)rp * k + 5 * rp) * 8, st>>>(Bk);
    KCHECK();
    return 0;
}

static int run_relin(nttb200_bfv *b, nttb200_mul_state *s, const MulWork &w, u64 *c_out, const u64 *y, unsigned batch, cudaStream_t st)
{
    if (!s->evk) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r, rp = s->rp;
    u64 *Dd = w.Dd, *acc = Dd + (size_t)batch * rp * rp * n;
    k_relin_lift<<<pair_grid3(n, rp * rp, batch), 256, 0, st>>>(y + (size_t)2 * rp * n, (size_t)3 * rp * n, Dd, n, rp, s->modQ);
    KCHECK();
    NTTB200_TRY(ntt_call(b->ctx, false, Dd, batch * rp * rp, rp, 0, 0, st));
    k_relin_accum<<<dim3((batch + kAccumItems - 1) / kAccumItems, (n + 255) / 256, rp), 128, 0, st>>>(Dd, s->evk, acc, n, rp, batch, s->h,
                                                                                                       s->modQ, s->r64Q);
    KCHECK();
    NTTB200_TRY(ntt_call(b->ctx, true, acc, batch * 2 * rp, rp, 0, 0, st));
    k_relin_add<<<pair_grid3(n, rp, 2 * batch), 256, 0, st>>>(y, (size_t)3 * rp * n, acc, c_out, n, rp, r, s->modQ);
    KCHECK();
    return 0;
}

extern "C" {

int nttb200_find_ntt_primes(unsigned bits, unsigned n, unsigned count, const nttb200_u64 *exclude, unsigned nexclude, nttb200_u64 *q_out,
                            nttb200_u64 *psi_out)
{
    return find_primes(bits, n, count, exclude, nexclude, q_out, psi_out);
}

// evk_i = (-(a_i s + e_i) + g_i s^2, a_i), i < rp, from the secret key sk[r][n] (NTT domain, as nttb200_bfv_keygen leaves it).
// Digit i samples Salsa20 nonce nonce0 + i under the context's sampling key.
int nttb200_bfv_relin_keygen(nttb200_bfv *b, const nttb200_u64 *sk, nttb200_u64 nonce0, void *stream)
{
    if (!b || !sk) return NTTB200_EINVAL;
    nttb200_mul_state *s;
    NTTB200_TRY(mul_state(b, &s));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned n = b->n, rp = s->rp;
    const size_t words = (size_t)rp * 2 * rp * n;
    if (!s->evk) NTTB200_CHECK(cudaMalloc(&s->evk, words * 8));
    const size_t ks_stride = (size_t)rp * n * 8 + (size_t)n * 4;
    unsigned char *ks = nullptr;
    u64 *E = nullptr;
    NTTB200_CHECK(cudaMalloc(&ks, ks_stride * rp));
    if (cudaMalloc(&E, (size_t)rp * rp * n * 8) != cudaSuccess) { cudaFree(ks); return nttb200_trace_error((int)cudaErrorMemoryAllocation, __FILE__, __LINE__); }
    const u64 nblk = ks_stride / 64;
    k_salsa20_keystream<<<grid_for(nblk * rp, 256), 256, 0, st>>>(ks, nblk, (u64)rp, ks_stride, bfv_salsa_key(b), nonce0);
    k_relin_sample<<<pair_grid3(n, rp * rp, 1), 256, 0, st>>>(ks, ks_stride, s->evk, E, n, rp, s->modQ);
    int rc = (int)cudaGetLastError();
    if (!rc) rc = ntt_call(b->ctx, false, E, rp * rp, rp, 0, 0, st);
    if (!rc) {
        k_relin_combine<<<pair_grid3(n, rp * rp, 1), 256, 0, st>>>(s->evk, E, sk, n, rp, s->modQ);
        rc = (int)cudaGetLastError();
    }
    if (!rc) rc = (int)cudaStreamSynchronize(st);
    cudaFree(ks); cudaFree(E);
    return rc;
}
int nttb200_bfv_relin_key(nttb200_bfv *b, const nttb200_u64 **evk, size_t *words)
{
    if (!b || !b->mul || !b->mul->evk) return NTTB200_EINVAL;
    if (evk) *evk = b->mul->evk;
    if (words) *words = (size_t)b->mul->rp * 2 * b->mul->rp * b->n;
    return 0;
}
// y[batch][3][r-1][n] (coefficient domain, canonical): the degree-2 ciphertext round(t/Q * (c_a (x) c_b)); Dec = y0 + y1 s + y2 s^2
int nttb200_bfv_mul_tensor(nttb200_bfv *b, nttb200_u64 *y, const nttb200_u64 *c_a, const nttb200_u64 *c_b, unsigned batch, void *stream)
{
    if (!b || !y || !c_a || !c_b || !batch || batch > 10000) return NTTB200_EINVAL;
    nttb200_mul_state *s;
    NTTB200_TRY(mul_state(b, &s));
    const bool square = c_a == c_b;
    NTTB200_TRY(mul_reserve(b, s, batch, square, false, false));
    const size_t rn = (size_t)b->r * b->n, yn = (size_t)3 * s->rp * b->n;
    return mul_split(b, batch, (cudaStream_t)stream, [&](cudaStream_t st, unsigned first, unsigned cnt) {
        return run_tensor(b, s, mul_carve(b, s, first, square), y + first * yn, c_a + first * 2 * rn, square ? c_a + first * 2 * rn : c_b + first * 2 * rn, cnt, st);
    });
}
int nttb200_bfv_relinearize(nttb200_bfv *b, nttb200_u64 *c_out, const nttb200_u64 *y, unsigned batch, void *stream)
{
    if (!b || !y || !c_out || !batch || batch > 10000 || !b->mul) return NTTB200_EINVAL;
    nttb200_mul_state *s = b->mul;
    if (!s->evk) return NTTB200_EINVAL;
    u64 *p;
    NTTB200_TRY(mul_buf(s, kDd, (size_t)batch * ((size_t)s->rp * s->rp + 2 * s->rp) * b->n, &p));
    const size_t rn = (size_t)b->r * b->n, yn = (size_t)3 * s->rp * b->n;
    return mul_split(b, batch, (cudaStream_t)stream, [&](cudaStream_t st, unsigned first, unsigned cnt) {
        return run_relin(b, s, mul_carve(b, s, first, false), c_out + first * 2 * rn, y + first * yn, cnt, st);
    });
}
// c_out <- relin(c_a * c_b): Dec(c_out) = m_a * m_b mod (X^n + 1, t).  c_out may alias an input.  Needs nttb200_bfv_relin_keygen.
int nttb200_bfv_mul(nttb200_bfv *b, nttb200_u64 *c_out, const nttb200_u64 *c_a, const nttb200_u64 *c_b, unsigned batch, void *stream)
{
    if (!b || !c_out || !c_a || !c_b || !batch || batch > 10000) return NTTB200_EINVAL;
    nttb200_mul_state *s;
    NTTB200_TRY(mul_state(b, &s));
    if (!s->evk) return NTTB200_EINVAL;
    const bool square = c_a == c_b;
    NTTB200_TRY(mul_reserve(b, s, batch, square, true, true));
    const size_t rn = (size_t)b->r * b->n;
    return mul_split(b, batch, (cudaStream_t)stream, [&](cudaStream_t st, unsigned first, unsigned cnt) {
        const MulWork w = mul_carve(b, s, first, square);
        const u64 *a = c_a + first * 2 * rn, *bb = square ? a : c_b + first * 2 * rn;
        NTTB200_TRY(run_tensor(b, s, w, w.Y, a, bb, cnt, st));
        return run_relin(b, s, w, c_out + first * 2 * rn, w.Y, cnt, st);
    });
}
// the auxiliary base of the multiplication (for tests / oracles): k = r primes
int nttb200_bfv_mul_aux_base(nttb200_bfv *b, nttb200_u64 *p_out, unsigned *count)
{
    if (!b) return NTTB200_EINVAL;
    nttb200_mul_state *s;
    NTTB200_TRY(mul_state(b, &s));
    if (count) *count = s->k;
    if (p_out) for (unsigned j = 0; j < s->k; j++) p_out[j] = s->p[j];
    return 0;
}

}  // extern "C"
