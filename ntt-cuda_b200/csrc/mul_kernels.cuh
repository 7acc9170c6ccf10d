// mul_kernels.cuh -- kernels of BFV ciphertext x ciphertext multiplication and relinearisation (the paper's stated
// future work, Article.pdf p.29; the natural next caller of the batched NTT and of half_poly_mul_device, poly_arithmetic.cuh:303).
// The reference has no such operation: the arithmetic follows the published RNS variant of Halevi, Polyakov and Shoup
// ("An Improved RNS Variant of the BFV Homomorphic Encryption Scheme", CT-RSA 2019): base extension Q -> P and P -> Q with a
// floating-point estimate of the CRT overflow, and the "simple scaling" round(t/Q * d) computed limb-wise in the auxiliary base P;
// relinearisation is the RNS-digit key switch (one digit per limb of Q).
// All floating-point steps use explicit round-to-nearest multiplies / adds in a fixed order (no FMA contraction), so the CPU oracle
// (oracle/bfv_mul_oracle.py) reproduces them bit for bit.
#pragma once
#include "modarith.cuh"
#include "bfv_kernels.cuh"

namespace nttb200 {

struct ModC { u64 q, ratio, mu; u32 qbit, pad; };  // modulus, floor(2^64 / q), Barrett mu = floor(2^(2 qbit) / q), bit length
struct ShoupC { u64 c, cs; };                      // constant below q and its companion floor(c * 2^64 / q)

__host__ __device__ __forceinline__ double dmul_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
__host__ __device__ __forceinline__ double dadd_rn(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
// canonical product of two canonical residues, no companion: the reference's Barrett estimate (at most 2 short) + corrections
__host__ __device__ __forceinline__ u64 mulmod_canon(u64 x, u64 y, const ModC &m)
{
    return csub(csub(barrett_lazy(x, y, m.q, m.mu, (int)m.qbit), 2 * m.q), m.q);
}
__host__ __device__ __forceinline__ u64 shoup_canon(u64 x, const ShoupC &c, u64 q) { return csub(shoup_mul(x, c.c, c.cs, q), q); }

constexpr int kBaseMax = 32;                 // limbs of one base

// ---- lazy accumulation of products of residues ---------------------------------------------------------------------------------------
// sum of y_i * c_i over <= kBaseMax + 1 terms, every factor below 2^bits.  Both factors are split at h = ceil(bits / 2) bits -- y in the
// kernel, once per coefficient; constants on the host (SplitC) -- so that each of the four partial products is below 2^(2h) and the
// three partial SUMS (low, cross, high) cannot carry: 2 * terms * 2^(2h) < 2^64 is checked in mul_state.  A term is then four
// 32x32->64 multiply-adds (an exact Shoup product: six wide multiplies, four narrow ones and the carry chain of a mul.hi), and ONE
// reduction per sum replaces one per term.  Exact: the reduced value is the canonical residue of the true sum.
struct SplitC { u32 c0, c1; };               // c = c1 * 2^h + c0
struct Acc3 { u64 ll, mid, hh; };
__host__ __device__ __forceinline__ void acc_init(Acc3 &a, u64 v) { a.ll = v; a.mid = 0; a.hh = 0; }
__host__ __device__ __forceinline__ void acc_mac(Acc3 &a, u32 y0, u32 y1, const SplitC c)
{
#if defined(__CUDA_ARCH__)
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.ll) : "r"(y0), "r"(c.c0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(y0), "r"(c.c1));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.mid) : "r"(y1), "r"(c.c0));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a.hh) : "r"(y1), "r"(c.c1));
#else
    a.ll += (u64)y0 * c.c0;
    a.mid += (u64)y0 * c.c1;
    a.mid += (u64)y1 * c.c0;
    a.hh += (u64)y1 * c.c1;
#endif
}
// value = ll + mid * 2^h + hh * 2^(2h) (below 2^128), reduced with r64 = 2^64 mod q and its companion
__host__ __device__ __forceinline__ u64 acc_reduce(const Acc3 &a, unsigned h, const ModC &m, const ShoupC &r64)
{
    const u64 m_lo = a.mid << h, h_lo = a.hh << (2 * h);
    u64 lo = a.ll + m_lo;
    u64 hi = (a.mid >> (64 - h)) + (a.hh >> (64 - 2 * h)) + (lo < m_lo ? 1u : 0u);
    lo += h_lo;
    hi += lo < h_lo ? 1u : 0u;
    return csub(csub(shoup_mul(hi, r64.c, r64.cs, m.q) + mod_exact(lo, m.q, m.ratio), 2 * m.q), m.q);
}

// ---- fast base conversion with overflow estimate (HPS eq. 2-3) ------------------------------------------------------------------
// x[item][lin][n] (canonical residues mod the input base) -> out[item][lout][n] mod the output base:
//   y_i = [x_i * (B/b_i)^-1]_{b_i};  v = round(sum_i y_i / b_i);  out_j = (sum_i y_i * [B/b_i]_{o_j} - v * [B]_{o_j}) mod o_j
// i.e. the residues of the CENTRED representative of x (or of one shifted by B when the estimate v is off by one, which only happens
// within 2^-48 of the boundary and is harmless: any representative below B in magnitude works downstream).
// One thread per coefficient: the lin residues stay in registers (loops fully unrolled, uniform early exit), the lout x lin matrix
// is staged in shared memory; per output limb one lazy sum.
struct BconvArgs {
    const u64 *x; size_t in_item;            // input, [lin][n] per item
    u64 *out; size_t out_item;               // output, [lout][n] per item
    const ShoupC *pre;                       // [lin]   (B/b_i)^-1 mod b_i
    const ModC *bin;                         // [lin]   input moduli
    const double *binv;                      // [lin]   1 / b_i
    const SplitC *M;                         // [lout][lin]  B/b_i mod o_j
    const u64 *corr;                         // [lout]  B mod o_j
    const ShoupC *r64;                       // [lout]  2^64 mod o_j
    const ModC *bout;                        // [lout]  output moduli
    unsigned lin, lout, n, h;                // h: split position of the lazy sums
};
template <int LMAX>      // register slots for the input residues: 16 (most parameter sets) or kBaseMax
NTT_KERNEL void __launch_bounds__(128) k_bconv(BconvArgs A)
{
    NTT_DYN_SMEM(mul_raw);
    SplitC *mul_smem = reinterpret_cast<SplitC *>(mul_raw);
    u64 *outc = reinterpret_cast<u64 *>(mul_raw) + (size_t)A.lin * A.lout;      // per output limb: q, ratio, corr, r64.c, r64.cs
    for (unsigned i = threadIdx.x; i < A.lin * A.lout; i += blockDim.x) mul_smem[i] = A.M[i];
    for (unsigned i = threadIdx.x; i < A.lout; i += blockDim.x) {
        outc[5 * i] = A.bout[i].q; outc[5 * i + 1] = A.bout[i].ratio; outc[5 * i + 2] = A.corr[i];
        outc[5 * i + 3] = A.r64[i].c; outc[5 * i + 4] = A.r64[i].cs;
    }
    __syncthreads();
    const size_t k = blockIdx.y;
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.n) return;
    const u64 *x = A.x + k * A.in_item + j;
    u64 *o = A.out + k * A.out_item + j;
    u32 y0[LMAX], y1[LMAX];
    const u64 mask = (1ull << A.h) - 1;
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < LMAX; i++) {
        if (i >= (int)A.lin) break;
        const u64 y = shoup_canon(x[(size_t)i * A.n], A.pre[i], A.bin[i].q);
        s = dadd_rn(s, dmul_rn((double)y, A.binv[i]));
        y0[i] = (u32)(y & mask); y1[i] = (u32)(y >> A.h);
    }
    const u64 v = (u64)(long long)floor(dadd_rn(s, 0.5));
    for (unsigned l = 0; l < A.lout; l++) {
        ModC m; m.q = outc[5 * l]; m.ratio = outc[5 * l + 1];
        ShoupC r64; r64.c = outc[5 * l + 3]; r64.cs = outc[5 * l + 4];
        const SplitC *row = mul_smem + (size_t)l * A.lin;
        Acc3 acc;
        acc_init(acc, (u64)(A.lin + 1) * m.q - v * outc[5 * l + 2]);                                      // v <= lin, corr < o_j
#pragma unroll
        for (int i = 0; i < LMAX; i++) {
            if (i >= (int)A.lin) break;
            acc_mac(acc, y0[i], y1[i], row[i]);
        }
        o[(size_t)l * A.n] = acc_reduce(acc, A.h, m, r64);
    }
}

// ---- HPS "simple scaling": y = round(t/Q * d) in base P from d in base Q u P ------------------------------------------------------
// d[item][comp][rp + k][n] (coefficient domain, canonical; Q limbs first) -> y[item][comp][k][n]:
//   yt_i = [d_i * (QP/q_i)^-1]_{q_i};  y_j = ( sum_i yt_i * [omega_i]_{p_j} + d'_j * lambda_j + round(sum_i yt_i * theta_i) ) mod p_j
// with t*P/q_i = omega_i + theta_i (integer + fraction) and lambda_j = [(QP/p_j)^-1 * t * P/p_j]_{p_j}.  Same thread layout as k_bconv.
struct ScaleArgs {
    const u64 *d; u64 *y;
    const ShoupC *preQ;                      // [rp]     (QP/q_i)^-1 mod q_i
    const ModC *modQ, *modP;                 // [rp], [k]
    const double *theta;                     // [rp]     frac(t * P / q_i)
    const SplitC *W;                         // [k][rp]  floor(t * P / q_i) mod p_j
    const SplitC *lam;                       // [k]
    const ShoupC *r64;                       // [k]      2^64 mod p_j
    unsigned rp, k, n, h;
};
template <int LMAX>
NTT_KERNEL void __launch_bounds__(128) k_scale(ScaleArgs A)
{
    NTT_DYN_SMEM(mul_raw);
    SplitC *mul_smem = reinterpret_cast<SplitC *>(mul_raw);
    u64 *outc = reinterpret_cast<u64 *>(mul_raw) + (size_t)A.rp * A.k;          // per output limb: q, ratio, lambda (split), r64.c, r64.cs
    for (unsigned i = threadIdx.x; i < A.rp * A.k; i += blockDim.x) mul_smem[i] = A.W[i];
    for (unsigned i = threadIdx.x; i < A.k; i += blockDim.x) {
        outc[5 * i] = A.modP[i].q; outc[5 * i + 1] = A.modP[i].ratio; outc[5 * i + 2] = (u64)A.lam[i].c0 | ((u64)A.lam[i].c1 << 32);
        outc[5 * i + 3] = A.r64[i].c; outc[5 * i + 4] = A.r64[i].cs;
    }
    __syncthreads();
    const size_t kc = blockIdx.y;             // item * comps + comp
    const size_t L = (size_t)A.rp + A.k;
    const u32 j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= A.n) return;
    const u64 *d = A.d + kc * L * A.n + j;
    u64 *y = A.y + kc * (size_t)A.k * A.n + j;
    u32 y0[LMAX], y1[LMAX];
    const u64 mask = (1ull << A.h) - 1;
    double f = 0.0;
#pragma unroll
    for (int i = 0; i < LMAX; i++) {
        if (i >= (int)A.rp) break;
        const u64 yt = shoup_canon(d[(size_t)i * A.n], A.preQ[i], A.modQ[i].q);
        f = dadd_rn(f, dmul_rn((double)yt, A.theta[i]));
        y0[i] = (u32)(yt & mask); y1[i] = (u32)(yt >> A.h);
    }
    const u64 v = (u64)(long long)floor(dadd_rn(f, 0.5));
    for (unsigned l = 0; l < A.k; l++) {
        ModC m; m.q = outc[5 * l]; m.ratio = outc[5 * l + 1];
        ShoupC r64; r64.c = outc[5 * l + 3]; r64.cs = outc[5 * l + 4];
        SplitC lam; lam.c0 = (u32)outc[5 * l + 2]; lam.c1 = (u32)(outc[5 * l + 2] >> 32);
        const SplitC *row = mul_smem + (size_t)l * A.rp;
        const u64 dp = d[((size_t)A.rp + l) * A.n];
        Acc3 acc;
        acc_init(acc, v);                                                                                 // v <= rp
        acc_mac(acc, (u32)(dp & mask), (u32)(dp >> A.h), lam);
#pragma unroll
        for (int i = 0; i < LMAX; i++) {
            if (i >= (int)A.rp) break;
            acc_mac(acc, y0[i], y1[i], row[i]);
        }
        y[(size_t)l * A.n] = acc_reduce(acc, A.h, m, r64);
    }
}

// ---- tensor product in the NTT domain ---------------------------------------------------------------------------------------------
// a, b: [item][2][L][n] (component, limb); d: [item][3][L][n]:  d0 = a0 b0, d1 = a0 b1 + a1 b0, d2 = a1 b1, limb l modulo mods[l].
// Inputs canonical.  grid (x, L, items)
NTT_KERNEL void k_tensor(const u64 *a, const u64 *b, u64 *d, unsigned n, unsigned L, const ModC *mods)
{
    const unsigned l = blockIdx.y;
    const size_t k = blockIdx.z, Ln = (size_t)L * n;
    const ModC m = mods[l];
    const u64 *a0 = a + k * 2 * Ln + (size_t)l * n, *a1 = a0 + Ln, *b0 = b + k * 2 * Ln + (size_t)l * n, *b1 = b0 + Ln;
    u64 *d0 = d + k * 3 * Ln + (size_t)l * n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 x0 = ld2(a0 + j), x1 = ld2(a1 + j), y0 = ld2(b0 + j), y1 = ld2(b1 + j);
        st2(d0 + j, mulmod_canon(x0.x, y0.x, m), mulmod_canon(x0.y, y0.y, m));
        st2(d0 + Ln + j, csub(mulmod_canon(x0.x, y1.x, m) + mulmod_canon(x1.x, y0.x, m), m.q),
            csub(mulmod_canon(x0.y, y1.y, m) + mulmod_canon(x1.y, y0.y, m), m.q));
        st2(d0 + 2 * Ln + j, mulmod_canon(x1.x, y1.x, m), mulmod_canon(x1.y, y1.y, m));
    }
}

// gathers the rp limbs of both halves of reference-layout ciphertexts c[item][2][r][n] into w[item][2][L][n] slots [0, rp)
// (the P slots [rp, L) are filled by k_bconv).  grid (x, rp, 2 * items)
NTT_KERNEL void k_gather_q(const u64 *c, u64 *w, unsigned n, unsigned r, unsigned L)
{
    const unsigned l = blockIdx.y;
    const size_t kh = blockIdx.z;
    const u64 *src = c + kh * r * n + (size_t)l * n;
    u64 *dst = w + kh * L * n + (size_t)l * n;
    NTT_PAIR_STRIDE(j, n) { const ulonglong2 v = ld2(src + j); st2(dst + j, v.x, v.y); }
}

// ---- relinearisation (RNS-digit key switch) ------------------------------------------------------------------------------------------
// digits: D[item][i][j][n] = [y2_i]_{q_j}, y2[item][rp][n] canonical mod q_i.  grid (x, rp * rp, items)
NTT_KERNEL void k_relin_lift(const u64 *y2, size_t y2_item, u64 *D, unsigned n, unsigned rp, const ModC *modQ)
{
    const unsigned i = blockIdx.y / rp, jl = blockIdx.y % rp;
    const size_t k = blockIdx.z;
    const ModC m = modQ[jl];
    const u64 *src = y2 + k * y2_item + (size_t)i * n;
    u64 *dst = D + ((k * rp + i) * rp + jl) * (size_t)n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 v = ld2(src + j);
        st2(dst + j, mod_exact(v.x, m.q, m.ratio), mod_exact(v.y, m.q, m.ratio));
    }
}
// acc[item][h][j][n] = sum_i D[item][i][j][n] * evk[i][h][j][n] mod q_j (NTT domain, canonical).  grid (ceil(items / kAccumItems), x, rp): the item
// groups of one (coefficient range, limb) are neighbours in launch order, so the key words the first one pulls from HBM are L2 hits for the
// others.  One thread = one PAIR of coefficients of limb j for BOTH halves and kAccumItems items: every digit word is read once (round-2
// first version: once per half, plus a Shoup companion per key word); the loads of digit i + 1 are issued before the products of digit i;
// the products are the lazy split-word sums of k_bconv, one reduction per output word.
constexpr int kAccumItems = 2;
struct AccumLoads { ulonglong2 e0, e1, d[kAccumItems]; };
NTT_KERNEL void __launch_bounds__(128) k_relin_accum(const u64 *D, const u64 *evk, u64 *acc, unsigned n, unsigned rp, unsigned items,
                                                     unsigned h, const ModC *modQ, const ShoupC *r64)
{
    const unsigned jl = blockIdx.z;
    const unsigned k0 = blockIdx.x * kAccumItems;
    const unsigned cnt = items - k0 < (unsigned)kAccumItems ? items - k0 : (unsigned)kAccumItems;
    const ModC m = modQ[jl];
    const ShoupC r = r64[jl];
    const u64 mask = (1ull << h) - 1;
    const size_t limb = (size_t)jl * n, rpn = (size_t)rp * n;
    auto split = [&](u64 v) { SplitC c; c.c0 = (u32)(v & mask); c.c1 = (u32)(v >> h); return c; };
    for (u32 j = 2 * (blockIdx.y * blockDim.x + threadIdx.x); j < n; j += 2 * gridDim.y * blockDim.x) {
        auto fetch = [&](unsigned i) {
            AccumLoads L;
            L.e0 = ld2(evk + ((size_t)i * 2) * rpn + limb + j);
            L.e1 = ld2(evk + ((size_t)i * 2 + 1) * rpn + limb + j);
#pragma unroll
            for (int t = 0; t < kAccumItems; t++) {
                const unsigned it = t < (int)cnt ? k0 + t : k0;               // a short last group re-reads its first item (not stored)
                L.d[t] = ld2(D + ((size_t)it * rp + i) * rpn + limb + j);
            }
            return L;
        };
        Acc3 a[kAccumItems][2][2];                                            // [item][half][coefficient of the pair]
#pragma unroll
        for (int t = 0; t < kAccumItems; t++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc_init(a[t][q >> 1][q & 1], 0);
        AccumLoads cur = fetch(0);
        for (unsigned i = 0; i < rp; i++) {
            const AccumLoads nxt = fetch(i + 1 < rp ? i + 1 : i);
            const SplitC e0x = split(cur.e0.x), e0y = split(cur.e0.y), e1x = split(cur.e1.x), e1y = split(cur.e1.y);
#pragma unroll
            for (int t = 0; t < kAccumItems; t++) {
                const SplitC dx = split(cur.d[t].x), dy = split(cur.d[t].y);
                acc_mac(a[t][0][0], dx.c0, dx.c1, e0x);
                acc_mac(a[t][0][1], dy.c0, dy.c1, e0y);
                acc_mac(a[t][1][0], dx.c0, dx.c1, e1x);
                acc_mac(a[t][1][1], dy.c0, dy.c1, e1y);
            }
            cur = nxt;
        }
#pragma unroll
        for (int t = 0; t < kAccumItems; t++) {
            if (t >= (int)cnt) break;
            st2(acc + ((size_t)(k0 + t) * 2) * rpn + limb + j, acc_reduce(a[t][0][0], h, m, r), acc_reduce(a[t][0][1], h, m, r));
            st2(acc + ((size_t)(k0 + t) * 2 + 1) * rpn + limb + j, acc_reduce(a[t][1][0], h, m, r), acc_reduce(a[t][1][1], h, m, r));
        }
    }
}
// out[item][h][r slots][n] limb j = (y[item][h][rp][n] + acc[item][h][rp][n]) mod q_j (coefficient domain).  grid (x, rp, 2 * items)
NTT_KERNEL void k_relin_add(const u64 *y, size_t y_item, const u64 *acc, u64 *out, unsigned n, unsigned rp, unsigned r, const ModC *modQ)
{
    const unsigned jl = blockIdx.y, h = blockIdx.z & 1u;
    const size_t k = blockIdx.z >> 1;
    const u64 q = modQ[jl].q;
    const u64 *a = y + k * y_item + ((size_t)h * rp + jl) * n, *b = acc + ((k * 2 + h) * rp + jl) * (size_t)n;
    u64 *o = out + ((k * 2 + h) * r + jl) * (size_t)n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 u = ld2(a + j), v = ld2(b + j);
        st2(o + j, csub(u.x + v.x, q), csub(u.y + v.y, q));
    }
}

// ---- relinearisation key generation ---------------------------------------------------------------------------------------------
// evk[i][1][j][n] = a_ij uniform mod q_j (taken as NTT-domain values), from 64-bit keystream words (uniform_value, the reference's
// converter bfv_keygen.cuh:37-44); E[i][j][n] = e_i lifted to limb j, e_i gaussian from 32-bit words.  Keystream of digit i:
// [a words: rp * n u64][e words: n u32].  grid (x, rp * rp)
NTT_KERNEL void k_relin_sample(const unsigned char *ks, size_t ks_stride, u64 *evk, u64 *E, unsigned n, unsigned rp, const ModC *modQ)
{
    const unsigned i = blockIdx.y / rp, jl = blockIdx.y % rp;
    const u64 q = modQ[jl].q;
    const unsigned char *s = ks + (size_t)i * ks_stride;
    const u64 *aw = reinterpret_cast<const u64 *>(s) + (size_t)jl * n;
    const u32 *ew = reinterpret_cast<const u32 *>(s + (size_t)rp * n * 8);
    u64 *a = evk + (((size_t)i * 2 + 1) * rp + jl) * (size_t)n, *e = E + ((size_t)i * rp + jl) * (size_t)n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 w = ld2(aw + j);
        st2(a + j, uniform_value(w.x, q), uniform_value(w.y, q));
        st2(e + j, signed_to_residue(gaussian_value(ew[j]), q), signed_to_residue(gaussian_value(ew[j + 1]), q));
    }
}
// evk[i][0][j] = -(a_ij * s_j + NTT(e_i)_j) + [i == j] * s_j^2   (NTT domain; sk[r][n] canonical).  grid (x, rp * rp)
NTT_KERNEL void k_relin_combine(u64 *evk, const u64 *Ehat, const u64 *sk, unsigned n, unsigned rp, const ModC *modQ)
{
    const unsigned i = blockIdx.y / rp, jl = blockIdx.y % rp;
    const ModC m = modQ[jl];
    const u64 *a = evk + (((size_t)i * 2 + 1) * rp + jl) * (size_t)n, *e = Ehat + ((size_t)i * rp + jl) * (size_t)n, *s = sk + (size_t)jl * n;
    u64 *o = evk + (((size_t)i * 2) * rp + jl) * (size_t)n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 av = ld2(a + j), ev = ld2(e + j), sv = ld2(s + j);
        u64 r0 = csub(mulmod_canon(av.x, sv.x, m) + ev.x, m.q), r1 = csub(mulmod_canon(av.y, sv.y, m) + ev.y, m.q);
        r0 = r0 ? m.q - r0 : 0; r1 = r1 ? m.q - r1 : 0;
        if (i == jl) { r0 = csub(r0 + mulmod_canon(sv.x, sv.x, m), m.q); r1 = csub(r1 + mulmod_canon(sv.y, sv.y, m), m.q); }
        st2(o + j, r0, r1);
    }
}

}  // namespace nttb200
