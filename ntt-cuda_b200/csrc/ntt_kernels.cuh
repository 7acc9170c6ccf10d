// ntt_kernels.cuh -- batched negacyclic NTT / INTT for sm_100a in TWO HBM passes.
//
// Replaces ntt_60bit.cuh:63-265, 388-606 (CTBasedNTTInner[Single][_batch], GSBasedINTTInner[Single][_batch]:
// one butterfly per thread per stage, 4-5 HBM passes at n = 32768) and the launch schedules :267-386, 608-697.
// Same transform, same tables, same layout: natural-order in -> bit-reversed out (forward), the mirror for the
// inverse, stage `length = m` uses table entries [m, 2m)  (psi^bitrev(i), parameter.h:5-12).
//
// Structure (forward; the inverse runs the same two kernels mirrored):
//   pass 1  "strided"  : the first K1 stages.  View the polynomial as [R = 2^K1 rows][n/R cols]; stage j pairs
//                        rows that differ in bit K1-1-j and all columns share the twiddle, so a CTA takes a
//                        [R][16]-column tile (128-byte rows, fetched by ONE 3-D TMA box), and every thread keeps
//                        16 coefficients in registers for 3-4 stages at a time (radix-16 / radix-8x2 rounds),
//                        exchanging through shared memory between rounds.
//   pass 2  "contig"   : the last K2 = 7|8 stages on contiguous 2^K2-element blocks.  A CTA takes 128 rows of
//                        16 coefficients (2-D TMA box, 128-byte swizzle); a block is owned by 8|16 lanes of ONE
//                        warp, so the only synchronisation between its two register rounds is __syncwarp().
// Butterflies are Harvey lazy butterflies in [0,4q) (forward) / [0,2q) (inverse) on Shoup twiddle companions
// (1 mul.hi.u64 + 2 mul.lo.u64 per modular multiplication); the n^-1 scaling of the inverse is folded into its
// last stage instead of one halving per stage (ntt_60bit.cuh:165,175) -- all results are canonical residues,
// hence bit-identical to the reference's.
#pragma once
#include "modarith.cuh"
#include "tile_io.cuh"
#include "epi.cuh"

#ifdef NTTB200_EMU
struct TensorMap { unsigned char opaque[128]; };
void emu_tma_2d(bool load, const TensorMap *m, void *smem, int c0, int c1);
void emu_tma_3d(bool load, const TensorMap *m, void *smem, int c0, int c1, int c2);
void emu_tma_4d(bool load, const TensorMap *m, void *smem, int c0, int c1, int c2, int c3);
#else
#include <cuda.h>
typedef CUtensorMap TensorMap;
#endif

namespace nttb200 {

struct NttArgs {
    u64 *a;               // [num][n] coefficients, in place
    const u64 *tw;        // psi (forward) / psiinv (inverse) tables, [limbs][n]
    const u64 *tws;       // Shoup companions floor(tw * 2^64 / q), [limbs][n]        (ShoupPolicy)
    const LimbConst *lc;  // [limbs]                                                  (ShoupPolicy)
    const u64 *qv;        // BarrettPolicy: device arrays q / mu / qbit per limb (the reference's q_cons,
    const u64 *muv;       //   mu_cons, q_bit_cons symbols, ntt_60bit.cuh:8-10) or NULL -> scalars below
    const u32 *qbitv;
    u64 q, mu;
    u32 qbit;
    u32 num, division;    // poly p uses limb p % division  (ntt_60bit.cuh:391)
    u32 group_polys;      // polynomial p starts at a + (p / group_polys) * group_stride + (p % group_polys) * n;
    size_t group_stride;  //   group_polys = num for one contiguous [num][n] array (the reference's layout)
    u32 use_tma;          // bit 0: TMA tile movement; bits 1/2 (profiling only): skip the butterflies / skip the tile traffic
    u32 pf_dist;          // > 0: CTA b prefetches the tile of CTA b + pf_dist into L2 (about one wave of resident CTAs ahead)
    // forward strided pass only: when set, the input is not read from `a` but GENERATED -- coefficient j of every polynomial of
    // group g is ternary_value(gen_src[g * gen_stride + j], q_limb)  (encryption's u, bfv_encryption.cuh:23-36)
    const unsigned char *gen_src;
    size_t gen_stride;
    // p / group_polys and p / division as multiply-shift (ntt_args_finish): exact for p < 2^31; mul == 0 -> plain division.  The two
    // runtime divisions cost every CTA ~60 uniform-datapath instructions (I2F / MUFU.RCP sequences) before its first butterfly.
    u32 gp_mul, gp_sh, div_mul, div_sh;
};
__host__ __device__ __forceinline__ void fastdiv_make(u32 d, u32 &mul, u32 &sh)
{
    u32 lg = 0;
    while ((1ull << lg) < d) lg++;
    sh = 31 + lg;
    mul = (u32)((((u64)1 << sh) + d - 1) / d);          // ceil(2^sh / d) < 2^32 for d >= 1
    if (d <= 1) { mul = 0; sh = 0; }
}
// FD = false: plain runtime division.  Measured per kernel on one box (scripts/ab_ntt.py, profiles/r02_experiments.md): the multiply-shift
// form wins 5 % in the inverse strided pass and LOSES 1-6 % in the other three (ptxas schedules the uniform-datapath division under the
// tile wait; the shorter form perturbs its register allocation), so each kernel picks its own.
template <bool FD>
__host__ __device__ __forceinline__ u32 fastdiv(u32 p, u32 d, u32 mul, u32 sh)
{
    if (FD) return mul ? (u32)(((u64)p * mul) >> sh) : (d <= 1 ? p : p / d);
    return p / d;
}
__host__ __device__ __forceinline__ void ntt_args_finish(NttArgs &A)
{
    fastdiv_make(A.group_polys, A.gp_mul, A.gp_sh);
    fastdiv_make(A.division, A.div_mul, A.div_sh);
}
template <bool FD>
__device__ __forceinline__ u32 ntt_limb_of(const NttArgs &A, u32 p) { return p - fastdiv<FD>(p, A.division, A.div_mul, A.div_sh) * A.division; }

// ---- stage split per ring degree: K1 = S1+S2+S3 strided stages (rounds of 3 or 4), K2 contiguous stages -------
template <int LOGN> struct Sched;
#define NTTB200_SCHED(L, s1, s2, s3, k2, nt)                                             \
    template <> struct Sched<L> {                                                        \
        static constexpr int S1 = s1, S2 = s2, S3 = s3, K1 = s1 + s2 + s3, K2 = k2;      \
        static constexpr int NT = nt; /* 16-column tiles per CTA in the strided pass */  \
        static_assert(K1 + K2 == L, "bad split");                                        \
    };
NTTB200_SCHED(11, 4, 0, 0, 7, 8)
NTTB200_SCHED(12, 4, 0, 0, 8, 8)
NTTB200_SCHED(13, 3, 3, 0, 7, 2)
NTTB200_SCHED(14, 4, 3, 0, 7, 1)
#ifdef NTT_SCHED15_438
NTTB200_SCHED(15, 4, 3, 0, 8, 1)
#else
NTTB200_SCHED(15, 4, 4, 0, 7, 1)
#endif
NTTB200_SCHED(16, 4, 4, 0, 8, 1)
NTTB200_SCHED(17, 3, 3, 3, 8, 1)
#undef NTTB200_SCHED

constexpr int kContigRows = 128;  // rows (of 16 coefficients) per CTA in the contiguous pass
// Minimum resident CTAs per SM the register allocator must leave room for (occupancy vs. per-thread ILP trade-off,
// tuned on the B200: see DESIGN.md "occupancy sweep").
#ifndef NTT_MINB_S
#define NTT_MINB_S 3
#endif
#ifndef NTT_MINB_C
#define NTT_MINB_C 6
#endif

__device__ __forceinline__ void prefetch_l1(const void *p)
{
#if !defined(NTTB200_EMU) && !defined(NTT_PF_HINT)
    // a real load whose result is discarded: unlike the prefetch.global.L1 hint it is never dropped (contig pass -2 %)
    u64 sink;
    asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(sink) : "l"(p));
#elif !defined(NTTB200_EMU)
    asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}
// Twiddles of a round of S stages starting at table index twbase: stage s uses entries [(twbase << s), +2^s).
template <int S, class P>
__device__ __forceinline__ void prefetch_round(const P &pol, u32 twbase)
{
    NTT_UNROLL
    for (int s = 0; s < S; s++) pol.prefetch(twbase << s);      // 2^s * 8 bytes <= 64: one line per stage and table
}

// ---- arithmetic policies ------------------------------------------------------------------------------------
struct ShoupPolicy {
    static constexpr bool kLazyGS = false;
    u64 q, twoq, nq;
    const u64 *w, *ws;
    const LimbConst *l;
    struct Tw { u64 w, ws; };
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n)
    {
        l = A.lc + limb;
        q = l->q; twoq = l->twoq; nq = l->negq;
        w = A.tw + (size_t)limb * n;
#ifdef NTT_TWZ   /* experiment (profiles/r02_experiments.md): A.tws is an INTERLEAVED table {w, floor(w 2^64 / q)}, 2n words per limb */
        ws = A.tws + (size_t)limb * 2 * n;
#else
        ws = A.tws + (size_t)limb * n;
#endif
    }
#ifdef NTT_TWZ
    __device__ __forceinline__ Tw load(u32 i) const
    {
        const ulonglong2 e = __ldg(reinterpret_cast<const ulonglong2 *>(ws) + i);
        Tw t; t.w = e.x; t.ws = e.y; return t;
    }
    __device__ __forceinline__ void load2(u32 i, Tw &t0, Tw &t1) const
    {
#ifdef NTT_TWZ256
        asm("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];" : "=l"(t0.w), "=l"(t0.ws), "=l"(t1.w), "=l"(t1.ws) : "l"(ws + 2 * (size_t)i));
#else
        t0 = load(i); t1 = load(i + 1);
#endif
    }
    __device__ __forceinline__ void prefetch(u32 i) const { prefetch_l1(ws + 2 * (size_t)i); }
#else
#ifdef NTT_DBG_REGTW   /* profiling only (wrong results): twiddles from registers, no table loads */
    __device__ __forceinline__ Tw load(u32 i) const { Tw t; t.w = q - 12345u; t.ws = nq; (void)i; return t; }
    __device__ __forceinline__ void load2(u32 i, Tw &t0, Tw &t1) const { t0 = load(i); t1.w = twoq - q - 777u; t1.ws = nq + 99u; }
    __device__ __forceinline__ void load2_ldg(u32 i, Tw &t0, Tw &t1) const
#else
    __device__ __forceinline__ Tw load(u32 i) const { Tw t; t.w = __ldg(w + i); t.ws = __ldg(ws + i); return t; }
    __device__ __forceinline__ void load2(u32 i, Tw &t0, Tw &t1) const
#endif
    {
        ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(w + i));
        ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(ws + i));
        t0.w = a.x; t1.w = a.y; t0.ws = b.x; t1.ws = b.y;
    }
    // pull the line holding table entry i into L1 (no destination register: can be issued long before the use)
    __device__ __forceinline__ void prefetch(u32 i) const { prefetch_l1(w + i); prefetch_l1(ws + i); }
#endif
    // forward (Cooley-Tukey), X,Y in [0,4q) -> [0,4q)
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 x = csub(X, twoq);
        u64 T = shoup_mul_n(Y, t.w, t.ws, nq);
        X = x + T;
        Y = x - T + twoq;
    }
    __device__ __forceinline__ u64 fwd_final(u64 x) const { return csub(csub(x, twoq), q); }
    __device__ __forceinline__ void fwd_final_all(u64 (&v)[16]) const
    {
        NTT_UNROLL
        for (int i = 0; i < 16; i++) v[i] = fwd_final(v[i]);
    }
    // inverse (Gentleman-Sande), U,V in [0,2q) -> [0,2q)
    __device__ __forceinline__ void gs(u64 &U, u64 &V, const Tw &t) const
    {
        u64 s = U + V, d = U - V + twoq;
        U = csub(s, twoq);
        V = shoup_mul_n(d, t.w, t.ws, nq);
    }
    // lazy product with a key coefficient (companion ks): any 64-bit x -> [0, 2q), a valid inverse-round input
    __device__ __forceinline__ u64 mul_key(u64 x, u64 k, u64 ks) const { return shoup_mul_n(x, k, ks, nq); }
    // product of two CANONICAL residues with no companion (polynomial products): [0, 2q), a valid inverse-round input
    __device__ __forceinline__ u64 mul_generic(u64 x, u64 y) const { return csub(barrett_lazy(x, y, q, l->mu, (int)l->qbit), q); }
    // last inverse stage (length = 1) with n^-1 folded in; canonical outputs
    __device__ __forceinline__ void gs_last(u64 &U, u64 &V) const
    {
        u64 s = U + V, d = U - V + twoq;
        U = csub(shoup_mul_n(s, l->ninv, l->ninv_s, nq), q);
        V = csub(shoup_mul_n(d, l->w1ninv, l->w1ninv_s, nq), q);
    }
};

// Forward transform for q < 2^57: no conditional subtraction at all inside the transform, and an approximate Shoup
// quotient (3 wide multiplies instead of 4).  The approximate product is < 4q whatever its input, so X' = X + T and
// Y' = X - T + 4q stay non-negative and grow by at most 4q per stage: canonical inputs end below
// (4 log2(n) + 1) q <= 69 q < 2^64.  One reduction by floor(2^64/q) at the very end brings the outputs to [0, q).
// The inverse transform is ShoupPolicy's.
struct ShoupLazyPolicy : ShoupPolicy {
    u64 ratio, fourq;
    u32 pm_delta, pm_bits;      // q = 2^pm_bits - pm_delta with 69 * pm_delta < q (pm_bits = 0: not of that form)
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n)
    {
        ShoupPolicy::init(A, limb, n);
        ratio = l->ratio;
        fourq = twoq + twoq;
        const u32 b = l->qbit;
        const u64 d = (1ull << b) - q;                       // qbit <= 57 under this policy
        const bool ok = d < (1ull << 32) && d * 69 < q;
        pm_delta = (u32)d;
        pm_bits = ok ? b : 0u;
    }
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 T = shoup_mul_m(Y, t.w, t.ws, nq);
        u64 x = X;
        X = x + T;
        Y = x - T + fourq;
    }
    // x < 69 q  ->  [0, q).
    //  * q = 2^b - delta with a small delta (every prime of the reference's parameter sets: NTT primes are picked just below a
    //    power of two): x = hi * 2^b + lo gives x - hi * q = lo + hi * delta < q + 69 delta < 2q -- a shift, a mask, one 32x32
    //    multiplication and one conditional subtraction.  Measured: the general path below costs 14 % of the contiguous pass.
    //  * otherwise, for q > 2^32: floor(x / q) is at most 2 above hi32((x >> 32) * ratio) because ratio = floor(2^64 / q) < 2^32
    //    (dropped: x_lo * ratio / 2^64 < 1, x / 2^64 < 1, the floor).
    __device__ __forceinline__ u64 final_pm(u64 x) const
    {
        const u32 hi = (u32)(x >> pm_bits);                  // x < 69 q < 2^(b + 7)
        const u64 r = (x & ((1ull << pm_bits) - 1)) + (u64)hi * (u64)pm_delta;
        return csub(r, q);
    }
    __device__ __forceinline__ u64 final_q32(u64 x) const
    {
        const u64 qe = ((u64)(u32)(x >> 32) * (u64)(u32)ratio) >> 32;          // 32-bit quotient estimate
        u64 r = x - ((u64)(u32)qe * (u64)(u32)q + (((u64)(u32)qe * (q >> 32)) << 32));   // x - qe * q  in [0, 3q)
        return csub(csub(r, twoq), q);
    }
    __device__ __forceinline__ u64 final_any(u64 x) const
    {
        u64 r = x + mulhi64(x, ratio) * nq;     // x - floor(x * ratio / 2^64) * q  in [0, 2q)
        return csub(r, q);
    }
    __device__ __forceinline__ u64 fwd_final(u64 x) const
    {
        return pm_bits != 0 ? final_pm(x) : (ratio >> 32) == 0 ? final_q32(x) : final_any(x);
    }
    // the (CTA-uniform) choice is made ONCE for the 16 coefficients of a thread: each arm is then straight-line code with 16
    // independent reductions (with the choice inside the loop ptxas emitted 16 branchy, serial blocks and the cheap arm bought nothing)
    __device__ __forceinline__ void fwd_final_all(u64 (&v)[16]) const
    {
        if (pm_bits != 0) {
            NTT_UNROLL
            for (int i = 0; i < 16; i++) v[i] = final_pm(v[i]);
        } else if ((ratio >> 32) == 0) {
            NTT_UNROLL
            for (int i = 0; i < 16; i++) v[i] = final_q32(v[i]);
        } else {
            NTT_UNROLL
            for (int i = 0; i < 16; i++) v[i] = final_any(v[i]);
        }
    }
};

// Inverse transform for q < 2^57: approximate quotient and NO per-butterfly correction inside a register round.
// With B = 4q (the bound of every approximate Shoup product), a Gentleman-Sande butterfly on U < bB, V < bB gives
// U' = U + V < 2bB and V' = ((U - V + bB) * w) < B.  Starting a round of S <= 4 stages from values < B, the bound of
// register row p after k stages is B * 2^(k - bitlength(p)) -- a fixed pattern known at compile time -- so the bias of
// every subtraction is a constant shift of 4q and each row is brought back below B once, at the end of the round, by a
// ladder of (S - bitlength(p)) conditional subtractions: 15 per 32 butterflies instead of 32.  Largest intermediate:
// 16 B = 64 q < 2^63.  The last stage of the transform (n^-1 folded in) uses the exact quotient and canonicalises.
struct ShoupLazyInvPolicy : ShoupPolicy {
    static constexpr bool kLazyGS = true;
    u64 fourq;
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n)
    {
        ShoupPolicy::init(A, limb, n);
        fourq = twoq + twoq;
    }
    __device__ __forceinline__ u64 mul_key(u64 x, u64 k, u64 ks) const { return shoup_mul_m(x, k, ks, nq); }   // < B = 4q
    __device__ __forceinline__ u64 mul_generic(u64 x, u64 y) const { return barrett_lazy(x, y, q, l->mu, (int)l->qbit); }   // < 3q < B
    // e: log2 of the common bound multiplier of U and V (compile-time after unrolling)
    __device__ __forceinline__ void gs_lazy(u64 &U, u64 &V, const Tw &t, int e) const
    {
        const u64 s = U + V, d = U - V + (fourq << e);
        U = s;
        V = shoup_mul_m(d, t.w, t.ws, nq);
    }
    // x < 2^e * B  ->  x < B
    __device__ __forceinline__ u64 reduce_to_B(u64 x, int e) const
    {
        NTT_UNROLL
        for (int k = 4; k >= 1; k--)
            if (k <= e) x = csub(x, fourq << (k - 1));
        return x;
    }
    // last stage: U, V < 2^e * B; canonical outputs
    __device__ __forceinline__ void gs_last_lazy(u64 &U, u64 &V, int e) const
    {
        const u64 s = U + V, d = U - V + (fourq << e);
        U = csub(shoup_mul_n(s, l->ninv, l->ninv_s, nq), q);
        V = csub(shoup_mul_n(d, l->w1ninv, l->w1ninv_s, nq), q);
    }
};

// The reference's own arithmetic, operation for operation (ntt_60bit.cuh:424-440 forward, :494-513 inverse):
// canonical values, Barrett with the driver's (q, mu, qbit), one halving per inverse stage.  Needs nothing but
// the reference's tables and constants, so the header drop-in can call it statelessly.
struct BarrettPolicy {
    static constexpr bool kLazyGS = false;
    u64 q, mu, q2;
    int qbit;
    const u64 *w;
    struct Tw { u64 w; };
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n)
    {
        if (A.qv) { q = __ldg(A.qv + limb); mu = __ldg(A.muv + limb); qbit = (int)__ldg(A.qbitv + limb); }
        else { q = A.q; mu = A.mu; qbit = (int)A.qbit; }
        q2 = (q + 1) >> 1;
        w = A.tw + (size_t)limb * n;
    }
    __device__ __forceinline__ Tw load(u32 i) const { Tw t; t.w = __ldg(w + i); return t; }
    __device__ __forceinline__ void load2(u32 i, Tw &t0, Tw &t1) const
    {
        ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(w + i));
        t0.w = a.x; t1.w = a.y;
    }
    __device__ __forceinline__ void prefetch(u32 i) const { prefetch_l1(w + i); }
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 u = X;
        u64 v = barrett_ref(Y, t.w, q, mu, qbit);
        u64 s = u + v;
        s -= q * (u64)(s >= q);
        X = s;
        u += q * (u64)(u < v);
        Y = u - v;
    }
    __device__ __forceinline__ u64 fwd_final(u64 x) const { return x; }
    __device__ __forceinline__ void fwd_final_all(u64 (&)[16]) const {}
    __device__ __forceinline__ void gs(u64 &U, u64 &V, const Tw &t) const
    {
        u64 u = U, v = V;
        u64 s = u + v;
        s -= q * (u64)(s >= q);
        U = (s >> 1) + q2 * (s & 1);
        u += q * (u64)(u < v);
        u64 d = barrett_ref(u - v, t.w, q, mu, qbit);
        V = (d >> 1) + q2 * (d & 1);
    }
    __device__ __forceinline__ void gs_last(u64 &U, u64 &V) const { gs(U, V, load(1)); }
};

// ---- S consecutive stages on 16 registers laid out [2^S rows][NC = 16 >> S cols] -------------------------------
// Stage s of the round pairs rows i and i + 2^(S-1-s); block b = i >> (S-s) uses table entry (twbase << s) + b,
// where twbase = m0 + g0 (m0 = `length` of the round's first stage, g0 = this thread's group index there).
template <int S, int NC, int s, bool GS, bool LAST, class P>
__device__ __forceinline__ void one_stage(u64 (&v)[16], u32 twbase, const P &pol)
{
    constexpr int half = 1 << (S - 1 - s);
    const u32 base = twbase << s;
    // lazy inverse rounds: stages already done in this round = S-1-s, pair position k < half = 2^(S-1-s):
    // common bound exponent of the pair = (S-1-s) - bitlength(k)
    [[maybe_unused]] auto bexp = [](int k) { int bl = 0; while ((k >> bl) != 0) bl++; return (S - 1 - s) - bl; };
    if constexpr (s == 0) {
        if constexpr (GS && LAST) {
            NTT_UNROLL
            for (int k = 0; k < half; k++) {
                NTT_UNROLL
                for (int c = 0; c < NC; c++) {
                    if constexpr (P::kLazyGS) pol.gs_last_lazy(v[k * NC + c], v[(k + half) * NC + c], bexp(k));
                    else pol.gs_last(v[k * NC + c], v[(k + half) * NC + c]);
                }
            }
        } else {
            typename P::Tw t = pol.load(base);
            NTT_UNROLL
            for (int k = 0; k < half; k++) {
                NTT_UNROLL
                for (int c = 0; c < NC; c++) {
                    if constexpr (GS && P::kLazyGS) pol.gs_lazy(v[k * NC + c], v[(k + half) * NC + c], t, bexp(k));
                    else if constexpr (GS) pol.gs(v[k * NC + c], v[(k + half) * NC + c], t);
                    else pol.ct(v[k * NC + c], v[(k + half) * NC + c], t);
                }
            }
        }
    } else {
        NTT_UNROLL
        for (int b = 0; b < (1 << s); b += 2) {
            typename P::Tw t0, t1;
            pol.load2(base + b, t0, t1);
            NTT_UNROLL
            for (int k = 0; k < half; k++) {
                const int i = (b << (S - s)) + k, j = ((b + 1) << (S - s)) + k;
                NTT_UNROLL
                for (int c = 0; c < NC; c++) {
                    if constexpr (GS && P::kLazyGS) {
                        pol.gs_lazy(v[i * NC + c], v[(i + half) * NC + c], t0, bexp(k));
                        pol.gs_lazy(v[j * NC + c], v[(j + half) * NC + c], t1, bexp(k));
                    } else if constexpr (GS) {
                        pol.gs(v[i * NC + c], v[(i + half) * NC + c], t0);
                        pol.gs(v[j * NC + c], v[(j + half) * NC + c], t1);
                    } else {
                        pol.ct(v[i * NC + c], v[(i + half) * NC + c], t0);
                        pol.ct(v[j * NC + c], v[(j + half) * NC + c], t1);
                    }
                }
            }
        }
    }
}
template <int S, int NC, int s, class P>
__device__ __forceinline__ void ct_from(u64 (&v)[16], u32 twbase, const P &pol)
{
    if constexpr (s < S) {
        one_stage<S, NC, s, false, false>(v, twbase, pol);
        ct_from<S, NC, s + 1>(v, twbase, pol);
    }
}
template <int S, int NC, int s, bool LAST, class P>
__device__ __forceinline__ void gs_from(u64 (&v)[16], u32 twbase, const P &pol)
{
    if constexpr (s >= 0) {
        one_stage<S, NC, s, true, LAST>(v, twbase, pol);
        gs_from<S, NC, s - 1, LAST>(v, twbase, pol);
    }
}
template <int S, int NC, class P>
__device__ __forceinline__ void ct_stages(u64 (&v)[16], u32 twbase, const P &pol) { ct_from<S, NC, 0>(v, twbase, pol); }
template <int S, int NC, bool LAST, class P>
__device__ __forceinline__ void gs_stages(u64 (&v)[16], u32 twbase, const P &pol)
{
    gs_from<S, NC, S - 1, LAST>(v, twbase, pol);
    if constexpr (P::kLazyGS && !LAST) {
        // row p of the round ends below B * 2^(S - bitlength(p)): bring every row back below B
        NTT_UNROLL
        for (int p = 0; p < (1 << S); p++) {
            int bl = 0;
            while ((p >> bl) != 0) bl++;
            NTT_UNROLL
            for (int c = 0; c < NC; c++) v[p * NC + c] = pol.reduce_to_B(v[p * NC + c], S - bl);
        }
    }
}

// ---- register <-> tile moves ----------------------------------------------------------------------------------
// 2^S rows (rbase + (i << rsh)) x NC adjacent columns starting at col0
template <int S, bool SWZ, bool LOAD>
__device__ __forceinline__ void regs_rows(u64 *tile, u32 rbase, u32 rsh, u32 col0, u64 (&v)[16])
{
    constexpr int NC = 16 >> S;
    static_assert(NC == 1 || NC == 2, "rounds are radix-16 (1 col) or radix-8 (2 cols)");
    NTT_UNROLL
    for (int i = 0; i < (1 << S); i++) {
        const u32 row = rbase + ((u32)i << rsh);
        if (NC == 1) {
            u64 *p = tile + tile_off<SWZ>(row, col0);
            if (LOAD) v[i] = *p; else *p = v[i];
        } else {
            ulonglong2 *p = reinterpret_cast<ulonglong2 *>(tile + tile_off<SWZ>(row, col0));
            if (LOAD) { ulonglong2 t = *p; v[2 * i] = t.x; v[2 * i + 1] = t.y; }
            else *p = make_ulonglong2(v[2 * i], v[2 * i + 1]);
        }
    }
}
// one whole row (16 contiguous coefficients)
template <bool SWZ, bool LOAD>
__device__ __forceinline__ void regs_row(u64 *tile, u32 row, u64 (&v)[16])
{
    NTT_UNROLL
    for (int c = 0; c < 8; c++) {
        ulonglong2 *p = reinterpret_cast<ulonglong2 *>(tile + tile_off<SWZ>(row, 2 * c));
        if (LOAD) { ulonglong2 t = *p; v[2 * c] = t.x; v[2 * c + 1] = t.y; }
        else *p = make_ulonglong2(v[2 * c], v[2 * c + 1]);
    }
}

// One round of the strided pass: stages [J0, J0+S) of K1 on thread-in-tile index u in [0, 2^K1).
// EPI (final inverse round only): applied to the canonical outputs; coefficient index of (row, col) = row * C + cbase + col
template <class P, int K1, int J0, int S, bool INV, class EPI = NoEpi, bool SWZ = false>
__device__ __forceinline__ void strided_round(u64 *tile, u32 u, const P &pol, const EPI *epi = nullptr, u32 C = 0, u32 cbase = 0)
{
    constexpr int NC = 16 >> S;
    constexpr int LOB = K1 - J0 - S;  // row bits below the round's bits
    const u32 cg = u & ((1u << S) - 1u);
    const u32 rest = u >> S;
    const u32 lo = rest & ((1u << LOB) - 1u), hi = rest >> LOB;
    const u32 rbase = (hi << (K1 - J0)) + lo;
    u64 v[16];
    regs_rows<S, SWZ, true>(tile, rbase, LOB, cg * NC, v);
    if (!INV) ct_stages<S, NC>(v, (1u << J0) + hi, pol);
    else gs_stages<S, NC, J0 == 0>(v, (1u << J0) + hi, pol);
    if constexpr (INV && J0 == 0 && EPI::kMode != kEpiNone) {
        NTT_UNROLL
        for (int i = 0; i < (1 << S); i++) {
            const u32 j = (rbase + ((u32)i << LOB)) * C + cbase + cg * NC;
            NTT_UNROLL
            for (int c = 0; c < NC; c++) v[i * NC + c] = epi->apply(v[i * NC + c], j + c);
        }
    }
    regs_rows<S, SWZ, false>(tile, rbase, LOB, cg * NC, v);
}

// Dynamic shared memory is only guaranteed 16-byte aligned; the swizzled TMA image needs 1024.  The pad is added
// to the __shared__ array itself so the compiler keeps the shared address space (LDS/STS, not generic LD/ST).
__device__ __forceinline__ u64 *align_1024(unsigned char *p)
{
#ifdef NTTB200_EMU
    return reinterpret_cast<u64 *>((reinterpret_cast<uintptr_t>(p) + 1023) & ~(uintptr_t)1023);
#else
    const u32 pad = (1024u - ((u32)__cvta_generic_to_shared(p) & 1023u)) & 1023u;
    return reinterpret_cast<u64 *>(p + pad);
#endif
}

// Tiles a CTA processes back to back, all of the SAME polynomial (same limb constants, same table base), every tile's TMA load
// issued up front into its own buffer.  Measured on the B200 at n = 2^15 (profiles/r01_experiments.md): 2 tiles per CTA is 7 %
// SLOWER than 1 -- the doubled shared-memory carve-out (200 KB per SM) leaves ~28 KB of L1 for the twiddle lines, and their
// misses cost more than the hidden tile latency gains.  Kept as a build-time knob; the product uses 1.
#ifndef NTT_TPC
#define NTT_TPC 1
#endif
__host__ __device__ constexpr int tiles_per_cta(u32 tiles_per_poly) { return tiles_per_poly % NTT_TPC == 0 ? NTT_TPC : 1; }

// ---- pass "strided": grid (num * tiles / TPC), tiles = n / 2^K1 / 16 / NT column tiles per polynomial; 2^K1 * NT threads ----------------------
template <class P, int LOGN, bool INV, class EPI = NoEpi>
__global__ void __launch_bounds__((1 << Sched<LOGN>::K1) * Sched<LOGN>::NT, ((1 << Sched<LOGN>::K1) * Sched<LOGN>::NT) > 256 ? 2 : NTT_MINB_S)
ntt_strided_pass(const __grid_constant__ TensorMap tmap, NttArgs A, EpiArgs E)
{
    NTT_PDL_ENTER();
    using SC = Sched<LOGN>;
    constexpr int K1 = SC::K1, R = 1 << K1, NT = SC::NT, THREADS = R * NT;
    constexpr u32 n = 1u << LOGN, C = n >> K1;          // C columns per row
    constexpr int RB = R > 256 ? 256 : R;               // TMA box rows (box dims are capped at 256)
    constexpr u32 TILES = (C >> 4) / NT;
    constexpr int TPC = tiles_per_cta(TILES);
    constexpr u32 TG = TILES / TPC;                     // CTAs per polynomial
    constexpr size_t TILE_ELEMS = (size_t)NT * R * 16;
    NTT_DYN_SMEM(raw);
    u64 *tiles0 = align_1024(raw);
    u64 *bar = tiles0 + TPC * TILE_ELEMS;
    const u32 tid = threadIdx.x, p = blockIdx.x / TG;
    const u32 colbase = (blockIdx.x % TG) * (TPC * NT * 16);
    const u32 grp = fastdiv<INV>(p, A.group_polys, A.gp_mul, A.gp_sh), idx = p - grp * A.group_polys;   // polynomial idx of group grp
    const bool dbg_nocompute = (A.use_tma & 2u) != 0, dbg_nomem = (A.use_tma & 4u) != 0;
    const bool tma = (A.use_tma & 1u) != 0;
#ifndef NTTB200_EMU
    // The tile loads are issued FIRST, by the thread that also initialises their barriers (no CTA barrier in between): limb constants,
    // index arithmetic and twiddle prefetches of everybody else then run under the HBM latency.
    const bool gen = !INV && A.gen_src != nullptr;
    if (!dbg_nomem && tma && !gen && tid == 0) {
        for (int tt = 0; tt < TPC; tt++) mbar_init(bar + tt, 1);
        fence_mbar_init();
        for (int tt = 0; tt < TPC; tt++) {
            mbar_expect_tx(bar + tt, (u32)(NT * R * 128));
            for (int k = 0; k < NT; k++)
                for (int rc = 0; rc < R / RB; rc++)
                    tma_load_4d(tiles0 + tt * TILE_ELEMS + ((size_t)k * R + rc * RB) * 16, &tmap, bar + tt, (int)colbase + (tt * NT + k) * 16,
                                rc * RB, (int)idx, (int)grp);
        }
        const u32 fb = blockIdx.x + A.pf_dist;           // the CTA that will run here about one wave later
        if (A.pf_dist != 0 && fb < gridDim.x) {
            const u32 fp = fb / TG, fcol = (fb % TG) * (TPC * NT * 16);
            const u32 fgrp = fastdiv<INV>(fp, A.group_polys, A.gp_mul, A.gp_sh), fidx = fp - fgrp * A.group_polys;
            for (int k = 0; k < TPC * NT; k++)
                for (int rc = 0; rc < R / RB; rc++) tma_prefetch_4d(&tmap, (int)fcol + k * 16, rc * RB, (int)fidx, (int)fgrp);
        }
    }
#else
    const bool gen = !INV && A.gen_src != nullptr;
#endif
    P pol;
    pol.init(A, ntt_limb_of<INV>(A, p), n);
    u64 *g = A.a + (size_t)grp * A.group_stride + ((size_t)idx << LOGN) + colbase;
    {   // twiddle lines of every round of this thread, requested before the tile wait so they arrive under it
        const u32 uu = tid & (R - 1);
        if constexpr (SC::S2 != 0) prefetch_round<SC::S2>(pol, (1u << SC::S1) + ((uu >> SC::S2) >> (K1 - SC::S1 - SC::S2)));
        if constexpr (SC::S3 != 0) prefetch_round<SC::S3>(pol, (1u << (SC::S1 + SC::S2)) + (uu >> SC::S3));
        if (uu < 32) prefetch_round<SC::S1>(pol, 1u);
    }

    if (dbg_nomem) {
        __syncthreads();
    } else if (gen) {
        // one 16-byte shared-memory chunk (2 coefficients = 2 keystream bytes) per thread and step: eight neighbouring lanes cover one
        // 128-byte tile row, so the stores of a warp are 512 contiguous bytes (no bank conflicts; the row-per-thread version had
        // 8-way conflicts and ran the pass at 207 us instead of 148)
        const unsigned char *src = A.gen_src + (size_t)grp * A.gen_stride + colbase;
        constexpr u32 GEN_ITERS = (u32)(TPC * NT) * R * 8 / THREADS;        // = 8: every thread converts eight 2-byte chunks
        static_assert(GEN_ITERS * THREADS == (u32)(TPC * NT) * R * 8, "tile is a whole number of chunk sweeps");
        u32 b2[GEN_ITERS];
        NTT_UNROLL
        for (u32 it = 0; it < GEN_ITERS; it++) {                            // all keystream loads in flight before the first conversion
            const u32 e = tid + it * THREADS, seg = e >> 3, c = e & 7u, k = seg / R, row = seg % R;
            b2[it] = *reinterpret_cast<const unsigned short *>(src + (size_t)row * C + k * 16 + 2 * c);
        }
        NTT_UNROLL
        for (u32 it = 0; it < GEN_ITERS; it++) {
            const u32 e = tid + it * THREADS, seg = e >> 3, c = e & 7u;
            *reinterpret_cast<ulonglong2 *>(tiles0 + (size_t)seg * 16 + 2 * c) =
                make_ulonglong2(ternary_value((unsigned char)(b2[it] & 0xffu), pol.q), ternary_value((unsigned char)(b2[it] >> 8), pol.q));
        }
        __syncthreads();
    } else if (tma) {
#ifdef NTTB200_EMU
        if (tid == 0)
            for (int tt = 0; tt < TPC; tt++)
                for (int k = 0; k < NT; k++)
                    for (int rc = 0; rc < R / RB; rc++)
                        emu_tma_4d(true, &tmap, tiles0 + tt * TILE_ELEMS + ((size_t)k * R + rc * RB) * 16, (int)colbase + (tt * NT + k) * 16, rc * RB,
                                   (int)idx, (int)grp);
        __syncthreads();
#else
        __syncthreads();                                     // the barriers thread 0 initialised before issuing the loads
#endif
    } else {
        for (int k = 0; k < TPC * NT; k++) tile_copy_coop<false, true>(tiles0 + (size_t)k * R * 16, g + k * 16, C, R, tid, THREADS);
        __syncthreads();
    }

    const u32 u = tid & (R - 1);
#pragma unroll 1
    for (int tt = 0; tt < TPC; tt++) {
        u64 *tiles = tiles0 + tt * TILE_ELEMS;
        u64 *tile = tiles + (size_t)(tid >> K1) * R * 16;
#ifndef NTTB200_EMU
        if (!dbg_nomem && tma && !gen) mbar_wait(bar + tt, 0);
#endif
        if (dbg_nocompute) {
        } else if (!INV) {
            strided_round<P, K1, 0, SC::S1, false>(tile, u, pol);
#ifdef NTT_DBG_NOBAR   /* profiling only (wrong results): no CTA barrier between the rounds */
            if constexpr (SC::S2 != 0) { __syncwarp(); strided_round<P, K1, SC::S1, SC::S2, false>(tile, u, pol); }
#else
            if constexpr (SC::S2 != 0) { __syncthreads(); strided_round<P, K1, SC::S1, SC::S2, false>(tile, u, pol); }
#endif
            if constexpr (SC::S3 != 0) { __syncthreads(); strided_round<P, K1, SC::S1 + SC::S2, SC::S3, false>(tile, u, pol); }
        } else {
            if constexpr (SC::S3 != 0) { strided_round<P, K1, SC::S1 + SC::S2, SC::S3, true>(tile, u, pol); __syncthreads(); }
            if constexpr (SC::S2 != 0) { strided_round<P, K1, SC::S1, SC::S2, true>(tile, u, pol); __syncthreads(); }
            if constexpr (EPI::kMode != kEpiNone) {
                EPI epi;
                epi.init(E, grp, idx, n);
                strided_round<P, K1, 0, SC::S1, true, EPI>(tile, u, pol, &epi, C, colbase + ((u32)tt * NT + (tid >> K1)) * 16u);
            } else {
                strided_round<P, K1, 0, SC::S1, true>(tile, u, pol);
            }
        }

        if (dbg_nomem) {
        } else if (tma) {
#ifdef NTTB200_EMU
            __syncthreads();
            if (tid == 0)
                for (int k = 0; k < NT; k++)
                    for (int rc = 0; rc < R / RB; rc++)
                        emu_tma_4d(false, &tmap, tiles + ((size_t)k * R + rc * RB) * 16, (int)colbase + (tt * NT + k) * 16, rc * RB, (int)idx, (int)grp);
#else
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                for (int k = 0; k < NT; k++)
                    for (int rc = 0; rc < R / RB; rc++)
                        tma_store_4d(&tmap, tiles + ((size_t)k * R + rc * RB) * 16, (int)colbase + (tt * NT + k) * 16, rc * RB, (int)idx, (int)grp);
                tma_store_commit();
            }
#endif
        } else {
            __syncthreads();
            for (int k = 0; k < NT; k++) tile_copy_coop<false, false>(tiles + (size_t)k * R * 16, g + (tt * NT + k) * 16, C, R, tid, THREADS);
        }
    }
#ifndef NTTB200_EMU
    if (!dbg_nomem && tma && tid == 0) tma_store_wait_read<0>();
#endif
    (void)bar;
}

// ---- pass "contig": grid (num * n / 16 / 128 / TPC), 128 threads; CTA = TPC x 128 consecutive rows of one polynomial ---------------------------------
template <class P, int LOGN, bool INV>
__global__ void __launch_bounds__(kContigRows, Sched<LOGN>::K2 == 8 ? 4 : NTT_MINB_C)   // radix-16 first round needs > 80 registers
ntt_contig_pass(const __grid_constant__ TensorMap tmap, NttArgs A)
{
    NTT_PDL_ENTER();
    using SC = Sched<LOGN>;
    constexpr int K1 = SC::K1, K2 = SC::K2, SA = K2 - 4, NC = 16 >> SA, RT = kContigRows;
    constexpr u32 n = 1u << LOGN;
    constexpr u32 TILES = (n >> 4) / RT;
    constexpr int TPC = tiles_per_cta(TILES);
    constexpr u32 TG = TILES / TPC;
    NTT_DYN_SMEM(raw);
    u64 *tile0 = align_1024(raw);
    u64 *bar = tile0 + (size_t)TPC * RT * 16;
    const u32 tid = threadIdx.x;
    const u32 p = blockIdx.x / TG;
    const u32 ripbase = (blockIdx.x % TG) * (TPC * RT);      // first row (of 16 coefficients) inside the polynomial
    const u32 grp = fastdiv<false>(p, A.group_polys, A.gp_mul, A.gp_sh), idx = p - grp * A.group_polys;
    const int growbase = (int)(idx * (n >> 4) + ripbase);    // row inside the group
    const bool dbg_nocompute = (A.use_tma & 2u) != 0, dbg_nomem = (A.use_tma & 4u) != 0;
    const bool tma = (A.use_tma & 1u) != 0;
#ifndef NTTB200_EMU
    if (!dbg_nomem && tma && tid == 0) {                      // loads first (see ntt_strided_pass)
        for (int tt = 0; tt < TPC; tt++) mbar_init(bar + tt, 1);
        fence_mbar_init();
        for (int tt = 0; tt < TPC; tt++) {
            mbar_expect_tx(bar + tt, (u32)(RT * 128));
            tma_load_3d(tile0 + (size_t)tt * RT * 16, &tmap, bar + tt, 0, growbase + tt * RT, (int)grp);
        }
        const u32 fb = blockIdx.x + A.pf_dist;
        if (A.pf_dist != 0 && fb < gridDim.x) {
            const u32 fp = fb / TG, frip = (fb % TG) * (TPC * RT);
            const u32 fgrp = fastdiv<false>(fp, A.group_polys, A.gp_mul, A.gp_sh), fidx = fp - fgrp * A.group_polys;
            for (int tt = 0; tt < TPC; tt++) tma_prefetch_3d(&tmap, 0, (int)(fidx * (n >> 4) + frip) + tt * RT, (int)fgrp);
        }
    }
#endif
    P pol;
    pol.init(A, ntt_limb_of<false>(A, p), n);
    u64 *g = A.a + (size_t)grp * A.group_stride + ((size_t)idx << LOGN) + (size_t)ripbase * 16;
    prefetch_round<4>(pol, (n >> 4) + ripbase + tid);
    prefetch_round<SA>(pol, (1u << K1) + (ripbase >> SA) + (tid >> SA));

    if (dbg_nomem) {
        __syncthreads();
    } else if (tma) {
#ifdef NTTB200_EMU
        if (tid == 0)
            for (int tt = 0; tt < TPC; tt++) emu_tma_3d(true, &tmap, tile0 + (size_t)tt * RT * 16, 0, growbase + tt * RT, (int)grp);
        __syncthreads();
#else
        __syncthreads();                                     // the barriers thread 0 initialised before issuing the loads
#endif
    } else {
        tile_copy_coop<true, true>(tile0, g, 16, TPC * RT, tid, RT);
        __syncthreads();
    }

    const u32 t = tid & ((1u << SA) - 1u), bl = tid >> SA;     // lane-in-block, block-in-tile
#pragma unroll 1
    for (int tt = 0; tt < TPC; tt++) {
        u64 *tile = tile0 + (size_t)tt * RT * 16;
        const u32 rip0 = ripbase + tt * RT;
        const u32 twA = (1u << K1) + (rip0 >> SA) + bl;            // block index inside the polynomial
        const u32 twB = (n >> 4) + rip0 + tid;                     // row index inside the polynomial
        if (tt + 1 < TPC) {                                        // next tile's twiddle lines
            prefetch_round<4>(pol, twB + RT);
            prefetch_round<SA>(pol, twA + (RT >> SA));
        }
#ifndef NTTB200_EMU
        if (!dbg_nomem && tma) mbar_wait(bar + tt, 0);
#endif
        u64 v[16];
        if (dbg_nocompute) {
        } else if (!INV) {
            regs_rows<SA, true, true>(tile, bl << SA, 0, t * NC, v);
            ct_stages<SA, NC>(v, twA, pol);
            regs_rows<SA, true, false>(tile, bl << SA, 0, t * NC, v);
            __syncwarp();
            regs_row<true, true>(tile, tid, v);
            ct_stages<4, 1>(v, twB, pol);
#ifndef NTT_DBG_NOFINAL   /* profiling only (non-canonical results): what the final reduction costs */
            pol.fwd_final_all(v);
#endif
            regs_row<true, false>(tile, tid, v);
        } else {
            regs_row<true, true>(tile, tid, v);
            gs_stages<4, 1, false>(v, twB, pol);
            regs_row<true, false>(tile, tid, v);
            __syncwarp();
            regs_rows<SA, true, true>(tile, bl << SA, 0, t * NC, v);
            gs_stages<SA, NC, false>(v, twA, pol);
            regs_rows<SA, true, false>(tile, bl << SA, 0, t * NC, v);
        }

        if (dbg_nomem) {
        } else if (tma) {
#ifdef NTTB200_EMU
            __syncthreads();
            if (tid == 0) emu_tma_3d(false, &tmap, tile, 0, growbase + tt * RT, (int)grp);
#else
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) { tma_store_3d(&tmap, tile, 0, growbase + tt * RT, (int)grp); tma_store_commit(); }
#endif
        }
    }
    if (!dbg_nomem && !tma) {
        __syncthreads();
        tile_copy_coop<true, false>(tile0, g, 16, TPC * RT, tid, RT);
    }
#ifndef NTTB200_EMU
    if (!dbg_nomem && tma && tid == 0) tma_store_wait_read<0>();
#endif
    (void)bar;
}

// ---- whole transform in ONE kernel for n <= 4096 and FEW polynomials: the latency path -----------------------------------------------
// BASELINE config 1 (a single N = 4096 transform, ntt_60bit.cuh:314-386: forwardNTT = one <<<1,1024>>> launch, inverseNTT = two) is pure
// latency.  The two-kernel schedule gives one small polynomial to four 128-thread CTAs with 96 dependent butterflies per thread (20 us
// against the rebuilt reference's 16); a first one-kernel version with 16 coefficients per thread was no faster (profiles/
// r02_experiments.md).  Here one CTA of n/4 threads keeps the polynomial in shared memory and every thread owns FOUR coefficients per
// round = two stages in registers (24 butterflies per thread at n = 4096, five CTA barriers), which is what latency wants.
// Round k (stages 2k, 2k+1): groups {i, i + d, i + 2d, i + 3d}, d = n >> (2k + 2); table entries as in the reference
// (stage `length = m` reads psi[m + block]).  Used when the call has at most kSmallNttMaxPolys polynomials; batches keep the two-kernel path.
constexpr unsigned kSmallNttMaxPolys = 64;
template <class P, int LOGN, bool INV>
__global__ void __launch_bounds__(1 << (LOGN - 2)) ntt_single_pass(NttArgs A)
{
    NTT_PDL_ENTER();
    static_assert(!P::kLazyGS, "the latency kernel uses the per-butterfly-corrected policies");
    constexpr u32 n = 1u << LOGN, T = n >> 2;
    constexpr int PAIRS = LOGN / 2;                 // rounds of two stages; LOGN odd: one single-stage round more
    NTT_DYN_SMEM(raw);
    u64 *tile = reinterpret_cast<u64 *>(raw);
    const u32 t = threadIdx.x, p = blockIdx.x;
    const u32 grp = p / A.group_polys, idx = p - grp * A.group_polys;
    P pol;
    pol.init(A, p % A.division, n);
    u64 *g = A.a + (size_t)grp * A.group_stride + ((size_t)idx << LOGN);
    {   // 32 bytes per thread, coalesced
        const ulonglong2 x0 = reinterpret_cast<const ulonglong2 *>(g)[t], x1 = reinterpret_cast<const ulonglong2 *>(g)[t + T];
        // every round's table lines are requested now (real loads, results discarded: prefetch_l1) so that the loads inside the rounds hit
        // L1 (measured: no change at n = 4096 -- one SM's issue rate, not table latency, bounds this kernel: profiles/r02_experiments.md)
        NTT_UNROLL
        for (int k = 0; k < PAIRS; k++) {
            const u32 d = n >> (2 * k + 2), b0 = t / d, m0 = 1u << (2 * k);
            if (k > 0 || !INV) pol.prefetch(m0 + b0);
            pol.prefetch(2 * m0 + 2 * b0);
        }
        if constexpr (LOGN & 1) { pol.prefetch((n >> 1) + t); pol.prefetch((n >> 1) + t + T); }
        reinterpret_cast<ulonglong2 *>(tile)[t] = x0;
        reinterpret_cast<ulonglong2 *>(tile)[t + T] = x1;
    }
    __syncthreads();
    if (!INV) {
        NTT_UNROLL
        for (int k = 0; k < PAIRS; k++) {
            const u32 d = n >> (2 * k + 2), b0 = t / d, i = b0 * 4 * d + (t % d), m0 = 1u << (2 * k);
            u64 x0 = tile[i], x1 = tile[i + d], x2 = tile[i + 2 * d], x3 = tile[i + 3 * d];
            const typename P::Tw w0 = pol.load(m0 + b0);
            typename P::Tw w1a, w1b;
            pol.load2(2 * m0 + 2 * b0, w1a, w1b);
            pol.ct(x0, x2, w0); pol.ct(x1, x3, w0);
            pol.ct(x0, x1, w1a); pol.ct(x2, x3, w1b);
            if (2 * k + 2 == LOGN) { x0 = pol.fwd_final(x0); x1 = pol.fwd_final(x1); x2 = pol.fwd_final(x2); x3 = pol.fwd_final(x3); }
            tile[i] = x0; tile[i + d] = x1; tile[i + 2 * d] = x2; tile[i + 3 * d] = x3;
            __syncthreads();
        }
        if constexpr (LOGN & 1) {                   // last stage alone: length n/2, pairs (2j, 2j + 1), two butterflies per thread
            NTT_UNROLL
            for (int h = 0; h < 2; h++) {
                const u32 j = t + h * T;
                u64 x0 = tile[2 * j], x1 = tile[2 * j + 1];
                pol.ct(x0, x1, pol.load((n >> 1) + j));
                tile[2 * j] = pol.fwd_final(x0); tile[2 * j + 1] = pol.fwd_final(x1);
            }
            __syncthreads();
        }
    } else {
        if constexpr (LOGN & 1) {
            NTT_UNROLL
            for (int h = 0; h < 2; h++) {
                const u32 j = t + h * T;
                u64 x0 = tile[2 * j], x1 = tile[2 * j + 1];
                pol.gs(x0, x1, pol.load((n >> 1) + j));
                tile[2 * j] = x0; tile[2 * j + 1] = x1;
            }
            __syncthreads();
        }
        NTT_UNROLL
        for (int k = PAIRS - 1; k >= 0; k--) {
            const u32 d = n >> (2 * k + 2), b0 = t / d, i = b0 * 4 * d + (t % d), m0 = 1u << (2 * k);
            u64 x0 = tile[i], x1 = tile[i + d], x2 = tile[i + 2 * d], x3 = tile[i + 3 * d];
            typename P::Tw w1a, w1b;
            pol.load2(2 * m0 + 2 * b0, w1a, w1b);
            pol.gs(x0, x1, w1a); pol.gs(x2, x3, w1b);
            if (k == 0) { pol.gs_last(x0, x2); pol.gs_last(x1, x3); }
            else { const typename P::Tw w0 = pol.load(m0 + b0); pol.gs(x0, x2, w0); pol.gs(x1, x3, w0); }
            tile[i] = x0; tile[i + d] = x1; tile[i + 2 * d] = x2; tile[i + 3 * d] = x3;
            __syncthreads();
        }
    }
    reinterpret_cast<ulonglong2 *>(g)[t] = reinterpret_cast<const ulonglong2 *>(tile)[t];
    reinterpret_cast<ulonglong2 *>(g)[t + T] = reinterpret_cast<const ulonglong2 *>(tile)[t + T];
}

// ---- fused "contig forward pass  (.) key  ->  contig inverse pass" ----------------------------------------------------
// BFV never needs NTT(x) itself, only INTT(NTT(x) (.) key).  With the key's Shoup companions in HBM the
// product is a lazy Shoup multiplication, so the thread that finishes the forward row round multiplies its 16 coefficients
// in registers and walks straight into the inverse row round: no canonicalisation, no pointwise kernel, no HBM round trip
// of NTT(x), one shared-memory round trip less.  NOUT = 2 produces both halves of a ciphertext from one NTT(u) (pk0, pk1).
struct FusedArgs {
    NttArgs A;                 // data array + FORWARD tables (tw = psi, tws = psi_s); group = one item
    const u64 *twi, *twis;     // inverse tables
    const u64 *key, *key_s;    // key polynomials (NTT domain, canonical) and floor(key * 2^64 / q); [item?][NOUT][r][n]
    size_t key_item_stride;    // 0: one key for every item
    size_t key_half_stride;    // distance between the two key halves (NOUT = 2)
    u32 r;                     // polynomials per half = limbs; CTA index = ((item * r) + limb) * tiles + tile
    u32 in_off, out_off[2];    // polynomial index inside the item's group of the input / outputs for limb 0
    u32 items;
};

template <class PF, class PI, int LOGN, int NOUT>
__global__ void __launch_bounds__(kContigRows, 4)
ntt_contig_fused_mul(const __grid_constant__ TensorMap tmap, FusedArgs F)
{
    NTT_PDL_ENTER();
    using SC = Sched<LOGN>;
    constexpr int K1 = SC::K1, K2 = SC::K2, SA = K2 - 4, NC = 16 >> SA, RT = kContigRows;
    constexpr u32 n = 1u << LOGN, TILES = (n >> 4) / RT;
    const NttArgs &A = F.A;
    NTT_DYN_SMEM(raw);
    u64 *tile = align_1024(raw);
    u64 *tile2 = tile + (size_t)RT * 16;                 // second output (NOUT = 2)
    u64 *bar = tile + (size_t)RT * 16 * NOUT;
    const u32 tid = threadIdx.x;
    const u32 pl = blockIdx.x / TILES, rip0 = (blockIdx.x % TILES) * RT;
    const u32 item = pl / F.r, limb = pl - item * F.r;
    const int row_in = (int)((F.in_off + limb) * (n >> 4) + rip0);
    const int row_o0 = (int)((F.out_off[0] + limb) * (n >> 4) + rip0);
    const int row_o1 = (int)((F.out_off[1] + limb) * (n >> 4) + rip0);
#ifndef NTTB200_EMU
    if ((A.use_tma & 1u) && tid == 0) {                      // load first: constants and twiddle prefetches run under its latency
        mbar_init(bar, 1); fence_mbar_init();
        mbar_expect_tx(bar, (u32)(RT * 128)); tma_load_3d(tile, &tmap, bar, 0, row_in, (int)item);
    }
#endif
    PF pf;
    pf.init(A, limb, n);
    NttArgs Ai = A;
    Ai.tw = F.twi; Ai.tws = F.twis;
    PI pi;
    pi.init(Ai, limb, n);
    u64 *gbase = A.a + (size_t)item * A.group_stride;

    if (A.use_tma & 1u) {
#ifdef NTTB200_EMU
        if (tid == 0) emu_tma_3d(true, &tmap, tile, 0, row_in, (int)item);
        __syncthreads();
#else
        __syncthreads();
        mbar_wait(bar, 0);
#endif
    } else {
        tile_copy_coop<true, true>(tile, gbase + (size_t)row_in * 16, 16, RT, tid, RT);
        __syncthreads();
    }

    const u32 t = tid & ((1u << SA) - 1u), bl = tid >> SA;
    const u32 twA = (1u << K1) + (rip0 >> SA) + bl, twB = (n >> 4) + rip0 + tid;
    u64 v[16];
    // forward: column round, row round (values stay lazy)
    regs_rows<SA, true, true>(tile, bl << SA, 0, t * NC, v);
    ct_stages<SA, NC>(v, twA, pf);
    regs_rows<SA, true, false>(tile, bl << SA, 0, t * NC, v);
    __syncwarp();
    regs_row<true, true>(tile, tid, v);
    ct_stages<4, 1>(v, twB, pf);
    // (.) key, inverse row round, per output.  Two outputs: the lazy forward row is PARKED in the thread's own tile row and re-read
    // per output (the __syncwarp keeps the compiler from forwarding the stored registers), so no 16-coefficient vector stays live
    // across an output: 128 registers / four CTAs per SM instead of 168 / three.  Output 0 goes to tile2, output 1 overwrites the
    // parked row in place (only this lane touches its row before the next __syncwarp).
    if (NOUT == 2) { regs_row<true, false>(tile, tid, v); __syncwarp(); }
    u64 *const out_tile0 = NOUT == 2 ? tile2 : tile, *const out_tile1 = tile;
    const size_t krow = (size_t)item * F.key_item_stride + ((size_t)limb << LOGN) + ((size_t)(rip0 + tid) << 4);
    NTT_UNROLL
    for (int o = 0; o < NOUT; o++) {
        const u64 *kp = F.key + krow + (size_t)o * F.key_half_stride, *ks = F.key_s + krow + (size_t)o * F.key_half_stride;
        if (NOUT == 2) regs_row<true, true>(tile, tid, v);
        u64 x[16];
        NTT_UNROLL
        for (int c = 0; c < 8; c++) {
            const ulonglong2 kv = __ldg(reinterpret_cast<const ulonglong2 *>(kp) + c), sv = __ldg(reinterpret_cast<const ulonglong2 *>(ks) + c);
            x[2 * c] = pi.mul_key(v[2 * c], kv.x, sv.x);
            x[2 * c + 1] = pi.mul_key(v[2 * c + 1], kv.y, sv.y);
        }
        gs_stages<4, 1, false>(x, twB, pi);
        regs_row<true, false>(o == 0 ? out_tile0 : out_tile1, tid, x);
        if (NOUT == 2 && o == 0) __syncwarp();
    }
    __syncwarp();
    // inverse column round on each output tile
    NTT_UNROLL
    for (int o = 0; o < NOUT; o++) {
        u64 *dst = o == 0 ? out_tile0 : out_tile1;
        regs_rows<SA, true, true>(dst, bl << SA, 0, t * NC, v);
        gs_stages<SA, NC, false>(v, twA, pi);
        regs_rows<SA, true, false>(dst, bl << SA, 0, t * NC, v);
    }

    if (A.use_tma & 1u) {
#ifdef NTTB200_EMU
        __syncthreads();
        if (tid == 0) {
            emu_tma_3d(false, &tmap, out_tile0, 0, row_o0, (int)item);
            if (NOUT == 2) emu_tma_3d(false, &tmap, out_tile1, 0, row_o1, (int)item);
        }
#else
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tma_store_3d(&tmap, out_tile0, 0, row_o0, (int)item);
            if (NOUT == 2) tma_store_3d(&tmap, out_tile1, 0, row_o1, (int)item);
            tma_store_commit();
            tma_store_wait_read<0>();
        }
#endif
    } else {
        __syncthreads();
        tile_copy_coop<true, false>(out_tile0, gbase + (size_t)row_o0 * 16, 16, RT, tid, RT);
        if (NOUT == 2) tile_copy_coop<true, false>(out_tile1, gbase + (size_t)row_o1 * 16, 16, RT, tid, RT);
    }
    (void)bar;
}

// ---- fused polynomial product: "contig forward pass of a and b  (.)  ->  contig inverse pass" ---------------------------------
// c = a * b in Z_q[X]/(X^n + 1) needs NTT(a), NTT(b) only inside the product (full_poly_mul_device / half_poly_mul_device,
// poly_arithmetic.cuh:296-310: forwardNTTdouble, barrett, inverseNTT = 7 kernels there).  After the strided forward passes of a
// and b this kernel takes the same tile of both, finishes the two forward transforms, multiplies coefficient-wise (Barrett on
// canonical values: neither operand has a Shoup companion) and runs the contiguous inverse pass; the strided inverse pass
// follows.  A_FWD / B_FWD = false: that operand is already in the NTT domain (bit-reversed order, canonical) and is only read --
// keygen's  INTT(NTT(s) (.) a)  with both operands transformed is the <false, false> instantiation.
struct PolymulArgs {
    NttArgs A;                 // operand a; forward tables
    u64 *out;                  // result array, a's group description (== A.a: in place)
    const u64 *b;              // operand b, [num][n], same polynomial order as a; clobbered when B_FWD
    const u64 *twi, *twis;     // inverse tables
    u32 b_group_polys;         // b's own group description (polynomial p of b at (p / gp) * stride + (p % gp) * n)
    size_t b_group_stride;
};

template <class PF, class PI, int LOGN, bool A_FWD, bool B_FWD>
__global__ void __launch_bounds__(kContigRows, 4)
ntt_contig_polymul(const __grid_constant__ TensorMap tmap_a, const __grid_constant__ TensorMap tmap_b,
                   const __grid_constant__ TensorMap tmap_out, PolymulArgs F)
{
    using SC = Sched<LOGN>;
    constexpr int K1 = SC::K1, K2 = SC::K2, SA = K2 - 4, NC = 16 >> SA, RT = kContigRows;
    constexpr u32 n = 1u << LOGN, TILES = (n >> 4) / RT;
    const NttArgs &A = F.A;
    NTT_DYN_SMEM(raw);
    u64 *tile = align_1024(raw);
    u64 *tile2 = tile + (size_t)RT * 16;
    u64 *bar = tile + (size_t)RT * 16 * 2;
    const u32 tid = threadIdx.x;
    const u32 p = blockIdx.x / TILES, rip0 = (blockIdx.x % TILES) * RT;
    const u32 grp = p / A.group_polys, idx = p - grp * A.group_polys;
    const u32 bgrp = p / F.b_group_polys, bidx = p - bgrp * F.b_group_polys;
    const u32 limb = p % A.division;
    const int row_a = (int)(idx * (n >> 4) + rip0), row_b = (int)(bidx * (n >> 4) + rip0);
#ifndef NTTB200_EMU
    if ((A.use_tma & 1u) && tid == 0) {                      // loads first
        mbar_init(bar, 1); fence_mbar_init();
        mbar_expect_tx(bar, (u32)(2 * RT * 128));
        tma_load_3d(tile, &tmap_a, bar, 0, row_a, (int)grp);
        tma_load_3d(tile2, &tmap_b, bar, 0, row_b, (int)bgrp);
    }
#endif
    PF pf;
    pf.init(A, limb, n);
    NttArgs Ai = A;
    Ai.tw = F.twi; Ai.tws = F.twis;
    PI pi;
    pi.init(Ai, limb, n);
    u64 *ga = A.a + (size_t)grp * A.group_stride + ((size_t)idx << LOGN) + (size_t)rip0 * 16;
    u64 *go = F.out + (size_t)grp * A.group_stride + ((size_t)idx << LOGN) + (size_t)rip0 * 16;
    const u64 *gb = F.b + (size_t)bgrp * F.b_group_stride + ((size_t)bidx << LOGN) + (size_t)rip0 * 16;

    if (A.use_tma & 1u) {
#ifdef NTTB200_EMU
        if (tid == 0) { emu_tma_3d(true, &tmap_a, tile, 0, row_a, (int)grp); emu_tma_3d(true, &tmap_b, tile2, 0, row_b, (int)bgrp); }
        __syncthreads();
#else
        __syncthreads();
        mbar_wait(bar, 0);
#endif
    } else {
        tile_copy_coop<true, true>(tile, ga, 16, RT, tid, RT);
        tile_copy_coop<true, true>(tile2, const_cast<u64 *>(gb), 16, RT, tid, RT);
        __syncthreads();
    }

    const u32 t = tid & ((1u << SA) - 1u), bl = tid >> SA;
    const u32 twA = (1u << K1) + (rip0 >> SA) + bl, twB = (n >> 4) + rip0 + tid;
    u64 v[16];
    // forward column rounds
    if constexpr (A_FWD) {
        regs_rows<SA, true, true>(tile, bl << SA, 0, t * NC, v);
        ct_stages<SA, NC>(v, twA, pf);
        regs_rows<SA, true, false>(tile, bl << SA, 0, t * NC, v);
    }
    if constexpr (B_FWD) {
        regs_rows<SA, true, true>(tile2, bl << SA, 0, t * NC, v);
        ct_stages<SA, NC>(v, twA, pf);
        regs_rows<SA, true, false>(tile2, bl << SA, 0, t * NC, v);
    }
    if constexpr (A_FWD || B_FWD) __syncwarp();
    // b: forward row round, canonical, parked in its own row of tile2 (only this lane touches the row from here on)
    if constexpr (B_FWD) {
        regs_row<true, true>(tile2, tid, v);
        ct_stages<4, 1>(v, twB, pf);
        pf.fwd_final_all(v);
        regs_row<true, false>(tile2, tid, v);
    }
    // a: forward row round, canonical, (.) b, inverse row round
    regs_row<true, true>(tile, tid, v);
    if constexpr (A_FWD) {
        ct_stages<4, 1>(v, twB, pf);
        pf.fwd_final_all(v);
    }
    NTT_UNROLL
    for (int c = 0; c < 8; c++) {
        const ulonglong2 bv = *reinterpret_cast<const ulonglong2 *>(tile2 + tile_off<true>(tid, 2 * c));
        v[2 * c] = pi.mul_generic(v[2 * c], bv.x);
        v[2 * c + 1] = pi.mul_generic(v[2 * c + 1], bv.y);
    }
    gs_stages<4, 1, false>(v, twB, pi);
    regs_row<true, false>(tile, tid, v);
    __syncwarp();
    // inverse column round
    regs_rows<SA, true, true>(tile, bl << SA, 0, t * NC, v);
    gs_stages<SA, NC, false>(v, twA, pi);
    regs_rows<SA, true, false>(tile, bl << SA, 0, t * NC, v);

    if (A.use_tma & 1u) {
#ifdef NTTB200_EMU
        __syncthreads();
        if (tid == 0) emu_tma_3d(false, &tmap_out, tile, 0, row_a, (int)grp);
#else
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) { tma_store_3d(&tmap_out, tile, 0, row_a, (int)grp); tma_store_commit(); tma_store_wait_read<0>(); }
#endif
    } else {
        __syncthreads();
        tile_copy_coop<true, false>(tile, go, 16, RT, tid, RT);
    }
    (void)bar;
}

}  // namespace nttb200
