// bfly_variants.cuh -- butterfly formulations that were measured and NOT adopted (tools/bfly_ubench.cu keeps them
// reproducible; profiles/r01_bfly_ubench*.json holds the numbers).  Not included by the library.
//   ShoupLazy2Policy  all-chain quotient (ptxas re-splits the zero-extended accumulators: no gain)
//   ShoupLazyHPolicy  half-scale quotient on a 63-bit companion (one IMAD.HI replaces a WIDE: no gain)
//   ShoupLazyFPolicy  FP64-assisted cross terms (DFMA is not free next to the integer stream, I2F.F64 costs ~7 issue cycles: slower)
#pragma once
#include "../csrc/ntt_kernels.cuh"
namespace nttb200 {
// Approximate high product as one accumulate chain: yh*sl, then yl*sh + hi32(previous), then yh*sh + hi32(previous).
// Result in {exact-2 .. exact} (it keeps the carry between the two cross products that mulhi64_approx drops, never more).
// 3 IMAD.WIDE.U32 + 2 register moves for the zero-extended hi words.
__host__ __device__ __forceinline__ u64 mulhi64_approx_c(u64 y, u64 s)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 yl, yh, sl, sh, t, dm;\n\t"
        ".reg .u64 acc;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {sl, sh}, %2;\n\t"
        "mul.wide.u32 acc, yh, sl;\n\t"
        "mov.b64 {dm, t}, acc;\n\t"
        "cvt.u64.u32 acc, t;\n\t"
        "mad.wide.u32 acc, yl, sh, acc;\n\t"
        "mov.b64 {dm, t}, acc;\n\t"
        "cvt.u64.u32 acc, t;\n\t"
        "mad.wide.u32 acc, yh, sh, acc;\n\t"
        "mov.b64 %0, acc;\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(s));
    return r;
#else
    const u64 yl = (u32)y, yh = y >> 32, sl = (u32)s, sh = s >> 32;
    return yh * sh + ((yl * sh + ((yh * sl) >> 32)) >> 32);
#endif
}

// Shoup multiplication, chain formulation: result in [0, 4q) for any 64-bit y.
__host__ __device__ __forceinline__ u64 shoup_mul_c(u64 y, u64 w, u64 ws, u64 negq)
{
    return mullo_sum2_x(y, w, mulhi64_approx_c(y, ws), negq);
}

// Half-scale quotient for Y < 2^63 and a 63-bit companion s = floor(w * 2^63 / q): yh*sl + yl*sh cannot overflow 64 bits,
// so the two cross products share ONE mad.wide chain and only its high word is added to yh*sh.  Result in
// {floor(y*s/2^64) - 1, floor(y*s/2^64)}, i.e. y*w/(2q) - 2.5 < result <= y*w/(2q).
__host__ __device__ __forceinline__ u64 mulhi64_approx_h(u64 y, u64 s)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 yl, yh, sl, sh, t, dm, rl, rh;\n\t"
        ".reg .u64 acc, p1;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {sl, sh}, %2;\n\t"
        "mul.wide.u32 acc, yh, sl;\n\t"
        "mad.wide.u32 acc, yl, sh, acc;\n\t"
        "mul.wide.u32 p1, yh, sh;\n\t"
        "mov.b64 {dm, t}, acc;\n\t"
        "mov.b64 {rl, rh}, p1;\n\t"
        "add.cc.u32 rl, rl, t;\n\t"
        "addc.u32 rh, rh, 0;\n\t"
        "mov.b64 %0, {rl, rh};\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(s));
    return r;
#else
    const u64 yl = (u32)y, yh = y >> 32, sl = (u32)s, sh = s >> 32;
    return yh * sh + ((yl * sh + yh * sl) >> 32);
#endif
}
// y < 2^63, ws63 = floor(w * 2^63 / q), neg2q = 2^64 - 2q: result = y*w - 2q*quotient in [0, 5q).
__host__ __device__ __forceinline__ u64 shoup_mul_h(u64 y, u64 w, u64 ws63, u64 neg2q)
{
    return mullo_sum2_x(y, w, mulhi64_approx_h(y, ws63), neg2q);
}

// ---- FP64-assisted quotient -------------------------------------------------------------------------------------------
// On the B200 an IMAD.WIDE occupies the dispatch port for 4 cycles (IMAD: 2, ALU: 1) while DFMA issues on its own pipe
// (tools/ipipe2_ubench.cu).  Of the three 32x32 products of the approximate high product only yh*sh needs all 64 bits;
// the two cross products are needed divided by 2^32, i.e. to a relative precision of 2^-33 -- a double-precision FMA
// rounded DOWN delivers that:  cross = yh*sl + yl*sh < 1.25 * 2^64, computed error in (-2^13, 0], so
// floor(cross_fp / 2^33) >= floor(cross / 2^33) - 1 and the quotient returned here is in {exact-3 .. exact}.
// u32 -> double conversions are exact (2^52 + x built from the bits, minus 2^52).
struct ShoupF { u64 w; u32 sh; double dsl, dsh; };     // twiddle, high word of its companion, both companion words as doubles
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ double u32_to_double(u32 x) { return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0; }
#else
static inline double u32_to_double(u32 x) { return (double)x; }
#endif
__host__ __device__ __forceinline__ ShoupF make_shoupf(u64 w, u64 ws)
{
    ShoupF t;
    t.w = w; t.sh = (u32)(ws >> 32);
    t.dsl = u32_to_double((u32)ws); t.dsh = u32_to_double(t.sh);
    return t;
}
// Returns quotient + kFpBias where quotient is in {exact-2 .. exact}: the raw bits of the double 2^52 + floor(cross / 2^32)
// are used directly as the accumulator of the yh*sh multiply-add, so no double -> integer conversion is ever issued; the
// bias (the exponent field, 0x4330 << 48) is taken out again by constants folded into the butterfly's additions.
#define NTTB200_FP_BIAS 0x4330000000000000ull
#ifdef NTT_FP_I2F
#define NTTB200_U2D(x) __uint2double_rn(x)        /* I2F.F64.U32: one instruction, exact */
#else
#define NTTB200_U2D(x) u32_to_double(x)
#endif
__host__ __device__ __forceinline__ u64 mulhi64_approx_f(u64 y, const ShoupF &t)
{
    const u32 yl = (u32)y, yh = (u32)(y >> 32);
#if defined(__CUDA_ARCH__)
    const double c = __fma_rd(NTTB200_U2D(yh), t.dsl, __dmul_rd(NTTB200_U2D(yl), t.dsh));
    const double m = __fma_rd(c, 2.3283064365386963e-10 /* 2^-32 */, 4503599627370496.0);
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(yh), "r"(t.sh), "l"((u64)__double_as_longlong(m)));
    return r;
#else
    // the emulator computes the value the directed-rounding device sequence yields: cross rounded down twice
    const unsigned __int128 a = (unsigned __int128)yl * t.sh, b = (unsigned __int128)yh * (u32)(u64)t.dsl;
    auto rd53 = [](unsigned __int128 v) { int bl = 0; while (bl < 128 && (v >> bl) != 0) bl++; if (bl <= 53) return v; return (v >> (bl - 53)) << (bl - 53); };
    const unsigned __int128 c = rd53(rd53(a) + b);
    return (u64)yh * t.sh + (u64)(c >> 32) + NTTB200_FP_BIAS;
#endif
}
// y*w - quotient*q + kFpBias * negq  (mod 2^64): [0, 4q) plus the constant the caller removes
__host__ __device__ __forceinline__ u64 shoup_mul_f(u64 y, const ShoupF &t, u64 negq)
{
    return mullo_sum2_x(y, t.w, mulhi64_approx_f(y, t), negq);
}


// Same two policies on the chain formulation of the product (shoup_mul_c: 9 IMAD-class + 3 ALU instead of 9 + 5).
struct ShoupLazy2Policy : ShoupLazyPolicy {
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 T = shoup_mul_c(Y, t.w, t.ws, nq);
        u64 x = X;
        X = x + T;
        Y = x - T + fourq;
    }
};
struct ShoupLazyInv2Policy : ShoupLazyInvPolicy {
    __device__ __forceinline__ u64 mul_key(u64 x, u64 k, u64 ks) const { return shoup_mul_c(x, k, ks, nq); }
    __device__ __forceinline__ void gs_lazy(u64 &U, u64 &V, const Tw &t, int e) const
    {
        const u64 s = U + V, d = U - V + (fourq << e);
        U = s;
        V = shoup_mul_c(d, t.w, t.ws, nq);
    }
};

struct ShoupLazy3Policy_unused : ShoupLazyPolicy {     // split-carry quotient + chain low product
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 T = shoup_mul_m(Y, t.w, t.ws, nq);
        u64 x = X;
        X = x + T;
        Y = x - T + fourq;
    }
};
// Half-scale quotient (63-bit companions in tws, values < 2^63): products < 5q, bias 5q per stage.
struct ShoupLazyHPolicy : ShoupLazyPolicy {
    u64 n2q, fiveq;
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n)
    {
        ShoupLazyPolicy::init(A, limb, n);
        n2q = nq + nq;
        fiveq = fourq + q;
    }
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 T = shoup_mul_h(Y, t.w, t.ws, n2q);
        u64 x = X;
        X = x + T;
        Y = x - T + fiveq;
    }
};

// FP64-assisted quotient (modarith.cuh: shoup_mul_f): products < 4q (plus a bias constant removed by the additions).
struct ShoupLazyFPolicy : ShoupLazyPolicy {
    typedef ShoupF Tw;
    u64 cb, fourq_cb;     // -(kFpBias * negq) (mod 2^64) and 4q minus that
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n)
    {
        ShoupLazyPolicy::init(A, limb, n);
        cb = (u64)l->pad << 32;          // loaded, not derived: ptxas would re-derive a computed constant at every use
        fourq_cb = fourq - cb;
    }
    __device__ __forceinline__ Tw load(u32 i) const { return make_shoupf(__ldg(w + i), __ldg(ws + i)); }
    __device__ __forceinline__ void load2(u32 i, Tw &t0, Tw &t1) const
    {
        ulonglong2 a = __ldg(reinterpret_cast<const ulonglong2 *>(w + i));
        ulonglong2 b = __ldg(reinterpret_cast<const ulonglong2 *>(ws + i));
        t0 = make_shoupf(a.x, b.x); t1 = make_shoupf(a.y, b.y);
    }
    __device__ __forceinline__ void ct(u64 &X, u64 &Y, const Tw &t) const
    {
        u64 T = shoup_mul_f(Y, t, nq);      // true product - cb
        u64 x = X;
        X = x + T + cb;
        Y = x - T + fourq_cb;
    }
};


}  // namespace nttb200
