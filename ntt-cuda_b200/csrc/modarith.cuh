// modarith.cuh -- 64-bit modular arithmetic for the B200 integer pipe.
//
// Replaces the reference's device arithmetic layer: uint128.h:343-373 (mul64 = 9 x 32-bit mul/mad with
// carries, sub128) and ntt_60bit.cuh:44-61 (singleBarrett on the emulated 128-bit type).  Here a
// 64x64 product is `mul.lo.u64` + `mul.hi.u64` (IMAD.WIDE.U32 chains), and the NTT butterflies use
// Shoup/Harvey lazy multiplication by a precomputed twiddle companion (1 mul.hi + 2 mul.lo).
//
// Two arithmetic flavours are provided:
//   * shoup_mul / lazy [0,4q) butterflies   -- the fast path (context owns companion tables)
//   * barrett_ref                           -- bit-for-bit the reference's Barrett sequence, for the
//                                              pointwise kernels and the stateless drop-in NTT path
#pragma once
#include "compat.cuh"

namespace nttb200 {

// Per-limb constants kept in device global memory by a context (see nttb200.cu: build_limb_consts).
struct LimbConst {
    u64 q;          // modulus (< 2^62)
    u64 twoq;       // 2q
    u64 mu;         // floor(2^(2*qbit) / q)            (demo.cu:157-165)
    u64 ninv;       // n^-1 mod q
    u64 ninv_s;     // floor(ninv * 2^64 / q)
    u64 w1ninv;     // psiinv[1] * n^-1 mod q           (last inverse stage with the scaling folded in)
    u64 w1ninv_s;   // its Shoup companion
    u64 ratio;      // floor(2^64 / q)                  (final reduction of the lazy forward transform)
    u64 negq;       // 2^64 - q
    u32 qbit;       // floor(log2 q) + 1                (demo.cu:69)
    u32 pad;
};

// high 64 bits of the 128-bit product (mul.hi.u64 on the device)
__host__ __device__ __forceinline__ u64 mulhi64(u64 a, u64 b)
{
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

// x in [0, 2m) -> [0, m)
__host__ __device__ __forceinline__ u64 csub(u64 x, u64 m) { return x >= m ? x - m : x; }

// Shoup multiplication: y arbitrary 64-bit, w < q, ws = floor(w * 2^64 / q).  Result in [0, 2q).
__host__ __device__ __forceinline__ u64 shoup_mul(u64 y, u64 w, u64 ws, u64 q)
{
    u64 qhat = mulhi64(y, ws);
    return y * w - qhat * q;
}

// low 64 bits of a*b + c*d as one accumulation chain: 2 IMAD.WIDE.U32 + 4 IMAD, no carry / negate fix-ups (nvcc's own
// code for `a*b - c*q` is 10 instructions).  The a*b part (1 wide + 2 lo) is issued first and does not depend on c, so
// in a Shoup product it overlaps the mul.hi that produces c; only 1 wide + 2 lo remain on the critical path after it.
__host__ __device__ __forceinline__ u64 mullo_sum2(u64 a, u64 b, u64 c, u64 d)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, cl, ch, dl, dh, rl, rh;\n\t"
        ".reg .u64 acc;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "mov.b64 {cl, ch}, %3;\n\t"
        "mov.b64 {dl, dh}, %4;\n\t"
        "mul.wide.u32 acc, al, bl;\n\t"
        "mov.b64 {rl, rh}, acc;\n\t"
        "mad.lo.u32 rh, al, bh, rh;\n\t"
        "mad.lo.u32 rh, ah, bl, rh;\n\t"
        "mov.b64 acc, {rl, rh};\n\t"
        "mad.wide.u32 acc, cl, dl, acc;\n\t"
        "mov.b64 {rl, rh}, acc;\n\t"
        "mad.lo.u32 rh, cl, dh, rh;\n\t"
        "mad.lo.u32 rh, ch, dl, rh;\n\t"
        "mov.b64 %0, {rl, rh};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b), "l"(c), "l"(d));
    return r;
#else
    return a * b + c * d;
#endif
}

// Approximate high product: drops lo*lo and the carries out of the low words, so the result is in {exact-2 .. exact}.
// 3 IMAD.WIDE.U32 + 3 IADD3 (the exact mul.hi.u64 is 4 + 3).  The C fallback computes the identical value.
__host__ __device__ __forceinline__ u64 mulhi64_approx(u64 y, u64 s)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 yl, yh, sl, sh, a1, b1, rl, rh, dm;\n\t"
        ".reg .u64 t1, t2, t3;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {sl, sh}, %2;\n\t"
        "mul.wide.u32 t1, yh, sl;\n\t"
        "mul.wide.u32 t2, yl, sh;\n\t"
        "mul.wide.u32 t3, yh, sh;\n\t"
        "mov.b64 {dm, a1}, t1;\n\t"
        "mov.b64 {dm, b1}, t2;\n\t"
        "mov.b64 {rl, rh}, t3;\n\t"
        "add.cc.u32 rl, rl, a1;\n\t"
        "addc.u32 rh, rh, 0;\n\t"
        "add.cc.u32 rl, rl, b1;\n\t"
        "addc.u32 rh, rh, 0;\n\t"
        "mov.b64 %0, {rl, rh};\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(s));
    return r;
#else
    const u64 yl = (u32)y, yh = y >> 32, sl = (u32)s, sh = s >> 32;
    return yh * sh + ((yh * sl) >> 32) + ((yl * sh) >> 32);
#endif
}

// Shoup multiplication with the approximate quotient: result in [0, 4q) for any 64-bit y.
__host__ __device__ __forceinline__ u64 shoup_mul_a(u64 y, u64 w, u64 ws, u64 negq)
{
    return mullo_sum2(y, w, mulhi64_approx(y, ws), negq);
}

// Shoup multiplication on the fused chain: negq = 2^64 - q.  Result in [0, 2q).
__host__ __device__ __forceinline__ u64 shoup_mul_n(u64 y, u64 w, u64 ws, u64 negq)
{
    return mullo_sum2(y, w, mulhi64(y, ws), negq);
}

// low 64 bits of a*b + c*d: the four cross terms go to a separate 32-bit accumulator and the two lo*lo products to ONE
// 64-bit mad.wide chain (an aligned register pair from start to end), joined by a single carry-less 32-bit add.
// 2 IMAD.WIDE.U32 + 4 IMAD + 1 IADD3 (the split-accumulator version above costs ptxas 2 carry IADD3 more).
__host__ __device__ __forceinline__ u64 mullo_sum2_x(u64 a, u64 b, u64 c, u64 d)
{
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, cl, ch, dl, dh, rl, rh, h;\n\t"
        ".reg .u64 acc;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "mov.b64 {cl, ch}, %3;\n\t"
        "mov.b64 {dl, dh}, %4;\n\t"
        "mul.lo.u32 h, al, bh;\n\t"
        "mad.lo.u32 h, ah, bl, h;\n\t"
        "mul.wide.u32 acc, al, bl;\n\t"
        "mad.lo.u32 h, cl, dh, h;\n\t"
        "mad.lo.u32 h, ch, dl, h;\n\t"
        "mad.wide.u32 acc, cl, dl, acc;\n\t"
        "mov.b64 {rl, rh}, acc;\n\t"
        "add.u32 rh, rh, h;\n\t"
        "mov.b64 %0, {rl, rh};\n\t"
        "}"
        : "=l"(r)
        : "l"(a), "l"(b), "l"(c), "l"(d));
    return r;
#else
    return a * b + c * d;
#endif
}

// Shoup multiplication with the approximate quotient and the chain low product: result in [0, 4q) for any 64-bit y.
__host__ __device__ __forceinline__ u64 shoup_mul_m(u64 y, u64 w, u64 ws, u64 negq)
{
    return mullo_sum2_x(y, w, mulhi64_approx(y, ws), negq);
}

// low 64 bits of ((hi:lo) >> s), 0 <= s <= 64
__host__ __device__ __forceinline__ u64 shr128_lo(u64 hi, u64 lo, int s)
{
    if (s <= 0) return lo;
    if (s >= 64) return s >= 128 ? 0 : hi >> (s - 64);
    return (lo >> s) | (hi << (64 - s));
}

// The reference's Barrett reduction of the 128-bit product a*b, operation for operation
// (ntt_60bit.cuh:44-61; same sequence inlined at poly_arithmetic.cuh:16-33).  Canonical for a*b < 2^(2*qbit).
__host__ __device__ __forceinline__ u64 barrett_ref(u64 a, u64 b, u64 q, u64 mu, int qbit)
{
    u64 lo = a * b, hi = mulhi64(a, b);
    u64 x1 = shr128_lo(hi, lo, qbit - 2);
    u64 plo = x1 * mu, phi = mulhi64(x1, mu);
    u64 x2 = shr128_lo(phi, plo, qbit + 2);
    u64 r = lo - x2 * q;
    return r >= q ? r - q : r;
}

// Product of two canonical residues without any precomputed companion: the reference's Barrett quotient estimate (never more
// than 2 short) without its final correction.  Result in [0, 3q), a*b < 2^(2*qbit).
__host__ __device__ __forceinline__ u64 barrett_lazy(u64 a, u64 b, u64 q, u64 mu, int qbit)
{
    u64 lo = a * b, hi = mulhi64(a, b);
    u64 x1 = shr128_lo(hi, lo, qbit - 2);
    u64 plo = x1 * mu, phi = mulhi64(x1, mu);
    u64 x2 = shr128_lo(phi, plo, qbit + 2);
    return lo - x2 * q;
}

// (x * 2^-1) mod q for canonical x, as the reference does after every inverse stage (ntt_60bit.cuh:165, 494-513)
__host__ __device__ __forceinline__ u64 half_mod(u64 x, u64 q2) { return (x >> 1) + (q2 & (0 - (x & 1))); }

// The reference's ternary converter, bfv_keygen.cuh:18-30 / bfv_encryption.cuh:23-36: int(float(byte) / (255.0f/3)) - 1 in {-1, 0, 1, 2};
// negative -> q - 1.  255.0f/3 is exactly 85.0f and byte/85.0f crosses an integer only AT 85, 170 and 255 (84/85, 169/85, 254/85 are
// nowhere near one), so the float expression equals the threshold count below for all 256 bytes
// (tests/test_oracle_golden.py::test_ternary_thresholds_equal_the_float_formula checks all of them against the float formula):
// three compares instead of an IEEE division with its slow-path call.
__host__ __device__ __forceinline__ u64 ternary_value(unsigned char byte, u64 q)
{
    const int b = (int)(byte >= 85) + (int)(byte >= 170) + (int)(byte == 255) - 1;
    return (u64)(b < 0) * q + (u64)(long long)b;
}
// the literal float formula (kept for the exhaustive equivalence test and the emulator export)
__host__ __device__ __forceinline__ int ternary_float_formula(unsigned char byte)
{
    float d = (float)byte;
    d /= (255.0f / 3);
    return int(d) - 1;
}

}  // namespace nttb200
