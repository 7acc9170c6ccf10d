// epi.cuh -- epilogues fused into the store of the last inverse NTT kernel (BFV encryption's element-wise tail).
#pragma once
#include "modarith.cuh"

namespace nttb200 {

// ---- epilogues fused into the store of the LAST inverse kernel (strided pass, final round) --------------------------------
// BFV encryption (bfv_encryption.cuh:111-212) consumes the INTT outputs coefficient by coefficient: `+ e`, the rounding of the
// dropped limb, the modulus switch and `Delta * m`.  The thread that finishes the last inverse stage holds 16 canonical
// coefficients in registers, so those steps run there and the ciphertext is written ONCE (the separate epilogue pass re-read
// and re-wrote all of c: 15.5 MiB of the 49 MiB an encryption at (32768, 16 limbs) moved).  Ordering: the dropped limb is
// finished by an earlier launch (mode kEpiEncLast), the other limbs read its finished values `cl` (mode kEpiEncLimb).
struct EncEpiLimb { u64 q, twoq, inv_q_last, inv_q_last_s, qdt, bias, ratio, reduce_cl; };   // per limb below the dropped one; reduce_cl: q_last > 2 q
enum { kEpiNone = 0, kEpiEncLast = 1, kEpiEncLimb = 2 };
struct EpiArgs {
    const signed char *es;        // gaussian draws es[item][2][n] (signed, |e| <= 19)
    const u64 *cl;                // kEpiEncLimb: finished dropped-limb values cl[item][half][n]
    size_t cl_item_stride, cl_half_stride;
    const u64 *m;                 // plaintexts m[item][n] (half 0 only)
    size_t m_stride;
    const EncEpiLimb *K;          // [r-1], indexed by the GLOBAL limb
    u64 last, half_last, t, tfix; // dropped modulus, floor(last / 2), plaintext modulus (a power of two), (t + 1) / 2
    u32 tsh, first_limb;          // log2 t; global limb of the launch's local limb 0
};
// last limb of one half: += e (`>` quirk, bfv_encryption.cuh:187), += floor(q_last / 2) mod q_last (:121-124)
__host__ __device__ __forceinline__ u64 enc_last_limb_value(u64 v, int d, u64 last, u64 half_last)
{
    u64 x = v + (d < 0 ? last + (u64)(long long)d : (u64)d);
    if (x > last) x -= last;
    x += half_last;
    if (x >= last) x -= last;
    return x;
}
struct NoEpi {
    static constexpr int kMode = kEpiNone;
    __device__ __forceinline__ void init(const EpiArgs &, u32, u32, u32) {}
    __device__ __forceinline__ u64 apply(u64 v, u32) const { return v; }
};
// grp = item * 2 + half (the launch's groups are the (item, half) pairs), j = coefficient index
struct EncLastEpi {
    static constexpr int kMode = kEpiEncLast;
    const signed char *e;
    u64 last, half_last;
    __device__ __forceinline__ void init(const EpiArgs &E, u32 grp, u32, u32 n)
    {
        e = E.es + (size_t)grp * n;
        last = E.last; half_last = E.half_last;
    }
    __device__ __forceinline__ u64 apply(u64 v, u32 j) const { return enc_last_limb_value(v, (int)e[j], last, half_last); }
};
struct EncLimbEpi {
    static constexpr int kMode = kEpiEncLimb;
    const signed char *e;
    const u64 *cl, *m;
    u64 q, twoq, c, cs, qdt, bias, ratio, t, tfix;
    u32 tsh, reduce_cl;
    __device__ __forceinline__ void init(const EpiArgs &E, u32 grp, u32 limb, u32 n)
    {
        const u32 item = grp >> 1, half = grp & 1u;
        e = E.es + (size_t)grp * n;
        cl = E.cl + (size_t)item * E.cl_item_stride + (size_t)half * E.cl_half_stride;
        m = half == 0 ? E.m + (size_t)item * E.m_stride : nullptr;
        const EncEpiLimb &k = E.K[E.first_limb + limb];
        q = k.q; twoq = k.twoq; c = k.inv_q_last; cs = k.inv_q_last_s; qdt = k.qdt; bias = k.bias; ratio = k.ratio; reduce_cl = (u32)k.reduce_cl;
        t = E.t; tfix = E.tfix; tsh = E.tsh;
    }
    // (c_i + e - (c_last - half)) * q_last^-1 [+ Delta*m + round-fix]: k_encrypt_epilogue's ALL_LAZY arithmetic (bfv_kernels.cuh)
    __device__ __forceinline__ u64 apply(u64 v, u32 j) const
    {
        const int d = (int)e[j];
        u64 last = cl[j];
        if (reduce_cl) last -= mulhi64(last, ratio) * q;          // q_last > 2 q_i (mixed-size sets such as 16k_9q): bring c_last below 2 q_i first
        u64 x = shoup_mul(v + bias + (u64)(long long)d - last, c, cs, q);          // [0, 2q)
        if (m) {
            const u64 mj = m[j], f = (mj + tfix) >> tsh;
            if (mj < t) x = csub(x + (mj * qdt + f), twoq);
            else { const u64 y = csub(x, q) + (mj * qdt + f); x = csub(y - mulhi64(y, ratio) * q, q); }
        }
        return csub(x, q);
    }
};

}  // namespace nttb200
