// bfv_kernels.cuh -- Salsa20/20 keystream, distribution converters and the fused BFV pipeline kernels.
//
// Replaces distributions.cuh:48-218 (VecCrypt + convert_*), bfv_keygen.cuh:14-93, bfv_encryption.cuh:17-212 and
// bfv_decryption.cuh:13-57 together with the base-conversion / rounding kernels they call (poly_arithmetic.cuh:217-263).
// Arithmetic is the reference's, operation for operation (float/double conversions included; compiled WITHOUT
// fast-math so float division and normcdfinvf behave exactly like the reference build).  What changes is structure:
//   * every kernel takes a batch of keys / ciphertexts (item k draws Salsa20 nonce nonce0 + k; nonce 0 == reference),
//   * the element-wise chains between transforms are fused: encryption's c += e, mod-switch of the last limb and the
//     Delta*m term (4 reference launches, 12 passes over c and e) are one pass; decryption's c1 += c0, two scalar
//     multiplications, both base conversions, both final multiplications and the rounding (8 launches on 3 streams with
//     a latent race) are one pass that writes only the n plaintext coefficients,
//   * the gaussian draws are kept as n small signed integers per polynomial instead of r*n 64-bit residues.
#pragma once
#include "modarith.cuh"
#include "pointwise_kernels.cuh"

namespace nttb200 {

// ---- Salsa20/20 -------------------------------------------------------------------------------------------------------
struct SalsaKey { u32 k[8]; };

__host__ __device__ __forceinline__ u32 rotl32(u32 v, int c) { return (v << c) | (v >> (32 - c)); }

// One 64-byte keystream block: state layout and round structure of distributions.cuh:59-133 (standard Salsa20/20:
// "expand 32-byte k", 64-bit nonce, 64-bit block counter).
__host__ __device__ __forceinline__ void salsa20_block(u32 (&o)[16], const SalsaKey &key, u64 nonce, u64 blk)
{
    u32 j[16], x[16];
    j[0] = 0x61707865u; j[5] = 0x3320646eu; j[10] = 0x79622d32u; j[15] = 0x6b206574u;
    j[1] = key.k[0]; j[2] = key.k[1]; j[3] = key.k[2]; j[4] = key.k[3];
    j[11] = key.k[4]; j[12] = key.k[5]; j[13] = key.k[6]; j[14] = key.k[7];
    j[6] = (u32)nonce; j[7] = (u32)(nonce >> 32); j[8] = (u32)blk; j[9] = (u32)(blk >> 32);
    NTT_UNROLL
    for (int i = 0; i < 16; i++) x[i] = j[i];
#define NTT_QR(a, b, c, d)              \
    x[b] ^= rotl32(x[a] + x[d], 7);     \
    x[c] ^= rotl32(x[b] + x[a], 9);     \
    x[d] ^= rotl32(x[c] + x[b], 13);    \
    x[a] ^= rotl32(x[d] + x[c], 18);
    for (int r = 0; r < 10; r++) {
        NTT_QR(0, 4, 8, 12) NTT_QR(5, 9, 13, 1) NTT_QR(10, 14, 2, 6) NTT_QR(15, 3, 7, 11)
        NTT_QR(0, 1, 2, 3) NTT_QR(5, 6, 7, 4) NTT_QR(10, 11, 8, 9) NTT_QR(15, 12, 13, 14)
    }
#undef NTT_QR
    NTT_UNROLL
    for (int i = 0; i < 16; i++) o[i] = x[i] + j[i];
}

// VecCrypt distributions.cuh:48-155 on a zeroed buffer == the raw keystream.  `streams` independent streams of
// blocks_per_stream blocks each; stream s uses nonce0 + s and starts at out + s * stream_stride (bytes).
NTT_KERNEL void k_salsa20_keystream(unsigned char *out, u64 blocks_per_stream, u64 streams, size_t stream_stride, SalsaKey key, u64 nonce0)
{
    NTT_GRID_STRIDE(i, blocks_per_stream * streams) {
        const u64 s = i / blocks_per_stream, b = i - s * blocks_per_stream;
        u32 o[16];
        salsa20_block(o, key, nonce0 + s, b);
        uint4 *dst = reinterpret_cast<uint4 *>(out + s * stream_stride + b * 64);
        NTT_UNROLL
        for (int v = 0; v < 4; v++) { uint4 t; t.x = o[4 * v]; t.y = o[4 * v + 1]; t.z = o[4 * v + 2]; t.w = o[4 * v + 3]; dst[v] = t; }
    }
}

// ---- distribution converters: the reference's formulas -----------------------------------------------------------------
// bfv_keygen.cuh:18-30 / bfv_encryption.cuh:23-36: int(float(byte) / (255.0f/3)) - 1 in {-1, 0, 1, 2}; negative -> q - 1
__host__ __device__ __forceinline__ u64 ternary_value(unsigned char byte, u64 q)
{
    float d = (float)byte;
    d /= (255.0f / 3);
    int b = int(d) - 1;
    return (u64)(b < 0) * q + (u64)(long long)b;
}
// bfv_keygen.cuh:37-44 / distributions.cuh:195-201
__host__ __device__ __forceinline__ u64 uniform_value(u64 x, u64 q)
{
    double d = (double)x;
    d /= 18446744073709551615ULL;
    d *= (double)(q - 1);
    return (u64)d;
}
// bfv_keygen.cuh:51-73 / distributions.cuh:161-183 / bfv_encryption.cuh:49-71: the signed draw dd
__device__ __forceinline__ int gaussian_value(u32 x)
{
    float d = x;
    d /= 4294967295;
    if (d == 0)
        d += 1.192092896e-07F;
    else if (d == 1)
        d -= 1.192092896e-07F;
    d = normcdfinvf(d);
    d = d * (float)3.2 + 0;
    if (d > 19.2) {
        d = 19.2;
    } else if (d < -19.2) {
        d = -19.2;
    }
    return (int)d;
}
__host__ __device__ __forceinline__ u64 signed_to_residue(int dd, u64 q) { return dd < 0 ? q + (u64)(long long)dd : (u64)dd; }

// legacy single-limb converters, distributions.cuh:157-218
NTT_KERNEL void k_convert_gaussian(const u32 *in, u64 *out, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) out[i] = signed_to_residue(gaussian_value(in[i]), q);
}
NTT_KERNEL void k_convert_range(const u64 *in, u64 *out, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) out[i] = uniform_value(in[i], q);
}
NTT_KERNEL void k_convert_ternary(const unsigned char *in, u64 *out, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) {
        float d = (float)in[i];
        d /= (256.0f / 3);
        out[i] = d >= 2 ? 1 : (d >= 1 ? 0 : q - 1);
    }
}
// bfv_keygen.cuh:14-79 ("_xq": all limbs in one launch, limb = i / n)
NTT_KERNEL void k_ternary_dist_xq(const unsigned char *in, u64 *sk, unsigned n, size_t total, const u64 *q)
{
    NTT_GRID_STRIDE(i, total) sk[i] = ternary_value(in[i % n], q[i / n]);
}
NTT_KERNEL void k_uniform_dist_xq(const unsigned char *in, u64 *pk, unsigned n, size_t total, const u64 *q)
{
    NTT_GRID_STRIDE(i, total) pk[i] = uniform_value(reinterpret_cast<const u64 *>(in)[i], q[i / n]);
}
NTT_KERNEL void k_gaussian_dist_xq(const unsigned char *in, u64 *temp, unsigned n, size_t total, const u64 *q)
{
    NTT_GRID_STRIDE(i, total) temp[i] = signed_to_residue(gaussian_value(reinterpret_cast<const u32 *>(in)[i % n]), q[i / n]);
}
// bfv_encryption.cuh:17-109 convert_ternary_gaussian_x2: u into both halves of c, e0 / e1 into the halves of e
NTT_KERNEL void k_convert_ternary_gaussian_x2(const unsigned char *in, u64 *c, u64 *e, unsigned n, unsigned q_amount, const u64 *q)
{
    const size_t rn = (size_t)n * q_amount;
    NTT_GRID_STRIDE(i, rn) {
        const u64 qi = q[i / n];
        const u64 tv = ternary_value(in[i % n], qi);
        c[i] = tv; c[i + rn] = tv;
        e[i] = signed_to_residue(gaussian_value(reinterpret_cast<const u32 *>(in + n)[i % n]), qi);
        e[i + rn] = signed_to_residue(gaussian_value(reinterpret_cast<const u32 *>(in + (size_t)n * 5)[i % n]), qi);
    }
}

// ---- fused pipeline kernels (batched) ------------------------------------------------------------------------------------
// Layouts per item k: sk[r][n], pk[2][r][n] = [pk0 | pk1 = a], c[2][r][n]; keystream `in` as in the reference:
// keygen  (bfv_keygen.cuh:120-122):  bytes [0,n) ternary | u64 at n + 8*(l*n + j) uniform | u32 at n + 8rn + 4j gaussian
// encrypt (bfv_encryption.cuh:23,49,79): bytes [0,n) ternary u | u32 at n + 4j -> e0 | u32 at 5n + 4j -> e1

// keygen sampling: sk = ternary (all limbs), pk1 = uniform, es[k][j] = gaussian draw (signed)
NTT_KERNEL void k_keygen_sample(const unsigned char *in, size_t in_stride, u64 *sk, u64 *pk, int *es, unsigned n, unsigned r,
                                unsigned batch, const u64 *q)
{
    const size_t rn = (size_t)r * n;
    NTT_GRID_STRIDE(i, (size_t)batch * n) {
        const size_t k = i / n, j = i - k * n;
        const unsigned char *s = in + k * in_stride;
        const unsigned char byte = s[j];
        es[i] = gaussian_value(reinterpret_cast<const u32 *>(s + n + 8 * rn)[j]);
        for (unsigned l = 0; l < r; l++) {
            const u64 ql = q[l];
            sk[k * rn + (size_t)l * n + j] = ternary_value(byte, ql);
            pk[k * 2 * rn + rn + (size_t)l * n + j] = uniform_value(reinterpret_cast<const u64 *>(s + n)[(size_t)l * n + j], ql);
        }
    }
}
// pk0 = pk1 (.) sk   (barrett_batch_3param, bfv_keygen.cuh:132)
NTT_KERNEL void k_keygen_mul(u64 *pk, const u64 *sk, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    const size_t rn = (size_t)r * n;
    NTT_GRID_STRIDE(i, (size_t)batch * rn) {
        const size_t k = i / rn, rem = i - k * rn;
        const unsigned l = (unsigned)(rem / n);
        pk[k * 2 * rn + rem] = barrett_ref(pk[k * 2 * rn + rn + rem], sk[i], L.q[l], L.mu[l], (int)L.qbit[l]);
    }
}
// pk0 = -(pk0 + e)   (gaussian_dist_xq + poly_add_negate_xq, bfv_keygen.cuh:47-93)
NTT_KERNEL void k_keygen_add_negate(u64 *pk, const int *es, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    const size_t rn = (size_t)r * n;
    NTT_GRID_STRIDE(i, (size_t)batch * rn) {
        const size_t k = i / rn, rem = i - k * rn;
        const unsigned l = (unsigned)(rem / n);
        const u64 q = L.q[l];
        u64 ra = pk[k * 2 * rn + rem] + signed_to_residue(es[k * n + rem % n], q);
        if (ra >= q) ra -= q;
        ra = q - ra;
        pk[k * 2 * rn + rem] = ra * (u64)(ra != q);
    }
}

// encryption sampling: u (ternary) into half 0 of c for every limb; es[k][0][j], es[k][1][j] = the two gaussian draws
NTT_KERNEL void k_encrypt_sample(const unsigned char *in, size_t in_stride, u64 *c, int *es, unsigned n, unsigned r, unsigned batch,
                                 const u64 *q)
{
    const size_t rn = (size_t)r * n;
    NTT_GRID_STRIDE(i, (size_t)batch * n) {
        const size_t k = i / n, j = i - k * n;
        const unsigned char *s = in + k * in_stride;
        const unsigned char byte = s[j];
        es[k * 2 * n + j] = gaussian_value(reinterpret_cast<const u32 *>(s + n)[j]);
        es[k * 2 * n + n + j] = gaussian_value(reinterpret_cast<const u32 *>(s + (size_t)n * 5)[j]);
        for (unsigned l = 0; l < r; l++) c[k * 2 * rn + (size_t)l * n + j] = ternary_value(byte, q[l]);
    }
}
// c0 = NTT(u) (.) pk0, c1 = NTT(u) (.) pk1 -- the reference transforms u twice; here NTT(u) sits in
// half 0 and is read once.  pk_stride = 0: one public key for the whole batch.
NTT_KERNEL void k_encrypt_mul(u64 *c, const u64 *pk, size_t pk_stride, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    const size_t rn = (size_t)r * n;
    NTT_GRID_STRIDE(i, (size_t)batch * rn) {
        const size_t k = i / rn, rem = i - k * rn;
        const unsigned l = (unsigned)(rem / n);
        const u64 q = L.q[l], mu = L.mu[l];
        const int qb = (int)L.qbit[l];
        const u64 *pkk = pk + k * pk_stride;
        const u64 uh = c[k * 2 * rn + rem];
        c[k * 2 * rn + rem] = barrett_ref(uh, pkk[rem], q, mu, qb);
        c[k * 2 * rn + rn + rem] = barrett_ref(uh, pkk[rn + rem], q, mu, qb);
    }
}
// poly_add_xq + divide_and_round_q_last_inplace_add_x2 + ..._loop_xq + weird_m_stuff (bfv_encryption.cuh:111-212)
// in one pass; thread = (item, half, coefficient).  The dropped limb r-1 keeps the value the reference leaves there.
NTT_KERNEL void k_encrypt_epilogue(u64 *c, const int *es, const u64 *m_poly, size_t m_stride, unsigned n, unsigned r, unsigned batch, u64 t,
                                   const u64 *qi_div_t, LimbArrays L)
{
    const size_t rn = (size_t)r * n;
    const u64 last = L.q[r - 1], half_last = last >> 1;
    NTT_GRID_STRIDE(i, (size_t)batch * 2 * n) {
        const size_t k = i / (2 * (size_t)n), rem = i - k * 2 * n;
        const unsigned h = (unsigned)(rem / n);
        const size_t j = rem - (size_t)h * n;
        u64 *ch = c + k * 2 * rn + (size_t)h * rn;
        const int dd = es[k * 2 * n + (size_t)h * n + j];
        // last limb: += e (`>` quirk, :187), += floor(q_last / 2) mod q_last (:121-124)
        u64 cl = ch[(size_t)(r - 1) * n + j] + signed_to_residue(dd, last);
        if (cl > last) cl -= last;
        cl += half_last;
        if (cl >= last) cl -= last;
        ch[(size_t)(r - 1) * n + j] = cl;
        u64 m = 0, fix = 0;
        if (h == 0) { m = m_poly[k * m_stride + j]; fix = (m + ((t + 1) >> 1)) / t; }
        for (unsigned l = 0; l + 1 < r; l++) {
            const u64 q = L.q[l];
            u64 x = ch[(size_t)l * n + j] + signed_to_residue(dd, q);
            if (x > q) x -= q;
            const u64 half_mod = half_last % q;
            u64 tp = cl % q;
            if (tp < half_mod) tp += q;
            tp -= half_mod;
            if (x < tp) x += q;
            x -= tp;
            x = barrett_ref(x, L.inv_q_last_mod_q[l], q, L.mu[l], (int)L.qbit[l]);
            if (h == 0) x = (x + ((m * qi_div_t[l]) + fix)) % q;
            ch[(size_t)l * n + j] = x;
        }
    }
}

// c1 = NTT(c1) (.) sk    (barrett_batch, bfv_decryption.cuh:100).  c1 of item k = c + k*item_stride + c1_off.
NTT_KERNEL void k_decrypt_mul(u64 *c, size_t item_stride, size_t c1_off, const u64 *sk, size_t sk_stride, unsigned n, unsigned rp,
                              unsigned batch, LimbArrays L)
{
    const size_t rn = (size_t)rp * n;
    NTT_GRID_STRIDE(i, (size_t)batch * rn) {
        const size_t k = i / rn, rem = i - k * rn;
        const unsigned l = (unsigned)(rem / n);
        u64 *p = c + k * item_stride + c1_off + rem;
        *p = barrett_ref(*p, sk[k * sk_stride + rem], L.q[l], L.mu[l], (int)L.qbit[l]);
    }
}
struct DecryptConsts {
    u64 t, gamma, mu_gamma, gamma_div_2, neg_inv_t, neg_inv_gamma;
    int gamma_bits;
    unsigned rp;          // limbs after the drop (the driver's q_amount after q_amount--)
    const u64 *bcm;       // [2][rp]: prod_{i != j} q_i mod t | mod gamma   (demo.cu:248-264)
};
// poly_add_xq_d, poly_mul_int_xq_prodtgamma, poly_mul_int_xq_invpq (bfv_decryption.cuh:13-57), fast_convert_array_kernel_t,
// _gamma (poly_arithmetic.cuh:217-251), mod_t, barrett_int, dec_round_kernel (:128-141, :100-126, :253-263) in one pass.
// Writes n plaintext coefficients per item to out + k*out_stride.
NTT_KERNEL void k_decrypt_epilogue(const u64 *c, size_t item_stride, size_t c1_off, u64 *out, size_t out_stride, unsigned n, unsigned batch,
                                   DecryptConsts D, LimbArrays L)
{
    const u32 mask32 = (u32)(D.t - 1);
    NTT_GRID_STRIDE(i, (size_t)batch * n) {
        const size_t k = i / n, j = i - k * n;
        const u64 *c0 = c + k * item_stride, *c1 = c0 + c1_off;
        u64 acc_t = 0, acc_g = 0;
        for (unsigned l = 0; l < D.rp; l++) {
            const u64 q = L.q[l], mu = L.mu[l];
            const int qb = (int)L.qbit[l];
            u64 v = c1[(size_t)l * n + j] + c0[(size_t)l * n + j];
            if (v > q) v -= q;
            v = barrett_ref(v, L.prod_t_gamma_mod_q[l], q, mu, qb);
            v = barrett_ref(v, L.inv_punctured_q[l], q, mu, qb);
            acc_t += (v * D.bcm[l]) & (u64)mask32;
            acc_g = (acc_g + barrett_ref(v, D.bcm[l + D.rp], D.gamma, D.mu_gamma, D.gamma_bits)) % D.gamma;
        }
        u64 mt = acc_t & (u64)mask32;
        u64 mg = acc_g % D.gamma;
        mt = (mt * D.neg_inv_t) & (u64)mask32;                                   // mod_t
        mg = barrett_ref(mg, D.neg_inv_gamma, D.gamma, D.mu_gamma, D.gamma_bits);   // barrett_int
        const u64 tmask = D.t - 1;
        out[k * out_stride + j] = mg > D.gamma_div_2 ? ((mt + (D.gamma - mg)) & tmask) : ((mt - mg) & tmask);
    }
}

// ---- limb-sharded decryption (multi-GPU): the same arithmetic split at its only cross-limb reduction -------------------
// Each GPU owns `count` limbs [first, first+count) of every ciphertext as a compact shard c[item][2][count][n] and
// produces partial base-conversion sums  part[item][0][j] = sum (v_l * bcm_t[l] & mask)  (mod 2^64 wrap-around, masked at the end)
// and part[item][1][j] = sum Barrett_gamma(v_l * bcm_g[l]) mod gamma.  One all-reduce(SUM) of `part` across the GPUs
// (at most 8 addends < 2^61: no 64-bit overflow) followed by k_decrypt_finish reproduces k_decrypt_epilogue exactly:
// the reference's running `(acc + v) % gamma` and the plain modular sum are the same residue.
NTT_KERNEL void k_decrypt_partial(const u64 *c, size_t item_stride, size_t c1_off, u64 *part, unsigned n, unsigned batch, unsigned first,
                                  unsigned count, DecryptConsts D, LimbArrays L)
{
    const u32 mask32 = (u32)(D.t - 1);
    NTT_GRID_STRIDE(i, (size_t)batch * n) {
        const size_t k = i / n, j = i - k * n;
        const u64 *c0 = c + k * item_stride, *c1 = c0 + c1_off;
        u64 acc_t = 0, acc_g = 0;
        for (unsigned ll = 0; ll < count; ll++) {
            const unsigned l = first + ll;                 // global limb: constants are indexed globally, data locally
            const u64 q = L.q[l], mu = L.mu[l];
            const int qb = (int)L.qbit[l];
            u64 v = c1[(size_t)ll * n + j] + c0[(size_t)ll * n + j];
            if (v > q) v -= q;
            v = barrett_ref(v, L.prod_t_gamma_mod_q[l], q, mu, qb);
            v = barrett_ref(v, L.inv_punctured_q[l], q, mu, qb);
            acc_t += (v * D.bcm[l]) & (u64)mask32;
            acc_g = (acc_g + barrett_ref(v, D.bcm[l + D.rp], D.gamma, D.mu_gamma, D.gamma_bits)) % D.gamma;
        }
        part[k * 2 * n + j] = acc_t;
        part[k * 2 * n + n + j] = acc_g;
    }
}
NTT_KERNEL void k_decrypt_finish(const u64 *part_sum, u64 *out, size_t out_stride, unsigned n, unsigned batch, DecryptConsts D)
{
    const u32 mask32 = (u32)(D.t - 1);
    NTT_GRID_STRIDE(i, (size_t)batch * n) {
        const size_t k = i / n, j = i - k * n;
        u64 mt = part_sum[k * 2 * n + j] & (u64)mask32;
        u64 mg = part_sum[k * 2 * n + n + j] % D.gamma;
        mt = (mt * D.neg_inv_t) & (u64)mask32;
        mg = barrett_ref(mg, D.neg_inv_gamma, D.gamma, D.mu_gamma, D.gamma_bits);
        const u64 tmask = D.t - 1;
        out[k * out_stride + j] = mg > D.gamma_div_2 ? ((mt + (D.gamma - mg)) & tmask) : ((mt - mg) & tmask);
    }
}

}  // namespace nttb200
