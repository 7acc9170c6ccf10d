// pointwise_kernels.cuh -- coefficient-wise RNS kernels (HBM-bound: 16-24 bytes per modular multiplication).
//
// Replaces poly_arithmetic.cuh:9-263 and the "_xq" kernels of bfv_keygen.cuh:81-93, bfv_encryption.cuh:111-212,
// bfv_decryption.cuh:13-57.  Semantics are the reference's, quirk for quirk (the `>` comparisons of poly_add*,
// the no-op poly_sub, 32-bit masks in mod_t / fast_convert_t): these entry points are observable stand-alone.
// Differences are structural only: grid-stride loops over 16-byte vectors sized to the SM count instead of one
// 8-byte element per thread, limb constants read once per thread from global memory (no 16-limb __constant__ cap),
// and the fused variants further down do in one pass what the reference does in two or three.
#pragma once
#include "modarith.cuh"

namespace nttb200 {

// Per-limb constant arrays as the reference keeps them in __constant__ memory (ntt_60bit.cuh:8-13); any may be null
// where a kernel does not use it.
struct LimbArrays {
    const u64 *q, *mu;
    const u32 *qbit;
    const u64 *inv_q_last_mod_q, *inv_punctured_q, *prod_t_gamma_mod_q;
};

#define NTT_GRID_STRIDE(i, total) \
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (total); i += (size_t)gridDim.x * blockDim.x)

// ---- single-limb kernels (explicit q, mu, qbit) ---------------------------------------------------------------------
// barrett poly_arithmetic.cuh:9-34 (c == a), barrett_3param (c != a)
NTT_KERNEL void k_barrett(u64 *c, const u64 *a, const u64 *b, size_t n, u64 q, u64 mu, int qbit)
{
    NTT_GRID_STRIDE(i, n) c[i] = barrett_ref(a[i], b[i], q, mu, qbit);
}
// barrett_int poly_arithmetic.cuh:100-126
NTT_KERNEL void k_barrett_int(u64 *a, u64 b, size_t n, u64 q, u64 mu, int qbit)
{
    NTT_GRID_STRIDE(i, n) a[i] = barrett_ref(a[i], b, q, mu, qbit);
}
// mod_t poly_arithmetic.cuh:128-141 (the mask is a 32-bit unsigned, :139)
NTT_KERNEL void k_mod_t(u64 *a, u64 b, size_t n, u64 t)
{
    const u32 mask = (u32)(t - 1);
    NTT_GRID_STRIDE(i, n) a[i] = (a[i] * b) & (u64)mask;
}
// poly_add poly_arithmetic.cuh:143-153
NTT_KERNEL void k_poly_add(u64 *a, const u64 *b, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) { u64 r = a[i] + b[i]; if (r > q) r -= q; a[i] = r; }
}
// poly_add_integer poly_arithmetic.cuh:155-165
NTT_KERNEL void k_poly_add_integer(u64 *a, u64 b, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) { u64 r = a[i] + b; if (r > q) r -= q; a[i] = r; }
}
// poly_sub poly_arithmetic.cuh:167-178: adds q when a < b and never subtracts b (reference behaviour, kept)
NTT_KERNEL void k_poly_sub(u64 *a, const u64 *b, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) { u64 r = a[i]; if (r < b[i]) r += q; a[i] = r; }
}
// poly_negate poly_arithmetic.cuh:332-338
NTT_KERNEL void k_poly_negate(u64 *a, size_t n, u64 q)
{
    NTT_GRID_STRIDE(i, n) { u64 r = q - a[i]; a[i] = r * (u64)(r != q); }
}
// divide_and_round_q_last_inplace_loop poly_arithmetic.cuh:180-214
NTT_KERNEL void k_divide_and_round_q_last_inplace_loop(u64 *input_poly, const u64 *rns_poly_minus1, size_t n, u64 base_q_i, u64 half_mod,
                                                       u64 inv_q_last_mod_q_i, u64 mu, int qbit)
{
    NTT_GRID_STRIDE(i, n) {
        u64 t = rns_poly_minus1[i] % base_q_i;
        if (t < half_mod) t += base_q_i;
        t -= half_mod;
        u64 x = input_poly[i];
        if (x < t) x += base_q_i;
        x -= t;
        input_poly[i] = barrett_ref(x, inv_q_last_mod_q_i, base_q_i, mu, qbit);
    }
}
// fast_convert_array_kernel_t poly_arithmetic.cuh:217-234
NTT_KERNEL void k_fast_convert_t(const u64 *input_poly, u64 *result_poly, u64 t, const u64 *bcm, unsigned q_amount, size_t n)
{
    const u32 mask = (u32)(t - 1);
    NTT_GRID_STRIDE(k, n) {
        u64 acc = 0;
        for (unsigned i = 0; i < q_amount; i++) acc += (input_poly[k + (size_t)i * n] * bcm[i]) & (u64)mask;
        result_poly[k] = acc & (u64)mask;
    }
}
// fast_convert_array_kernel_gamma poly_arithmetic.cuh:237-251 (writes result_poly[k + n])
NTT_KERNEL void k_fast_convert_gamma(const u64 *input_poly, u64 *result_poly, u64 gamma, const u64 *bcm, unsigned q_amount, int gamma_bits,
                                     u64 mu_gamma, size_t n)
{
    NTT_GRID_STRIDE(k, n) {
        u64 acc = 0;
        for (unsigned i = 0; i < q_amount; i++) {
            u64 v = barrett_ref(input_poly[k + (size_t)i * n], bcm[i + q_amount], gamma, mu_gamma, gamma_bits);
            acc = (acc + v) % gamma;
        }
        result_poly[k + n] = acc % gamma;
    }
}
// dec_round_kernel poly_arithmetic.cuh:253-263
NTT_KERNEL void k_dec_round(const u64 *input_poly, u64 *result_poly, u64 t, u64 gamma, u64 gamma_div_2, size_t n)
{
    const u64 mask = t - 1;
    NTT_GRID_STRIDE(i, n) {
        const u64 a = input_poly[i], g = input_poly[i + n];
        result_poly[i] = g > gamma_div_2 ? ((a + (gamma - g)) & mask) : ((a - g) & mask);
    }
}

// ---- multi-limb kernels: element i of a [polys][n] array belongs to limb (i / n) % division ----------------------
// barrett_batch poly_arithmetic.cuh:36-66 (c == a) / barrett_batch_3param :68-98
NTT_KERNEL void k_barrett_batch(u64 *c, const u64 *a, const u64 *b, unsigned n, size_t total, unsigned division, LimbArrays L)
{
    NTT_GRID_STRIDE(i, total) {
        const unsigned l = (unsigned)((i / n) % division);
        c[i] = barrett_ref(a[i], b[i], L.q[l], L.mu[l], (int)L.qbit[l]);
    }
}
// poly_add_negate_xq bfv_keygen.cuh:81-93: a = -(a + b) mod q   (limb = i / n)
NTT_KERNEL void k_poly_add_negate_xq(u64 *a, const u64 *b, unsigned n, size_t total, LimbArrays L)
{
    NTT_GRID_STRIDE(i, total) {
        const u64 q = L.q[i / n];
        u64 r = a[i] + b[i];
        if (r >= q) r -= q;
        r = q - r;
        a[i] = r * (u64)(r != q);
    }
}
// poly_add_xq bfv_encryption.cuh:180-191: c += e on both halves ([2][q_amount][n]); `>` quirk
NTT_KERNEL void k_poly_add_xq(u64 *c, const u64 *e, unsigned n, unsigned q_amount, LimbArrays L)
{
    const size_t half = (size_t)n * q_amount;
    NTT_GRID_STRIDE(i, 2 * half) {
        const u64 q = L.q[(i % half) / n];
        u64 r = c[i] + e[i];
        if (r > q) r -= q;
        c[i] = r;
    }
}
// divide_and_round_q_last_inplace_add_x2 bfv_encryption.cuh:111-125: last limb of c0 and c1 += floor(q_last / 2)
NTT_KERNEL void k_divide_and_round_q_last_inplace_add_x2(u64 *c, unsigned n, unsigned q_amount, LimbArrays L)
{
    const u64 last = L.q[q_amount - 1], half = last >> 1;
    NTT_GRID_STRIDE(i, (size_t)2 * n) {
        const size_t k = (size_t)n * (q_amount - 1) + i % n + ((size_t)n * q_amount) * (size_t)(i >= n);
        u64 r = c[k] + half;
        if (r >= last) r -= last;
        c[k] = r;
    }
}
// divide_and_round_q_last_inplace_loop_xq bfv_encryption.cuh:127-178: limbs i < r-1 of c0 and c1
NTT_KERNEL void k_divide_and_round_q_last_inplace_loop_xq(u64 *c, unsigned q_amount, unsigned n, LimbArrays L)
{
    const u64 half_last = L.q[q_amount - 1] >> 1;
    const size_t per = (size_t)n * (q_amount - 1);
    NTT_GRID_STRIDE(i, 2 * per) {
        const size_t ii = i % n;
        const unsigned idx = (unsigned)((i % per) / n);
        const u64 q = L.q[idx];
        const u64 half_mod = half_last % q;
        const size_t second = (size_t)(i >= per);
        const u64 *lastp = c + second * ((size_t)n * q_amount) + (size_t)n * (q_amount - 1);
        u64 *inp = c + second * ((size_t)n * q_amount) + (size_t)n * idx;
        u64 t = lastp[ii] % q;
        if (t < half_mod) t += q;
        t -= half_mod;
        u64 x = inp[ii];
        if (x < t) x += q;
        x -= t;
        inp[ii] = barrett_ref(x, L.inv_q_last_mod_q[idx], q, L.mu[idx], (int)L.qbit[idx]);
    }
}
// weird_m_stuff bfv_encryption.cuh:193-212: c0_i += m * floor(q_i / t) + round-fix, i < r-1
NTT_KERNEL void k_weird_m_stuff(const u64 *m_poly, u64 *c0, u64 t, const u64 *qi_div_t, const u64 *q_array, unsigned q_amount, unsigned n)
{
    NTT_GRID_STRIDE(j, (size_t)n) {
        const u64 m = m_poly[j];
        const u64 fix = (m + ((t + 1) >> 1)) / t;
        for (unsigned i = 0; i + 1 < q_amount; i++) {
            const size_t k = j + (size_t)i * n;
            c0[k] = (c0[k] + ((m * qi_div_t[i]) + fix)) % q_array[i];
        }
    }
}
// poly_add_xq_d bfv_decryption.cuh:13-23: c1_i += c0_i (c1 at c + n * q_amount_plus1), `>` quirk
NTT_KERNEL void k_poly_add_xq_d(u64 *c, unsigned n, unsigned rp, unsigned q_amount_plus1, LimbArrays L)
{
    NTT_GRID_STRIDE(i, (size_t)n * rp) {
        const u64 q = L.q[i / n];
        u64 r = c[i + (size_t)n * q_amount_plus1] + c[i];
        if (r > q) r -= q;
        c[i + (size_t)n * q_amount_plus1] = r;
    }
}
// poly_mul_int_xq_prodtgamma / _invpq bfv_decryption.cuh:25-57: c_i *= k_i  (k = prod_t_gamma_mod_q or inv_punctured_q)
NTT_KERNEL void k_poly_mul_int_xq(u64 *c, unsigned n, size_t total, const u64 *k, LimbArrays L)
{
    NTT_GRID_STRIDE(i, total) {
        const unsigned l = (unsigned)(i / n);
        c[i] = barrett_ref(c[i], k[l], L.q[l], L.mu[l], (int)L.qbit[l]);
    }
}

}  // namespace nttb200
