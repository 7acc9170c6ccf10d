/*
 * ntt_oracle.c -- CPU restatement of the NTT-Cuda hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * link or call this file.  The product (libnttb200.so) never does.
 *
 * Every function restates, in plain scalar C (unsigned __int128), what one reference kernel or host
 * routine computes, *including its quirks*, and cites the reference file:line it follows
 * (paths relative to the reference's BFV_Scheme/ directory).
 *
 * Parity pin: the decryption known-answer vector embedded in the reference's decryption_test.cu
 * (c_host :348, sk_host :355, expected plaintext i % 10 :230-232) -- see tests/test_oracle_golden.py.
 * The gaussian converter (normcdfinvf) is the one stage a CPU cannot reproduce bit-for-bit:
 * "parity unpinned" on CPU for that stage only; it is pinned GPU-vs-GPU against the rebuilt
 * reference (oracle/_ref) instead.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef unsigned long long u64;
typedef unsigned __int128 u128;
typedef uint32_t u32;

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------------
 * L0: host number theory (helper.h, uint128.h host parts)
 * ---------------------------------------------------------------------------------------------- */

/* uint128.h:278-281 + :151-159 -- `x % y`; operator<(uint128_t,u64) is really "<=", so the
 * remainder of x == y is y, not 0.  Otherwise an ordinary remainder. */
static u64 ref_mod128(u128 x, u64 y)
{
    if ((u64)(x >> 64) == 0 && (u64)x <= y) return (u64)x;
    return (u64)(x % y);
}

/* helper.h:8-28 modpow128 (square-and-multiply, LSB first) */
ORC_API u64 orc_modpow(u64 a, u64 b, u64 mod)
{
    u64 res = 1;
    if (b & 1) res = a;
    while (b != 0) {
        b >>= 1;
        a = ref_mod128((u128)a * a, mod);
        if (b & 1) res = ref_mod128((u128)res * a, mod);
    }
    return res;
}

/* helper.h:52-56 modinv128 (Fermat; also called with the non-prime t, demo.cu:109) */
ORC_API u64 orc_modinv(u64 a, u64 q) { return orc_modpow(a, q - 2, q); }

/* helper.h:58-70 bitReverse */
ORC_API u64 orc_bitrev(u64 a, int bits)
{
    u64 r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (a & 1); a >>= 1; }
    return r;
}

static int ilog2(u64 n) { int l = 0; while ((1ull << l) < n) l++; return l; }

/* parameter.h:5-12 fillTablePsi128: psiTable[i] = psi^bitrev(i), psiinvTable[i] = psiinv^bitrev(i).
 * (Computed here by repeated multiplication in natural order then scattered: identical values.) */
ORC_API void orc_fill_psi_tables(u64 psi, u64 q, u64 psiinv, u64 *psiTable, u64 *psiinvTable, unsigned n)
{
    int lg = ilog2(n);
    u64 p = 1, pi = 1;
    for (unsigned e = 0; e < n; e++) {
        unsigned i = (unsigned)orc_bitrev(e, lg);
        psiTable[i] = p;
        if (psiinvTable) psiinvTable[i] = pi;
        p = (u64)((u128)p * psi % q);
        pi = (u64)((u128)pi * psiinv % q);
    }
}

/* demo.cu:69 q_bit = log2((double)q) + 1 ; demo.cu:157-165 mu = floor(2^(2*qbit) / q) */
ORC_API unsigned orc_qbit(u64 q) { return (unsigned)(log2((double)q) + 1); }
ORC_API u64 orc_mu(u64 q, unsigned qbit)
{
    /* uint128_t::exp2(2*qbit) / q, low word (uint128.h:73-83, 245-276) */
    if (2 * qbit >= 128) return 0;
    return (u64)(((u128)1 << (2 * qbit)) / q);
}

/* ------------------------------------------------------------------------------------------------
 * L1: device modular arithmetic
 * ---------------------------------------------------------------------------------------------- */

/* ntt_60bit.cuh:44-61 singleBarrett, operation for operation (128-bit shifts keep the low word). */
static inline u64 ref_barrett(u128 a, u64 q, u64 mu, int qbit)
{
    u64 x = (u64)(a >> (qbit - 2));                 /* rx = a >> (qbit-2)        */
    u128 p = (u128)x * mu;                          /* mul64(rx.low, mu, rx)     */
    x = (u64)(p >> (qbit + 2));                     /* shiftr(rx, qbit+2)        */
    p = (u128)x * q;                                /* mul64(rx.low, q, rx)      */
    a -= p;                                         /* sub128(a, rx)             */
    u64 lo = (u64)a;
    if (lo >= q) lo -= q;
    return lo;
}
ORC_API u64 orc_barrett_mul(u64 a, u64 b, u64 q, u64 mu, int qbit) { return ref_barrett((u128)a * b, q, mu, qbit); }

/* ------------------------------------------------------------------------------------------------
 * L2: NTT / INTT  (ntt_60bit.cuh)
 * ---------------------------------------------------------------------------------------------- */

/* One polynomial, all stages.  Butterfly body: ntt_60bit.cuh:200-222 (CTBasedNTTInner) and the
 * identical body in :86-110 / :412-441; index scheme step = n/(2*length), psi index length+ps. */
ORC_API void orc_forward_ntt(u64 *a, unsigned n, u64 q, u64 mu, int qbit, const u64 *psi)
{
    for (unsigned length = 1; length < n; length <<= 1) {
        unsigned step = (n / length) / 2;
        for (unsigned tid = 0; tid < n / 2; tid++) {
            unsigned ps = tid / step;
            unsigned j = ps * step * 2 + tid % step;
            u64 w = psi[length + ps];
            u64 u = a[j];
            u64 v = ref_barrett((u128)a[j + step] * w, q, mu, qbit);
            u64 s = u + v;
            s -= q * (s >= q);
            a[j] = s;
            u += q * (u < v);
            a[j + step] = u - v;
        }
    }
}

/* ntt_60bit.cuh:233-264 (GSBasedINTTInner) / :144-180 / :483-514: GS butterfly, each output halved. */
ORC_API void orc_inverse_ntt(u64 *a, unsigned n, u64 q, u64 mu, int qbit, const u64 *psiinv)
{
    u64 q2 = (q + 1) >> 1;
    for (unsigned length = n / 2; length >= 1; length >>= 1) {
        unsigned step = (n / length) / 2;
        for (unsigned tid = 0; tid < n / 2; tid++) {
            unsigned ps = tid / step;
            unsigned j = ps * step * 2 + tid % step;
            u64 w = psiinv[length + ps];
            u64 u = a[j];
            u64 v = a[j + step];
            u64 s = u + v;
            s -= q * (s >= q);
            a[j] = (s >> 1) + q2 * (s & 1);
            u += q * (u < v);
            u64 d = ref_barrett((u128)(u - v) * w, q, mu, qbit);
            a[j + step] = (d >> 1) + q2 * (d & 1);
        }
    }
}

/* Fast host scalar NTT with the same moduli/tables for CPU timing (canonical inputs only: results are
 * identical to orc_forward_ntt because all stored values are canonical residues).  Uses `%` instead
 * of the Barrett sequence.  BASELINE.md section 2 "CPU scalar NTT". */
ORC_API void orc_forward_ntt_fast(u64 *a, unsigned n, u64 q, const u64 *psi)
{
    unsigned t = n;
    for (unsigned m = 1; m < n; m <<= 1) {
        t >>= 1;
        for (unsigned i = 0; i < m; i++) {
            u64 w = psi[m + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (unsigned j = 0; j < t; j++) {
                u64 u = x[j];
                u64 v = (u64)((u128)y[j] * w % q);
                u64 s = u + v; if (s >= q) s -= q;
                u64 d = u >= v ? u - v : u + q - v;
                x[j] = s; y[j] = d;
            }
        }
    }
}
ORC_API void orc_inverse_ntt_fast(u64 *a, unsigned n, u64 q, const u64 *psiinv)
{
    u64 q2 = (q + 1) >> 1;
    unsigned t = 1;
    for (unsigned m = n / 2; m >= 1; m >>= 1) {
        for (unsigned i = 0; i < m; i++) {
            u64 w = psiinv[m + i];
            u64 *x = a + 2 * i * t, *y = x + t;
            for (unsigned j = 0; j < t; j++) {
                u64 u = x[j], v = y[j];
                u64 s = u + v; if (s >= q) s -= q;
                u64 d = u >= v ? u - v : u + q - v;
                d = (u64)((u128)d * w % q);
                x[j] = (s >> 1) + q2 * (s & 1);
                y[j] = (d >> 1) + q2 * (d & 1);
            }
        }
        t <<= 1;
    }
}

/* ntt_60bit.cuh:608-650 forwardNTT_batch: poly p uses limb p % division, table psi + limb*n. */
ORC_API void orc_forward_ntt_batch(u64 *a, unsigned n, const u64 *psi, unsigned num, unsigned division,
                                   const u64 *q, const u64 *mu, const unsigned *qbit)
{
    for (unsigned p = 0; p < num; p++) {
        unsigned l = p % division;
        orc_forward_ntt(a + (size_t)p * n, n, q[l], mu[l], (int)qbit[l], psi + (size_t)l * n);
    }
}
/* ntt_60bit.cuh:652-697 inverseNTT_batch */
ORC_API void orc_inverse_ntt_batch(u64 *a, unsigned n, const u64 *psiinv, unsigned num, unsigned division,
                                   const u64 *q, const u64 *mu, const unsigned *qbit)
{
    for (unsigned p = 0; p < num; p++) {
        unsigned l = p % division;
        orc_inverse_ntt(a + (size_t)p * n, n, q[l], mu[l], (int)qbit[l], psiinv + (size_t)l * n);
    }
}

/* ------------------------------------------------------------------------------------------------
 * L2': pointwise kernels (poly_arithmetic.cuh)
 * ---------------------------------------------------------------------------------------------- */

/* poly_arithmetic.cuh:9-34 barrett: a[i] = a[i]*b[i] mod q */
ORC_API void orc_barrett(u64 *a, const u64 *b, unsigned n, u64 q, u64 mu, int qbit)
{
    for (unsigned i = 0; i < n; i++) a[i] = ref_barrett((u128)a[i] * b[i], q, mu, qbit);
}
/* poly_arithmetic.cuh:36-66 barrett_batch: grid.y = polys, limb = y % division, offset y*n on a AND b */
ORC_API void orc_barrett_batch(u64 *a, const u64 *b, unsigned n, unsigned polys, unsigned division,
                               const u64 *q, const u64 *mu, const unsigned *qbit)
{
    for (unsigned y = 0; y < polys; y++) {
        unsigned l = y % division;
        for (unsigned i = 0; i < n; i++) {
            size_t k = (size_t)y * n + i;
            a[k] = ref_barrett((u128)a[k] * b[k], q[l], mu[l], (int)qbit[l]);
        }
    }
}
/* poly_arithmetic.cuh:68-98 barrett_batch_3param: c = a*b */
ORC_API void orc_barrett_batch_3param(u64 *c, const u64 *a, const u64 *b, unsigned n, unsigned polys, unsigned division,
                                      const u64 *q, const u64 *mu, const unsigned *qbit)
{
    for (unsigned y = 0; y < polys; y++) {
        unsigned l = y % division;
        for (unsigned i = 0; i < n; i++) {
            size_t k = (size_t)y * n + i;
            c[k] = ref_barrett((u128)a[k] * b[k], q[l], mu[l], (int)qbit[l]);
        }
    }
}
/* poly_arithmetic.cuh:100-126 barrett_int */
ORC_API void orc_barrett_int(u64 *a, u64 b, unsigned n, u64 q, u64 mu, int qbit)
{
    for (unsigned i = 0; i < n; i++) a[i] = ref_barrett((u128)a[i] * b, q, mu, qbit);
}
/* poly_arithmetic.cuh:128-141 mod_t: mask is a 32-bit `unsigned` (:139) */
ORC_API void orc_mod_t(u64 *a, u64 b, unsigned n, u64 t)
{
    u32 mask = (u32)(t - 1);
    for (unsigned i = 0; i < n; i++) a[i] = (a[i] * b) & mask;
}
/* poly_arithmetic.cuh:143-153 poly_add: subtract q only if > q (result may equal q) */
ORC_API void orc_poly_add(u64 *a, const u64 *b, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) { u64 r = a[i] + b[i]; if (r > q) r -= q; a[i] = r; }
}
/* poly_arithmetic.cuh:155-165 poly_add_integer */
ORC_API void orc_poly_add_integer(u64 *a, u64 b, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) { u64 r = a[i] + b; if (r > q) r -= q; a[i] = r; }
}
/* poly_arithmetic.cuh:167-178 poly_sub: adds q when a<b and never subtracts b (reference bug, kept) */
ORC_API void orc_poly_sub(u64 *a, const u64 *b, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) { u64 r = a[i]; if (r < b[i]) r += q; a[i] = r; }
}
/* poly_arithmetic.cuh:332-338 poly_negate */
ORC_API void orc_poly_negate(u64 *a, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) { u64 r = q - a[i]; a[i] = r * (r != q); }
}
/* poly_arithmetic.cuh:180-214 divide_and_round_q_last_inplace_loop (per-limb mod-switch step) */
ORC_API void orc_divide_and_round_q_last_inplace_loop(u64 *input_poly, const u64 *rns_poly_minus1, unsigned n, u64 base_q_i,
                                                      u64 half_mod, u64 inv_q_last_mod_q_i, u64 mu, int qbit)
{
    for (unsigned i = 0; i < n; i++) {
        u64 t = rns_poly_minus1[i] % base_q_i;
        if (t < half_mod) t += base_q_i;
        t -= half_mod;
        u64 x = input_poly[i];
        if (x < t) x += base_q_i;
        x -= t;
        input_poly[i] = ref_barrett((u128)x * inv_q_last_mod_q_i, base_q_i, mu, qbit);
    }
}
/* poly_arithmetic.cuh:217-234 fast_convert_array_kernel_t (32-bit mask) */
ORC_API void orc_fast_convert_t(const u64 *input_poly, u64 *result_poly, u64 t, const u64 *bcm, unsigned q_amount, unsigned n)
{
    u32 mask = (u32)(t - 1);
    for (unsigned k = 0; k < n; k++) {
        u64 acc = 0;
        for (unsigned i = 0; i < q_amount; i++) acc += (input_poly[k + (size_t)i * n] * bcm[i]) & mask;
        result_poly[k] = acc & mask;
    }
}
/* poly_arithmetic.cuh:237-251 fast_convert_array_kernel_gamma -> result_poly[k + n] */
ORC_API void orc_fast_convert_gamma(const u64 *input_poly, u64 *result_poly, u64 gamma, const u64 *bcm, unsigned q_amount,
                                    int gamma_bits, u64 mu_gamma, unsigned n)
{
    for (unsigned k = 0; k < n; k++) {
        u64 acc = 0;
        for (unsigned i = 0; i < q_amount; i++) {
            u64 v = ref_barrett((u128)input_poly[k + (size_t)i * n] * bcm[i + q_amount], gamma, mu_gamma, gamma_bits);
            acc = (acc + v) % gamma;
        }
        result_poly[k + n] = acc % gamma;
    }
}
/* poly_arithmetic.cuh:253-263 dec_round_kernel */
ORC_API void orc_dec_round(const u64 *input_poly, u64 *result_poly, u64 t, u64 gamma, u64 gamma_div_2, unsigned n)
{
    u64 mask = t - 1;
    for (unsigned i = 0; i < n; i++) {
        if (input_poly[i + n] > gamma_div_2) result_poly[i] = (input_poly[i] + (gamma - input_poly[i + n])) & mask;
        else result_poly[i] = (input_poly[i] - input_poly[i + n]) & mask;
    }
}

/* ------------------------------------------------------------------------------------------------
 * L2'': Salsa20/20 keystream + converters (distributions.cuh, bfv_keygen.cuh, bfv_encryption.cuh)
 * ---------------------------------------------------------------------------------------------- */

static inline u32 rotl32(u32 u, int c) { return (u << c) | (u >> (32 - c)); }
static inline u32 le32(const unsigned char *x) { return (u32)x[0] | ((u32)x[1] << 8) | ((u32)x[2] << 16) | ((u32)x[3] << 24); }

/* distributions.cuh:48-155 VecCrypt with blks_per_chunk = 1: 64-byte block `blk` of the keystream
 * (buffer is zeroed first, :244/:273, so XOR == store).  State layout :59-80, rounds :82-115. */
static void salsa20_block(unsigned char out[64], const unsigned char key[32], u64 nonce, u64 blk)
{
    static const unsigned char sigma[16] = "expand 32-byte k";
    u32 j[16], x[16];
    j[0] = le32(sigma + 0);  j[1] = le32(key + 0);   j[2] = le32(key + 4);   j[3] = le32(key + 8);
    j[4] = le32(key + 12);   j[5] = le32(sigma + 4); j[6] = (u32)nonce;      j[7] = (u32)(nonce >> 32);
    j[8] = (u32)blk;         j[9] = (u32)(blk >> 32); j[10] = le32(sigma + 8); j[11] = le32(key + 16);
    j[12] = le32(key + 20);  j[13] = le32(key + 24); j[14] = le32(key + 28); j[15] = le32(sigma + 12);
    memcpy(x, j, sizeof x);
#define QR(a, b, c, d) x[b] ^= rotl32(x[a] + x[d], 7); x[c] ^= rotl32(x[b] + x[a], 9); \
                       x[d] ^= rotl32(x[c] + x[b], 13); x[a] ^= rotl32(x[d] + x[c], 18);
    for (int r = 0; r < 10; r++) {
        QR(0, 4, 8, 12) QR(5, 9, 13, 1) QR(10, 14, 2, 6) QR(15, 3, 7, 11)
        QR(0, 1, 2, 3)  QR(5, 6, 7, 4)  QR(10, 11, 8, 9) QR(15, 12, 13, 14)
    }
#undef QR
    for (int i = 0; i < 16; i++) {
        u32 v = x[i] + j[i];
        out[4 * i] = (unsigned char)v; out[4 * i + 1] = (unsigned char)(v >> 8);
        out[4 * i + 2] = (unsigned char)(v >> 16); out[4 * i + 3] = (unsigned char)(v >> 24);
    }
}

/* Keystream of floor(nbytes/64) blocks starting at block counter 0 (distributions.cuh:224, 253: NBLKS = n/64). */
ORC_API void orc_salsa20_keystream(unsigned char *out, size_t nbytes, const unsigned char key[32], u64 nonce)
{
    size_t nblk = nbytes / 64;
    for (size_t b = 0; b < nblk; b++) salsa20_block(out + 64 * b, key, nonce, b);
}
/* distributions.cuh:249-276 generate_random_default: key = 32 x 0x01, nonce 0.  (orc_set_nonce lets the tests restate
 * the batched pipelines, where item k draws from nonce k; the reference itself always uses 0.) */
static u64 g_nonce = 0;
ORC_API void orc_set_nonce(u64 nonce) { g_nonce = nonce; }
ORC_API void orc_generate_random_default(unsigned char *out, unsigned nbytes)
{
    unsigned char key[32]; memset(key, 1, 32);
    orc_salsa20_keystream(out, nbytes, key, g_nonce);
}
/* distributions.cuh:220-247 generate_random: key = 0x4D but only the first 24 bytes are uploaded (:235);
 * bytes 24..31 keep whatever the `key` symbol held before (`prev_key_tail`: zeros on a fresh context,
 * 0x01 after a generate_random_default call). */
ORC_API void orc_generate_random(unsigned char *out, unsigned nbytes, const unsigned char prev_key_tail[8])
{
    unsigned char key[32]; memset(key, 77, 24); memcpy(key + 24, prev_key_tail, 8);
    orc_salsa20_keystream(out, nbytes, key, 0);
}

/* bfv_keygen.cuh:14-31 ternary formula (also bfv_encryption.cuh:23-36): int(float(b)/(255.0f/3)) - 1 */
static inline u64 ternary_value(unsigned char byte, u64 q)
{
    float d = (float)byte;
    d /= (255.0f / 3);
    int b = (int)d - 1;
    return (u64)(b < 0) * q + (u64)(long long)b;
}
/* distributions.cuh:204-218 convert_ternary (legacy converter: 256.0f/3 thresholds) */
ORC_API void orc_convert_ternary(const unsigned char *in, u64 *out, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) {
        float d = (float)in[i];
        d /= (256.0f / 3);
        if (d >= 2) out[i] = 1; else if (d >= 1) out[i] = 0; else out[i] = q - 1;
    }
}
/* distributions.cuh:191-202 convert_range / bfv_keygen.cuh:33-45 uniform_dist_xq */
static inline u64 uniform_value(u64 x, u64 q)
{
    double d = (double)x;
    d /= 18446744073709551615ULL;   /* UINT64_MAX -> (double) 2^64 */
    d *= (double)(q - 1);
    return (u64)d;
}
ORC_API void orc_convert_range(const u64 *in, u64 *out, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) out[i] = uniform_value(in[i], q);
}

/* Inverse normal CDF in double precision (Acklam's rational approximation + one Halley step with
 * erfc), rounded to float.  NOT bit-identical to CUDA's normcdfinvf: see the header comment. */
static double norm_cdf_inv(double p)
{
    static const double a[] = { -3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                                1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00 };
    static const double b[] = { -5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                                6.680131188771972e+01, -1.328068155288572e+01 };
    static const double c[] = { -7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                                -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00 };
    static const double d[] = { 7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                                3.754408661907416e+00 };
    double x, qq, r;
    if (p < 0.02425) {
        qq = sqrt(-2 * log(p));
        x = (((((c[0] * qq + c[1]) * qq + c[2]) * qq + c[3]) * qq + c[4]) * qq + c[5]) /
            ((((d[0] * qq + d[1]) * qq + d[2]) * qq + d[3]) * qq + 1);
    } else if (p <= 1 - 0.02425) {
        qq = p - 0.5; r = qq * qq;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * qq /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1);
    } else {
        qq = sqrt(-2 * log(1 - p));
        x = -(((((c[0] * qq + c[1]) * qq + c[2]) * qq + c[3]) * qq + c[4]) * qq + c[5]) /
            ((((d[0] * qq + d[1]) * qq + d[2]) * qq + d[3]) * qq + 1);
    }
    for (int it = 0; it < 2; it++) {
        double e = 0.5 * erfc(-x / sqrt(2.0)) - p;
        double u = e * sqrt(2 * M_PI) * exp(x * x / 2);
        x = x - u / (1 + x * u / 2);
    }
    return x;
}
/* bfv_keygen.cuh:47-79 gaussian_dist_xq / distributions.cuh:157-189 / bfv_encryption.cuh:49-76:
 * returns the signed sample dd (the caller maps negative to q + dd). */
static inline int gaussian_value(u32 x)
{
    float d = (float)x;
    d /= 4294967295;                 /* long literal -> float 2^32 */
    if (d == 0) d += 1.192092896e-07F;
    else if (d == 1) d -= 1.192092896e-07F;
    d = (float)norm_cdf_inv((double)d);
    d = d * (float)3.2 + 0;
    if (d > 19.2) d = 19.2; else if (d < -19.2) d = -19.2;
    return (int)d;
}
ORC_API void orc_gaussian_samples(const u32 *in, int *out, unsigned n)
{
    for (unsigned i = 0; i < n; i++) out[i] = gaussian_value(in[i]);
}
ORC_API void orc_convert_gaussian(const u32 *in, u64 *out, unsigned n, u64 q)
{
    for (unsigned i = 0; i < n; i++) { int dd = gaussian_value(in[i]); out[i] = dd < 0 ? q + (u64)(long long)dd : (u64)dd; }
}

/* bfv_keygen.cuh:14-31 ternary_dist_xq: same byte i % n for every limb */
ORC_API void orc_ternary_dist_xq(const unsigned char *in, u64 *sk, unsigned n, unsigned q_amount, const u64 *q)
{
    for (size_t i = 0; i < (size_t)n * q_amount; i++) sk[i] = ternary_value(in[i % n], q[i / n]);
}
/* bfv_keygen.cuh:33-45 uniform_dist_xq */
ORC_API void orc_uniform_dist_xq(const unsigned char *in, u64 *pk, unsigned n, unsigned q_amount, const u64 *q)
{
    for (size_t i = 0; i < (size_t)n * q_amount; i++) { u64 x; memcpy(&x, in + 8 * i, 8); pk[i] = uniform_value(x, q[i / n]); }
}
/* bfv_keygen.cuh:47-79 gaussian_dist_xq; `samples` (n signed values) overrides the CPU normcdfinv when non-NULL */
ORC_API void orc_gaussian_dist_xq(const unsigned char *in, u64 *temp, unsigned n, unsigned q_amount, const u64 *q, const int *samples)
{
    for (size_t i = 0; i < (size_t)n * q_amount; i++) {
        int dd;
        if (samples) dd = samples[i % n];
        else { u32 x; memcpy(&x, in + 4 * (i % n), 4); dd = gaussian_value(x); }
        temp[i] = dd < 0 ? q[i / n] + (u64)(long long)dd : (u64)dd;
    }
}
/* bfv_keygen.cuh:81-93 poly_add_negate_xq */
ORC_API void orc_poly_add_negate_xq(u64 *a, const u64 *b, unsigned n, unsigned q_amount, const u64 *q)
{
    for (size_t i = 0; i < (size_t)n * q_amount; i++) {
        u64 qi = q[i / n];
        u64 r = a[i] + b[i];
        if (r >= qi) r -= qi;
        r = qi - r;
        a[i] = r * (r != qi);
    }
}

/* ------------------------------------------------------------------------------------------------
 * L3: BFV pipelines
 * ---------------------------------------------------------------------------------------------- */

typedef struct {
    unsigned n, r;              /* ring degree, number of limbs (all limbs, before q_amount--) */
    const u64 *q, *mu;          /* [r] */
    const unsigned *qbit;       /* [r] */
    const u64 *psi, *psiinv;    /* [r][n] */
} orc_ring;

/* bfv_keygen.cuh:95-151 keygen_rns.  in: scratch of 9*r*n + 4*n bytes (filled here);
 * secret_key[r][n], public_key[2][r][n], temp[r][n].  e_samples: optional n signed gaussian draws. */
ORC_API void orc_keygen_rns(unsigned char *in, unsigned r, unsigned n, const u64 *q, const u64 *mu, const unsigned *qbit,
                            const u64 *psi, const u64 *psiinv, u64 *secret_key, u64 *public_key, u64 *temp, const int *e_samples)
{
    size_t rn = (size_t)r * n;
    orc_generate_random_default(in, (unsigned)((1 + 8) * rn + 4 * (size_t)n));          /* :99  */
    orc_ternary_dist_xq(in, secret_key, n, r, q);                                        /* :120 */
    orc_uniform_dist_xq(in + n, public_key + rn, n, r, q);                               /* :121 */
    orc_gaussian_dist_xq(in + n + 8 * rn, temp, n, r, q, e_samples);                     /* :122 */
    orc_forward_ntt_batch(secret_key, n, psi, r, r, q, mu, qbit);                        /* :129 */
    orc_barrett_batch_3param(public_key, public_key + rn, secret_key, n, r, r, q, mu, qbit); /* :132 */
    orc_inverse_ntt_batch(public_key, n, psiinv, r, r, q, mu, qbit);                     /* :133 */
    orc_poly_add_negate_xq(public_key, temp, n, r, q);                                   /* :144 */
    orc_forward_ntt_batch(public_key, n, psi, r, r, q, mu, qbit);                        /* :145 */
}

/* bfv_encryption.cuh:17-109 convert_ternary_gaussian_x2 */
static void convert_ternary_gaussian_x2(const unsigned char *in, u64 *c, u64 *e, unsigned n, unsigned r, const u64 *q,
                                        const int *e0_samples, const int *e1_samples)
{
    size_t rn = (size_t)r * n;
    for (size_t i = 0; i < rn; i++) {
        u64 qi = q[i / n];
        u64 tv = ternary_value(in[i % n], qi);
        c[i] = tv; c[i + rn] = tv;
        int d0, d1;
        if (e0_samples) d0 = e0_samples[i % n];
        else { u32 x; memcpy(&x, in + n + 4 * (i % n), 4); d0 = gaussian_value(x); }
        if (e1_samples) d1 = e1_samples[i % n];
        else { u32 x; memcpy(&x, in + 5 * (size_t)n + 4 * (i % n), 4); d1 = gaussian_value(x); }
        e[i] = d0 < 0 ? qi + (u64)(long long)d0 : (u64)d0;
        e[i + rn] = d1 < 0 ? qi + (u64)(long long)d1 : (u64)d1;
    }
}

/* bfv_encryption.cuh:223-290 encryption_rns.  c[2][r][n], e[2][r][n], in: 9n bytes scratch, m_poly[n],
 * inv_q_last_mod_q[r-1], qi_div_t[r].  After return limb r-1 of each half of c is padding. */
ORC_API void orc_encryption_rns(u64 *c, const u64 *public_key, unsigned char *in, u64 *e, unsigned n, unsigned r,
                                const u64 *q, const u64 *mu, const unsigned *qbit, const u64 *inv_q_last_mod_q,
                                const u64 *psi, const u64 *psiinv, const u64 *m_poly, const u64 *qi_div_t, u64 t,
                                const int *e0_samples, const int *e1_samples)
{
    size_t rn = (size_t)r * n;
    orc_generate_random_default(in, 9 * n);                                              /* :228 */
    convert_ternary_gaussian_x2(in, c, e, n, r, q, e0_samples, e1_samples);              /* :247 */
    orc_forward_ntt_batch(c, n, psi, 2 * r, r, q, mu, qbit);                             /* :268 */
    orc_barrett_batch(c, public_key, n, 2 * r, r, q, mu, qbit);                          /* :270 */
    orc_inverse_ntt_batch(c, n, psiinv, 2 * r, r, q, mu, qbit);                          /* :271 */
    /* poly_add_xq :180-191 (quirk: > instead of >=) */
    for (unsigned h = 0; h < 2; h++)
        for (size_t i = 0; i < rn; i++) {
            u64 ra = c[i + rn * h] + e[i + rn * h];
            if (ra > q[i / n]) ra -= q[i / n];
            c[i + rn * h] = ra;
        }
    /* divide_and_round_q_last_inplace_add_x2 :111-125 */
    u64 last = q[r - 1], half = last >> 1;
    for (unsigned i = 0; i < 2 * n; i++) {
        size_t k = (size_t)n * (r - 1) + i % n + rn * (i >= n);
        u64 ra = c[k] + half;
        if (ra >= last) ra -= last;
        c[k] = ra;
    }
    /* divide_and_round_q_last_inplace_loop_xq :127-178 */
    for (size_t i = 0; i < (size_t)2 * n * (r - 1); i++) {
        size_t ii = i % n;
        unsigned index = (unsigned)((i % ((size_t)n * (r - 1))) / n);
        u64 qi = q[index];
        u64 half_mod = half % qi;
        unsigned second = i >= (size_t)n * (r - 1);
        u64 *lastp = c + second * rn + (size_t)n * (r - 1);
        u64 *inp = c + second * rn + (size_t)n * index;
        u64 tp = lastp[ii] % qi;
        if (tp < half_mod) tp += qi;
        tp -= half_mod;
        u64 x = inp[ii];
        if (x < tp) x += qi;
        x -= tp;
        inp[ii] = ref_barrett((u128)x * inv_q_last_mod_q[index], qi, mu[index], (int)qbit[index]);
    }
    /* weird_m_stuff :193-212 */
    for (unsigned j = 0; j < n; j++) {
        u64 numerator = m_poly[j] + ((t + 1) >> 1);
        u64 fix = numerator / t;
        for (unsigned i = 0; i + 1 < r; i++)
            c[j + (size_t)i * n] = (c[j + (size_t)i * n] + ((m_poly[j] * qi_div_t[i]) + fix)) % q[i];
    }
}

/* bfv_decryption.cuh:76-138 decryption_rns.  rp = r - 1 limbs (the driver's q_amount after q_amount--);
 * c[2][rp+1][n]; secret_key[>=rp][n] (NTT domain); bcm[2][rp]; plaintext lands at c[(rp-1)*n .. rp*n). */
ORC_API void orc_decryption_rns(u64 *c, const u64 *secret_key, unsigned n, unsigned rp, const u64 *q, const u64 *mu,
                                const unsigned *qbit, const u64 *psi, const u64 *psiinv, const u64 *inv_punctured_q,
                                const u64 *prod_t_gamma_mod_q, const u64 *bcm, u64 t, u64 gamma, u64 mu_gamma,
                                int gamma_bits, const u64 *neg_inv_q_mod_t_gamma, u64 gamma_div_2)
{
    u64 *c1 = c + (size_t)(rp + 1) * n;
    orc_forward_ntt_batch(c1, n, psi, rp, rp + 1, q, mu, qbit);                           /* :98  */
    orc_barrett_batch(c1, secret_key, n, rp, rp, q, mu, qbit);                            /* :100 */
    orc_inverse_ntt_batch(c1, n, psiinv, rp, rp + 1, q, mu, qbit);                        /* :101 */
    for (size_t i = 0; i < (size_t)n * rp; i++) {                                         /* poly_add_xq_d :13-23 */
        u64 ra = c1[i] + c[i];
        if (ra > q[i / n]) ra -= q[i / n];
        c1[i] = ra;
    }
    for (size_t i = 0; i < (size_t)n * rp; i++)                                           /* :25-40 */
        c1[i] = ref_barrett((u128)c1[i] * prod_t_gamma_mod_q[i / n], q[i / n], mu[i / n], (int)qbit[i / n]);
    for (size_t i = 0; i < (size_t)n * rp; i++)                                           /* :42-57 */
        c1[i] = ref_barrett((u128)c1[i] * inv_punctured_q[i / n], q[i / n], mu[i / n], (int)qbit[i / n]);
    orc_fast_convert_t(c1, c, t, bcm, rp, n);                                             /* poly_arithmetic.cuh:272 */
    orc_fast_convert_gamma(c1, c, gamma, bcm, rp, gamma_bits, mu_gamma, n);               /* :274 */
    orc_mod_t(c, neg_inv_q_mod_t_gamma[0], n, t);                                         /* :133 */
    orc_barrett_int(c + n, neg_inv_q_mod_t_gamma[1], n, gamma, mu_gamma, gamma_bits);     /* :134 */
    orc_dec_round(c, c + (size_t)n * (rp - 1), t, gamma, gamma_div_2, n);                 /* :137 */
}

/* ------------------------------------------------------------------------------------------------
 * Derived RNS constants built by the driver (demo.cu:62-264); r = all limbs, rp = r-1.
 * Outputs: qbit[r], mu[r], inv_q_last_mod_q[rp], qi_div_t[r], psiinv_root[r], neg_inv[2],
 *          prod_t_gamma_mod_q[rp], mu_gamma, inv_punctured_q[rp], bcm[2*rp].
 * ---------------------------------------------------------------------------------------------- */
ORC_API void orc_derive_params(unsigned r, const u64 *q, const u64 *psi_root, u64 t, u64 gamma, int gamma_bits,
                               unsigned *qbit, u64 *mu, u64 *inv_q_last_mod_q, u64 *qi_div_t, u64 *psiinv_root,
                               u64 *neg_inv, u64 *prod_t_gamma_mod_q, u64 *mu_gamma, u64 *inv_punctured_q, u64 *bcm)
{
    unsigned rp = r - 1;
    u64 base[2] = { t, gamma };
    for (unsigned i = 0; i < r; i++) {
        qbit[i] = orc_qbit(q[i]);                                                          /* :69      */
        mu[i] = orc_mu(q[i], qbit[i]);                                                     /* :157-165 */
        qi_div_t[i] = q[i] / t;                                                            /* :84-88   */
        psiinv_root[i] = orc_modinv(psi_root[i], q[i]);                                    /* :96-97   */
    }
    for (unsigned i = 0; i < rp; i++) inv_q_last_mod_q[i] = orc_modinv(q[r - 1] % q[i], q[i]);  /* :75-79 */
    u64 mult_t = 1, mult_g = 1;
    for (unsigned i = 0; i < rp; i++) {                                                    /* :103-108 */
        mult_t = ref_mod128((u128)mult_t * q[i], t);
        mult_g = ref_mod128((u128)mult_g * q[i], gamma);
    }
    neg_inv[0] = t - orc_modinv(mult_t, t);                                                /* :109 */
    neg_inv[1] = gamma - orc_modinv(mult_g, gamma);                                        /* :110 */
    for (unsigned i = 0; i < rp; i++) prod_t_gamma_mod_q[i] = ref_mod128((u128)t * gamma, q[i]); /* :118-123 */
    *mu_gamma = (u64)(((u128)1 << (2 * gamma_bits)) / gamma);                              /* :221-226 */
    for (unsigned i = 0; i < rp; i++) {                                                    /* :229-243 */
        u64 tmp = 1;
        for (unsigned j = 0; j < rp; j++) if (j != i) tmp = ref_mod128((u128)tmp * q[j], q[i]);
        inv_punctured_q[i] = orc_modinv(tmp, q[i]);
    }
    for (unsigned k = 0; k < 2; k++)                                                       /* :248-264 */
        for (unsigned j = 0; j < rp; j++) {
            u64 tmp = 1;
            for (unsigned i = 0; i < rp; i++) if (i != j) tmp = ref_mod128((u128)tmp * q[i], base[k]);
            bcm[k * rp + j] = tmp;
        }
}

/* helper.h:95-126 refPolyMul128 (schoolbook negacyclic product; d must hold n values) */
ORC_API void orc_ref_poly_mul(const u64 *a, const u64 *b, u64 m, int n, u64 *d)
{
    u64 *c = (u64 *)calloc((size_t)2 * n, sizeof(u64));
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            c[i + j] = (u64)((u128)a[i] * b[j] % m) + c[i + j] % m;
            c[i + j] %= m;
        }
    for (int i = 0; i < n; i++) {
        if (c[i] < c[i + n]) c[i] += m;
        d[i] = (c[i] - c[i + n]) % m;
    }
    free(c);
}

/* ------------------------------------------------------------------------------------------------
 * Deterministic test inputs: splitmix64 with rejection, portable to Python.
 * ---------------------------------------------------------------------------------------------- */
static inline u64 splitmix64(u64 *s)
{
    u64 z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
ORC_API void orc_fill_uniform(u64 *a, size_t n, u64 q, u64 seed)
{
    int bits = 64; while (bits > 1 && !((q - 1) >> (bits - 1))) bits--;
    u64 mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
    u64 s = seed;
    for (size_t i = 0; i < n; i++) { u64 v; do v = splitmix64(&s) & mask; while (v >= q); a[i] = v; }
}

/* Multi-threaded CPU baseline helper: transforms polys [p0, p1) (called from one thread each). */
ORC_API void orc_forward_ntt_fast_range(u64 *a, unsigned n, const u64 *psi, unsigned division, const u64 *q,
                                        unsigned p0, unsigned p1)
{
    for (unsigned p = p0; p < p1; p++) { unsigned l = p % division; orc_forward_ntt_fast(a + (size_t)p * n, n, q[l], psi + (size_t)l * n); }
}
ORC_API void orc_inverse_ntt_fast_range(u64 *a, unsigned n, const u64 *psiinv, unsigned division, const u64 *q,
                                        unsigned p0, unsigned p1)
{
    for (unsigned p = p0; p < p1; p++) { unsigned l = p % division; orc_inverse_ntt_fast(a + (size_t)p * n, n, q[l], psiinv + (size_t)l * n); }
}
