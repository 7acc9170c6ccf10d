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

#define NTTB200_MAX_LIMBS_INTERNAL 64

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
// blk0: first block of every stream that is generated (out + s * stream_stride receives block blk0 first) -- the sharded pipelines
// draw only the part of an item's keystream they consume (Salsa20 is counter-addressable).
NTT_KERNEL void k_salsa20_keystream(unsigned char *out, u64 blocks_per_stream, u64 streams, size_t stream_stride, SalsaKey key, u64 nonce0,
                                    u64 blk0 = 0)
{
    NTT_PDL_ENTER();
    NTT_GRID_STRIDE(i, blocks_per_stream * streams) {
        const u64 s = i / blocks_per_stream, b = i - s * blocks_per_stream;
        u32 o[16];
        salsa20_block(o, key, nonce0 + s, blk0 + b);
        uint4 *dst = reinterpret_cast<uint4 *>(out + s * stream_stride + b * 64);
        NTT_UNROLL
        for (int v = 0; v < 4; v++) { uint4 t; t.x = o[4 * v]; t.y = o[4 * v + 1]; t.z = o[4 * v + 2]; t.w = o[4 * v + 3]; dst[v] = t; }
    }
}

// ---- distribution converters: the reference's formulas -----------------------------------------------------------------
// ternary_value(byte, q): modarith.cuh (shared with the NTT pass that generates u on the fly)
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

// Launch geometry of the fused kernels: blockIdx.x strides over coefficient PAIRS (16-byte accesses), blockIdx.y / .z carry
// the limb / half / item, so no thread ever divides by n or r.  Per-limb constants (and the few 64-bit divisions they
// need) are computed once per CTA into shared memory.
#define NTT_PAIR_STRIDE(j, n) \
    for (u32 j = (blockIdx.x * blockDim.x + threadIdx.x) * 2u; j < (n); j += gridDim.x * blockDim.x * 2u)

constexpr int kMaxLimbs = NTTB200_MAX_LIMBS_INTERNAL;

// exact x mod q for any 64-bit x, ratio = floor(2^64 / q): x - floor(x * ratio / 2^64) * q is in [0, 2q).  The remainder
// is unique, so this is bit-identical to the reference's `%` (bfv_encryption.cuh:146,209; poly_arithmetic.cuh:248).
__host__ __device__ __forceinline__ u64 mod_exact(u64 x, u64 q, u64 ratio) { return csub(x - mulhi64(x, ratio) * q, q); }
__host__ __device__ __forceinline__ u64 ratio_of(u64 q) { return ~0ull / q; }   // = floor(2^64 / q) for odd q > 1

// The reference's Barrett (one correction) is exact for every product a*b < 2^(2*qbit) iff its quotient estimate is never 2
// short; that is guaranteed when delta = frac(2^(2*qbit) / q) < 1/2 (estimate error < delta + 1/2 + frac).
// For such moduli any exact modular multiplication returns the reference's bits, so multiplications by per-limb CONSTANTS
// use a Shoup product (mul.hi + 2 mul.lo) instead of the 30-instruction Barrett sequence.  Moduli with delta >= 1/2
// (68719230977 and 274877202433 among the reference's) keep the literal sequence so even their glitches reproduce.
__host__ __device__ __forceinline__ bool barrett_is_exact(u64 q, u64 mu, int qbit)
{
    const u64 rem = (2 * qbit >= 64 ? 0ull : (1ull << (2 * qbit))) - mu * q;     // 2^(2 qbit) mod q (mod 2^64 arithmetic, rem < q)
    return qbit >= 3 && qbit <= 61 && rem < q && rem * 2 < q;
}
__host__ __device__ __forceinline__ u64 shoup_companion(u64 c, u64 q) { return (u64)((((unsigned __int128)c) << 64) / q); }
// v * c mod q, canonical, for a constant c < q with companion cs; `fast` as decided by barrett_is_exact
__host__ __device__ __forceinline__ u64 mul_const(u64 v, u64 c, u64 cs, bool fast, u64 q, u64 mu, int qbit)
{
    return fast ? csub(shoup_mul(v, c, cs, q), q) : barrett_ref(v, c, q, mu, qbit);
}

__device__ __forceinline__ ulonglong2 ld2(const u64 *p) { return *reinterpret_cast<const ulonglong2 *>(p); }
__device__ __forceinline__ void st2(u64 *p, u64 a, u64 b) { *reinterpret_cast<ulonglong2 *>(p) = make_ulonglong2(a, b); }

// keygen sampling: sk = ternary (all limbs), pk1 = uniform, es[k][j] = gaussian draw (signed).  grid (x, batch)
// sk == nullptr: the ternary secret is generated inside the first strided NTT pass instead (NttArgs::gen_src)
NTT_KERNEL void k_keygen_sample(const unsigned char *in, size_t in_stride, u64 *sk, u64 *pk, int *es, unsigned n, unsigned r,
                                unsigned batch, const u64 *q)
{
    NTT_PDL_ENTER();
    (void)batch;
    const size_t rn = (size_t)r * n, k = blockIdx.y;
    const unsigned char *s = in + k * in_stride;
    NTT_PAIR_STRIDE(j, n) {
        const unsigned char b0 = s[j], b1 = s[j + 1];
        const u32 *g = reinterpret_cast<const u32 *>(s + n + 8 * rn) + j;
        es[k * n + j] = gaussian_value(g[0]);
        es[k * n + j + 1] = gaussian_value(g[1]);
        for (unsigned l = 0; l < r; l++) {
            const u64 ql = q[l];
            if (sk) st2(sk + k * rn + (size_t)l * n + j, ternary_value(b0, ql), ternary_value(b1, ql));
            const ulonglong2 u = ld2(reinterpret_cast<const u64 *>(s + n) + (size_t)l * n + j);
            st2(pk + k * 2 * rn + rn + (size_t)l * n + j, uniform_value(u.x, ql), uniform_value(u.y, ql));
        }
    }
}
// pk0 = pk1 (.) sk   (barrett_batch_3param, bfv_keygen.cuh:132).  grid (x, r, batch)
NTT_KERNEL void k_keygen_mul(u64 *pk, const u64 *sk, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    NTT_PDL_ENTER();
    (void)batch;
    const unsigned l = blockIdx.y;
    const size_t rn = (size_t)r * n, k = blockIdx.z;
    const u64 q = L.q[l], mu = L.mu[l];
    const int qb = (int)L.qbit[l];
    u64 *p0 = pk + k * 2 * rn + (size_t)l * n;
    const u64 *p1 = p0 + rn, *s = sk + k * rn + (size_t)l * n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 a = ld2(p1 + j), b = ld2(s + j);
        st2(p0 + j, barrett_ref(a.x, b.x, q, mu, qb), barrett_ref(a.y, b.y, q, mu, qb));
    }
}
// pk0 = -(pk0 + e)   (gaussian_dist_xq + poly_add_negate_xq, bfv_keygen.cuh:47-93).  grid (x, r, batch)
NTT_KERNEL void k_keygen_add_negate(u64 *pk, const int *es, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    NTT_PDL_ENTER();
    (void)batch;
    const unsigned l = blockIdx.y;
    const size_t rn = (size_t)r * n, k = blockIdx.z;
    const u64 q = L.q[l];
    u64 *p0 = pk + k * 2 * rn + (size_t)l * n;
    const int *e = es + k * n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 a = ld2(p0 + j);
        u64 r0 = a.x + signed_to_residue(e[j], q), r1 = a.y + signed_to_residue(e[j + 1], q);
        if (r0 >= q) r0 -= q;
        if (r1 >= q) r1 -= q;
        r0 = q - r0; r1 = q - r1;
        st2(p0 + j, r0 * (u64)(r0 != q), r1 * (u64)(r1 != q));
    }
}

// encryption sampling: u (ternary) into half 0 of c for every limb; es[k][0][j], es[k][1][j] = the two gaussian draws
// grid (x, batch)
NTT_KERNEL void k_encrypt_sample(const unsigned char *in, size_t in_stride, u64 *c, int *es, unsigned n, unsigned r, unsigned batch,
                                 const u64 *q)
{
    NTT_PDL_ENTER();
    (void)batch;
    const size_t rn = (size_t)r * n, k = blockIdx.y;
    const unsigned char *s = in + k * in_stride;
    NTT_PAIR_STRIDE(j, n) {
        const unsigned char b0 = s[j], b1 = s[j + 1];
        const u32 *g0 = reinterpret_cast<const u32 *>(s + n) + j, *g1 = reinterpret_cast<const u32 *>(s + (size_t)n * 5) + j;
        int *e = es + k * 2 * n;
        e[j] = gaussian_value(g0[0]); e[j + 1] = gaussian_value(g0[1]);
        e[n + j] = gaussian_value(g1[0]); e[n + j + 1] = gaussian_value(g1[1]);
        for (unsigned l = 0; l < r; l++) {
            const u64 ql = q[l];
            st2(c + k * 2 * rn + (size_t)l * n + j, ternary_value(b0, ql), ternary_value(b1, ql));
        }
    }
}
// the two gaussian draws only: u is generated inside the first strided NTT pass straight from the keystream (NttArgs::gen_src),
// so the r*n ternary residues are never written to and re-read from HBM.  grid (x, batch)
// e_off: byte offset of the e0 words inside an item's stream (n for the full 9n-byte stream; 0 when only blocks [n/64, 9n/64) were drawn)
NTT_KERNEL void k_encrypt_gauss(const unsigned char *in, size_t in_stride, int *es, unsigned n, unsigned batch, size_t e_off)
{
    NTT_PDL_ENTER();
    (void)batch;
    const size_t k = blockIdx.y;
    const unsigned char *s = in + k * in_stride;
    NTT_PAIR_STRIDE(j, n) {
        const u32 *g0 = reinterpret_cast<const u32 *>(s + e_off) + j, *g1 = reinterpret_cast<const u32 *>(s + e_off + (size_t)n * 4) + j;
        int *e = es + k * 2 * n;
        e[j] = gaussian_value(g0[0]); e[j + 1] = gaussian_value(g0[1]);
        e[n + j] = gaussian_value(g1[0]); e[n + j + 1] = gaussian_value(g1[1]);
    }
}
// Sampling of the fused-epilogue encryption path: one launch, one thread per 64-byte Salsa20 block of an item's keystream
// (bfv_encryption.cuh:228: 9n bytes).  Blocks [0, n/64) are the ternary source of u and are stored as they are (n bytes per
// item, read by the first strided NTT pass); blocks [n/64, 9n/64) are 2n 32-bit words that become the gaussian draws e0 | e1,
// stored as signed bytes (|e| <= 19).  The 8n keystream bytes of e and the 8n bytes of 32-bit draws never reach HBM.
// item0 / items: the items of the batch that get their e part (sharded runs draw e only for the items they finish);
// every item gets its u part when `ub` is given.
NTT_KERNEL void k_encrypt_sample_fused(unsigned char *ub, signed char *es8, unsigned n, u64 items, SalsaKey key, u64 nonce0, int want_u, int want_e)
{
    NTT_PDL_ENTER();
    const u64 ublk = n / 64, eblk = 8 * (u64)n / 64;
    const u64 per = (want_u ? ublk : 0) + (want_e ? eblk : 0);
    NTT_GRID_STRIDE(i, per * items) {
        const u64 k = i / per;
        u64 b = i - k * per;
        if (!want_u) b += ublk;
        u32 o[16];
        salsa20_block(o, key, nonce0 + k, b);
        if (b < ublk) {
            uint4 *dst = reinterpret_cast<uint4 *>(ub + k * n + b * 64);
            NTT_UNROLL
            for (int v = 0; v < 4; v++) { uint4 t; t.x = o[4 * v]; t.y = o[4 * v + 1]; t.z = o[4 * v + 2]; t.w = o[4 * v + 3]; dst[v] = t; }
        } else {
            u32 w[4];
            NTT_UNROLL
            for (int v = 0; v < 4; v++) {
                w[v] = 0;
                NTT_UNROLL
                for (int e = 0; e < 4; e++) w[v] |= ((u32)(gaussian_value(o[4 * v + e]) & 0xff)) << (8 * e);
            }
            uint4 t; t.x = w[0]; t.y = w[1]; t.z = w[2]; t.w = w[3];
            *reinterpret_cast<uint4 *>(es8 + k * 2 * (u64)n + (b - ublk) * 16) = t;
        }
    }
}
// c0 = NTT(u) (.) pk0, c1 = NTT(u) (.) pk1 -- the reference transforms u twice; here NTT(u) sits in
// half 0 and is read once.  pk_stride = 0: one public key for the whole batch.  grid (x, r, batch)
NTT_KERNEL void k_encrypt_mul(u64 *c, const u64 *pk, size_t pk_stride, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    NTT_PDL_ENTER();
    (void)batch;
    const unsigned l = blockIdx.y;
    const size_t rn = (size_t)r * n, k = blockIdx.z;
    const u64 q = L.q[l], mu = L.mu[l];
    const int qb = (int)L.qbit[l];
    u64 *c0 = c + k * 2 * rn + (size_t)l * n, *c1 = c0 + rn;
    const u64 *p0 = pk + k * pk_stride + (size_t)l * n, *p1 = p0 + rn;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 uh = ld2(c0 + j), a = ld2(p0 + j), b = ld2(p1 + j);
        st2(c0 + j, barrett_ref(uh.x, a.x, q, mu, qb), barrett_ref(uh.y, a.y, q, mu, qb));
        st2(c1 + j, barrett_ref(uh.x, b.x, q, mu, qb), barrett_ref(uh.y, b.y, q, mu, qb));
    }
}

struct EncLimb { u64 q, mu, ratio, half_mod, inv_q_last, inv_q_last_s, qdt, bias; int qbit, fast, last_lt_2q, lazy; };
constexpr unsigned kEncChunk = 4;      // limbs per CTA row of k_encrypt_epilogue
// last limb of one half: += e (`>` quirk, bfv_encryption.cuh:187), += floor(q_last / 2) mod q_last (:121-124)
__host__ __device__ __forceinline__ u64 enc_last_limb(u64 v, int d, u64 last, u64 half_last)
{
    u64 x = v + signed_to_residue(d, last);
    if (x > last) x -= last;
    x += half_last;
    if (x >= last) x -= last;
    return x;
}
// poly_add_xq + divide_and_round_q_last_inplace_add_x2 + ..._loop_xq + weird_m_stuff (bfv_encryption.cuh:111-212) for the limbs
// below the dropped one, in one pass.  grid (x, 2 halves * ceil((r-1) / kEncChunk) limb chunks, batch): every CTA row handles
// kEncChunk limbs of one half, so (r-1)/4 times more 16-byte loads are in flight than with one thread walking all limbs (the
// single-row version was latency-bound at 3 TB/s).  Every row recomputes the dropped limb's value from the RAW INTT output,
// which is therefore left untouched here and finalised by k_encrypt_last_limb afterwards.
// ALL_LAZY (decided on the host for context-owned parameter sets): every limb takes the lazy path, the literal reference
// sequence is compiled out and the kernel fits four 256-thread CTAs per SM -- the pass is bound by loads in flight.
// Generalised for limb windows (sharded runs): the launch covers `cnt` limbs, local limb l = global limb first + l, stored at
// c[item][half][l][n] with the given strides; the dropped limb's values come from clp[item][half][n] -- RAW inverse-transform outputs
// (CL_FINISHED = false: `+ e` and the rounding offset are applied here, as the single-GPU path does on the padding slot) or the
// finished values an earlier launch left there (CL_FINISHED = true).  ES: int (reference-style scratch) or signed char draws.
template <bool ALL_LAZY, class ES, bool CL_FINISHED>
NTT_KERNEL void __launch_bounds__(256, ALL_LAZY ? 4 : 2)
k_encrypt_epilogue(u64 *c, size_t item_stride, size_t half_stride, const ES *es, const u64 *m_poly, size_t m_stride, unsigned n, unsigned r,
                   unsigned first, unsigned cnt, const u64 *clp, size_t cl_item_stride, size_t cl_half_stride, u64 t, const u64 *qi_div_t,
                   LimbArrays L)
{
    NTT_PDL_ENTER();
    NTT_SHARED EncLimb K[kEncChunk];
    const u64 last = L.q[r - 1], half_last = last >> 1;
    const unsigned chunks = (cnt + kEncChunk - 1) / kEncChunk;
    const unsigned h = blockIdx.y / chunks, l0 = (blockIdx.y % chunks) * kEncChunk;
    if (threadIdx.x < kEncChunk && l0 + threadIdx.x < cnt) {
        const unsigned l = first + l0 + threadIdx.x;
        EncLimb e;
        e.q = L.q[l]; e.mu = L.mu[l]; e.qbit = (int)L.qbit[l];
        e.ratio = ratio_of(e.q);
        e.half_mod = half_last % e.q;
        e.inv_q_last = L.inv_q_last_mod_q[l];
        e.fast = barrett_is_exact(e.q, e.mu, e.qbit) && e.inv_q_last < e.q;
        e.inv_q_last_s = e.fast ? shoup_companion(e.inv_q_last, e.q) : 0;
        e.qdt = qi_div_t[l];
        e.last_lt_2q = last <= 2 * e.q;      // c_last < q_last <= 2 q_i: its residue mod q_i is one conditional subtraction
        // lazy path: every intermediate of the reference is only ever consumed modulo q_i and the value stored is canonical, so
        // c_i + e - c_last + half_last may be formed as ONE biased sum in [0, 5q) and reduced by the (exact) Shoup product
        e.lazy = e.fast && e.last_lt_2q && e.q < (1ull << 60);
        e.bias = e.half_mod + 3 * e.q;
        K[threadIdx.x] = e;
    }
    __syncthreads();
    const size_t k = blockIdx.z;
    u64 *ch = c + k * item_stride + (size_t)h * half_stride;
    const u64 *clh = clp + k * cl_item_stride + (size_t)h * cl_half_stride;
    const ES *e = es + k * 2 * n + (size_t)h * n;
    const u64 *mp = m_poly + k * m_stride;
    const u64 tfix = (t + 1) >> 1;
    // t is a power of two in every reference parameter set; keep the general division for anything else
    const bool tpow2 = (t & (t - 1)) == 0;
    int tsh = 0;
    while (tpow2 && (1ull << tsh) < t) tsh++;
    NTT_PAIR_STRIDE(j, n) {
        ulonglong2 xv[kEncChunk];
        NTT_UNROLL
        for (unsigned i = 0; i < kEncChunk; i++)
            if (l0 + i < cnt) xv[i] = ld2(ch + (size_t)(l0 + i) * n + j);
        const int d0 = (int)e[j], d1 = (int)e[j + 1];
        const ulonglong2 lv = ld2(clh + j);
        const u64 cl0 = CL_FINISHED ? lv.x : enc_last_limb(lv.x, d0, last, half_last), cl1 = CL_FINISHED ? lv.y : enc_last_limb(lv.y, d1, last, half_last);
        u64 m0 = 0, m1 = 0, f0 = 0, f1 = 0;
        if (h == 0) {
            const ulonglong2 mv = ld2(mp + j);
            m0 = mv.x; m1 = mv.y;
            f0 = tpow2 ? (m0 + tfix) >> tsh : (m0 + tfix) / t;
            f1 = tpow2 ? (m1 + tfix) >> tsh : (m1 + tfix) / t;
        }
        NTT_UNROLL
        for (unsigned i = 0; i < kEncChunk; i++) {
            if (l0 + i >= cnt) break;
            const EncLimb &P = K[i];
            if (ALL_LAZY || P.lazy) {
                // (c_i + e - (c_last - half)) * q_last^-1: sum in (q - 20, 5q), Shoup product in [0, 2q)
                u64 x0 = shoup_mul(xv[i].x + P.bias + (u64)(long long)d0 - cl0, P.inv_q_last, P.inv_q_last_s, P.q);
                u64 x1 = shoup_mul(xv[i].y + P.bias + (u64)(long long)d1 - cl1, P.inv_q_last, P.inv_q_last_s, P.q);
                if (h == 0 && m0 < t && m1 < t) {          // + Delta*m + round-fix < q + 1: below 3q in total
                    x0 = csub(x0 + ((m0 * P.qdt) + f0), 2 * P.q);
                    x1 = csub(x1 + ((m1 * P.qdt) + f1), 2 * P.q);
                } else if (h == 0) {
                    x0 = mod_exact(csub(x0, P.q) + ((m0 * P.qdt) + f0), P.q, P.ratio);
                    x1 = mod_exact(csub(x1, P.q) + ((m1 * P.qdt) + f1), P.q, P.ratio);
                }
                st2(ch + (size_t)(l0 + i) * n + j, csub(x0, P.q), csub(x1, P.q));
                continue;
            }
            if (ALL_LAZY) continue;
            u64 x0 = xv[i].x + signed_to_residue(d0, P.q), x1 = xv[i].y + signed_to_residue(d1, P.q);
            if (x0 > P.q) x0 -= P.q;
            if (x1 > P.q) x1 -= P.q;
            u64 t0, t1;
            if (P.last_lt_2q) { t0 = csub(cl0, P.q); t1 = csub(cl1, P.q); }
            else { t0 = mod_exact(cl0, P.q, P.ratio); t1 = mod_exact(cl1, P.q, P.ratio); }
            if (t0 < P.half_mod) t0 += P.q;
            if (t1 < P.half_mod) t1 += P.q;
            t0 -= P.half_mod; t1 -= P.half_mod;
            if (x0 < t0) x0 += P.q;
            if (x1 < t1) x1 += P.q;
            x0 -= t0; x1 -= t1;
            x0 = mul_const(x0, P.inv_q_last, P.inv_q_last_s, P.fast != 0, P.q, P.mu, P.qbit);
            x1 = mul_const(x1, P.inv_q_last, P.inv_q_last_s, P.fast != 0, P.q, P.mu, P.qbit);
            if (h == 0) {
                // a plaintext coefficient below t gives m * floor(q/t) < q and a round-fix of at most 1: the sum is below 2q
                // and the reference's `% q` is one conditional subtraction; anything else takes the exact remainder
                const u64 s0 = x0 + ((m0 * P.qdt) + f0), s1 = x1 + ((m1 * P.qdt) + f1);
                if (m0 < t && m1 < t) { x0 = csub(s0, P.q); x1 = csub(s1, P.q); }
                else { x0 = mod_exact(s0, P.q, P.ratio); x1 = mod_exact(s1, P.q, P.ratio); }
            }
            st2(ch + (size_t)(l0 + i) * n + j, x0, x1);
        }
    }
}
// the dropped limb r-1 keeps the value the reference leaves there (padding of the ciphertext layout).  grid (x, 2, batch)
NTT_KERNEL void k_encrypt_last_limb(u64 *c, const int *es, unsigned n, unsigned r, unsigned batch, LimbArrays L)
{
    NTT_PDL_ENTER();
    (void)batch;
    const u64 last = L.q[r - 1], half_last = last >> 1;
    const unsigned h = blockIdx.y;
    const size_t rn = (size_t)r * n, k = blockIdx.z;
    u64 *p = c + k * 2 * rn + (size_t)h * rn + (size_t)(r - 1) * n;
    const int *e = es + k * 2 * n + (size_t)h * n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 lv = ld2(p + j);
        st2(p + j, enc_last_limb(lv.x, e[j], last, half_last), enc_last_limb(lv.y, e[j + 1], last, half_last));
    }
}

// c1 = NTT(c1) (.) sk    (barrett_batch, bfv_decryption.cuh:100).  c1 of item k = c + k*item_stride + c1_off.  grid (x, rp, batch)
NTT_KERNEL void k_decrypt_mul(u64 *c, size_t item_stride, size_t c1_off, const u64 *sk, size_t sk_stride, unsigned n, unsigned rp,
                              unsigned batch, LimbArrays L)
{
    NTT_PDL_ENTER();
    (void)batch; (void)rp;
    const unsigned l = blockIdx.y;
    const size_t k = blockIdx.z;
    const u64 q = L.q[l], mu = L.mu[l];
    const int qb = (int)L.qbit[l];
    u64 *p = c + k * item_stride + c1_off + (size_t)l * n;
    const u64 *s = sk + k * sk_stride + (size_t)l * n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 a = ld2(p + j), b = ld2(s + j);
        st2(p + j, barrett_ref(a.x, b.x, q, mu, qb), barrett_ref(a.y, b.y, q, mu, qb));
    }
}
struct DecryptConsts {
    u64 t, gamma, mu_gamma, gamma_div_2, neg_inv_t, neg_inv_gamma;
    int gamma_bits;
    unsigned rp;          // limbs after the drop (the driver's q_amount after q_amount--)
    const u64 *bcm;       // [2][rp]: prod_{i != j} q_i mod t | mod gamma   (demo.cu:248-264)
};
struct DecLimb { u64 q, mu, ptg, ipq, c12, c12_s, bt, bg, bg_s; int qbit, fast; };   // c12 = ptg * ipq mod q
// (acc + v) mod gamma for acc < gamma and v < 2*gamma (the reference's Barrett may leave v in [gamma, 2 gamma)): the sum is
// below 3*gamma, so two conditional subtractions are the exact remainder the reference's `%` computes.
__host__ __device__ __forceinline__ u64 add_mod_gamma(u64 acc, u64 v, u64 gamma)
{
    u64 s = acc + v;
    if (s >= gamma) s -= gamma;
    if (s >= gamma) s -= gamma;
    return s;
}
__device__ __forceinline__ void dec_stage_limbs(DecLimb *K, unsigned first, unsigned count, const DecryptConsts &D, const LimbArrays &L)
{
    if (threadIdx.x < count) {
        const unsigned l = first + threadIdx.x;
        DecLimb e;
        e.q = L.q[l]; e.mu = L.mu[l]; e.qbit = (int)L.qbit[l];
        e.ptg = L.prod_t_gamma_mod_q[l]; e.ipq = L.inv_punctured_q[l];
        e.bt = D.bcm[l]; e.bg = D.bcm[l + D.rp];
        // fast: both the limb's and gamma's Barrett are provably exact -> two Shoup products replace three Barrett sequences
        e.fast = barrett_is_exact(e.q, e.mu, e.qbit) && barrett_is_exact(D.gamma, D.mu_gamma, D.gamma_bits) && e.ptg < e.q && e.ipq < e.q &&
                 e.bg < D.gamma;
        e.c12 = e.fast ? (u64)((unsigned __int128)e.ptg * e.ipq % e.q) : 0;
        e.c12_s = e.fast ? shoup_companion(e.c12, e.q) : 0;
        e.bg_s = e.fast ? shoup_companion(e.bg, D.gamma) : 0;
        K[threadIdx.x] = e;
    }
    __syncthreads();
}
// one limb's contribution: c1 += c0 (`>` quirk), *= t*gamma, *= punctured inverse, accumulate both base conversions
template <bool ALL_FAST>
__device__ __forceinline__ void dec_accumulate(const DecLimb &P, u64 a1, u64 a0, u32 mask32, const DecryptConsts &D, u64 &acc_t, u64 &acc_g)
{
    u64 v = a1 + a0;
    if (!ALL_FAST && v > P.q) v -= P.q;        // the `>` quirk only matters to the literal Barrett sequence; the Shoup product takes any input
    u64 g;
    if (ALL_FAST || P.fast) {
        v = csub(shoup_mul(v, P.c12, P.c12_s, P.q), P.q);                  // (v * t*gamma) * punctured inverse, canonical
        g = csub(shoup_mul(v, P.bg, P.bg_s, D.gamma), D.gamma);
    } else if (!ALL_FAST) {
        v = barrett_ref(v, P.ptg, P.q, P.mu, P.qbit);
        v = barrett_ref(v, P.ipq, P.q, P.mu, P.qbit);
        g = barrett_ref(v, P.bg, D.gamma, D.mu_gamma, D.gamma_bits);
    } else {
        g = 0;
    }
    acc_t += (u64)(((u32)v * (u32)P.bt) & mask32);     // only the low 32 bits survive the mask: one 32-bit multiply (bt < t)
    acc_g = add_mod_gamma(acc_g, g, D.gamma);
}
// both coefficients of pair j over `count` limbs, four limbs' loads (8 x 16 bytes) in flight at a time
template <bool ALL_FAST>
__device__ __forceinline__ void dec_sum_limbs(const DecLimb *K, const u64 *c0, const u64 *c1, unsigned n, u32 j, unsigned count, u32 mask32,
                                              const DecryptConsts &D, u64 &at0, u64 &at1, u64 &ag0, u64 &ag1)
{
    for (unsigned l0 = 0; l0 < count; l0 += 4) {
        ulonglong2 a1[4], a0[4];
        NTT_UNROLL
        for (unsigned i = 0; i < 4; i++)
            if (l0 + i < count) { a1[i] = ld2(c1 + (size_t)(l0 + i) * n + j); a0[i] = ld2(c0 + (size_t)(l0 + i) * n + j); }
        NTT_UNROLL
        for (unsigned i = 0; i < 4; i++) {
            if (l0 + i >= count) break;
            dec_accumulate<ALL_FAST>(K[l0 + i], a1[i].x, a0[i].x, mask32, D, at0, ag0);
            dec_accumulate<ALL_FAST>(K[l0 + i], a1[i].y, a0[i].y, mask32, D, at1, ag1);
        }
    }
}
__device__ __forceinline__ u64 dec_finish_one(u64 acc_t, u64 acc_g, u32 mask32, const DecryptConsts &D)
{
    u64 mt = acc_t & (u64)mask32;
    u64 mg = acc_g >= D.gamma ? acc_g % D.gamma : acc_g;
    mt = (mt * D.neg_inv_t) & (u64)mask32;                                       // mod_t
    mg = barrett_ref(mg, D.neg_inv_gamma, D.gamma, D.mu_gamma, D.gamma_bits);       // barrett_int
    const u64 tmask = D.t - 1;
    return mg > D.gamma_div_2 ? ((mt + (D.gamma - mg)) & tmask) : ((mt - mg) & tmask);
}
// poly_add_xq_d, poly_mul_int_xq_prodtgamma, poly_mul_int_xq_invpq (bfv_decryption.cuh:13-57), fast_convert_array_kernel_t,
// _gamma (poly_arithmetic.cuh:217-251), mod_t, barrett_int, dec_round_kernel (:128-141, :100-126, :253-263) in one pass.
// Writes n plaintext coefficients per item to out + k*out_stride.  grid (x, batch)
template <bool ALL_FAST>
NTT_KERNEL void __launch_bounds__(256, ALL_FAST ? 3 : 2)
k_decrypt_epilogue(const u64 *c, size_t item_stride, size_t c1_off, u64 *out, size_t out_stride, unsigned n, unsigned batch,
                                   DecryptConsts D, LimbArrays L)
{
    NTT_PDL_ENTER();
    (void)batch;
    NTT_SHARED DecLimb K[kMaxLimbs];
    dec_stage_limbs(K, 0, D.rp, D, L);
    const u32 mask32 = (u32)(D.t - 1);
    const size_t k = blockIdx.y;
    const u64 *c0 = c + k * item_stride, *c1 = c0 + c1_off;
    NTT_PAIR_STRIDE(j, n) {
        u64 at0 = 0, at1 = 0, ag0 = 0, ag1 = 0;
        dec_sum_limbs<ALL_FAST>(K, c0, c1, n, j, D.rp, mask32, D, at0, at1, ag0, ag1);
        st2(out + k * out_stride + j, dec_finish_one(at0, ag0, mask32, D), dec_finish_one(at1, ag1, mask32, D));
    }
}

// ---- limb-sharded decryption (multi-GPU): the same arithmetic split at its only cross-limb reduction -------------------
// Each GPU owns `count` limbs [first, first+count) of every ciphertext as a compact shard c[item][2][count][n] and
// produces partial base-conversion sums  part[item][0][j] = sum (v_l * bcm_t[l] & mask)  (mod 2^64 wrap-around, masked at the end)
// and part[item][1][j] = sum Barrett_gamma(v_l * bcm_g[l]) mod gamma.  One all-reduce(SUM) of `part` across the GPUs
// (at most 8 addends < 2^61: no 64-bit overflow) followed by k_decrypt_finish reproduces k_decrypt_epilogue exactly:
// the reference's running `(acc + v) % gamma` and the plain modular sum are the same residue.
// PACKED (requires rp * (t - 1) < 2^16): per item [gamma sums: n u64][t sums: n 16-bit fields, four per u64] = 1.25 n words instead of
// 2 n -- the fields cannot carry into each other under a 64-bit SUM reduction, so the collective moves 10 bytes per coefficient, not 16.
template <bool PACKED>
NTT_KERNEL void k_decrypt_partial(const u64 *c, size_t item_stride, size_t c1_off, u64 *part, unsigned n, unsigned batch, unsigned first,
                                  unsigned count, DecryptConsts D, LimbArrays L)
{
    (void)batch;
    NTT_SHARED DecLimb K[kMaxLimbs];
    dec_stage_limbs(K, first, count, D, L);          // constants are indexed by the global limb, data by the local one
    const u32 mask32 = (u32)(D.t - 1);
    const size_t k = blockIdx.y;
    const u64 *c0 = c + k * item_stride, *c1 = c0 + c1_off;
    NTT_PAIR_STRIDE(j, n) {
        u64 at0 = 0, at1 = 0, ag0 = 0, ag1 = 0;
        dec_sum_limbs<false>(K, c0, c1, n, j, count, mask32, D, at0, at1, ag0, ag1);
        if (PACKED) {
            u64 *p = part + k * (size_t)(n + n / 4);
            st2(p + j, ag0, ag1);
            reinterpret_cast<u32 *>(p + n)[j >> 1] = (u32)at0 | ((u32)at1 << 16);
        } else {
            st2(part + k * 2 * n + j, at0, at1);
            st2(part + k * 2 * n + n + j, ag0, ag1);
        }
    }
}
// OUT16: plaintext coefficients as 16-bit words (t <= 2^16; what the final gather of a sharded decryption moves).
// slots > 1: part_sum holds `slots` contributions slot_stride words apart (the peer-to-peer exchange deposits every rank's partial sums
// in its own slot at the block's owner) and the 64-bit SUM a collective would have formed is taken here.
template <bool PACKED, bool OUT16>
NTT_KERNEL void k_decrypt_finish(const u64 *part_sum, void *out, size_t out_stride, unsigned n, unsigned batch, DecryptConsts D, unsigned slots,
                                 size_t slot_stride)
{
    (void)batch;
    const u32 mask32 = (u32)(D.t - 1);
    const size_t k = blockIdx.y;
    NTT_PAIR_STRIDE(j, n) {
        u64 t0 = 0, t1 = 0, g0 = 0, g1 = 0;
        for (unsigned sl = 0; sl < slots; sl++) {
            const u64 *base = part_sum + (size_t)sl * slot_stride;
            if (PACKED) {
                const u64 *p = base + k * (size_t)(n + n / 4);
                const ulonglong2 pg = ld2(p + j);
                const u32 w = reinterpret_cast<const u32 *>(p + n)[j >> 1];
                t0 += w & 0xffffu; t1 += w >> 16; g0 += pg.x; g1 += pg.y;
            } else {
                const ulonglong2 pt = ld2(base + k * 2 * n + j), pg = ld2(base + k * 2 * n + n + j);
                t0 += pt.x; t1 += pt.y; g0 += pg.x; g1 += pg.y;
            }
        }
        const u64 r0 = dec_finish_one(t0, g0 % D.gamma, mask32, D), r1 = dec_finish_one(t1, g1 % D.gamma, mask32, D);
        if (OUT16) reinterpret_cast<u32 *>(reinterpret_cast<unsigned short *>(out) + k * out_stride)[j >> 1] = (u32)r0 | ((u32)r1 << 16);
        else st2(reinterpret_cast<u64 *>(out) + k * out_stride + j, r0, r1);
    }
}
// 16-bit plaintext words -> the u64 coefficients of the reference layout.  blocks > 1 (grid y): block b reads `total` words at
// in + b * total and writes them at out + b * out_block_stride (the chunk-major staging of the sharded decryption back to item order).
NTT_KERNEL void k_expand16(const unsigned short *in, u64 *out, size_t total, size_t out_block_stride)
{
    const unsigned short *src = in + (size_t)blockIdx.y * total;
    u64 *dst = out + (size_t)blockIdx.y * out_block_stride;
    NTT_GRID_STRIDE(i, total / 2) {
        const u32 w = reinterpret_cast<const u32 *>(src)[i];
        st2(dst + 2 * i, (u64)(w & 0xffffu), (u64)(w >> 16));
    }
}

// ---- homomorphic operations on ciphertexts in the reference layout (the next callers of the batched NTT) ----------
// ca += cb on the limbs below the dropped one, both halves (canonical in, canonical out).  grid (x, r-1, 2*batch)
NTT_KERNEL void k_ct_add(u64 *ca, const u64 *cb, unsigned n, unsigned r, unsigned batch, const u64 *q)
{
    (void)batch;
    const unsigned l = blockIdx.y;
    const size_t off = (size_t)blockIdx.z * r * n + (size_t)l * n;       // blockIdx.z = item * 2 + half
    const u64 ql = q[l];
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 a = ld2(ca + off + j), b = ld2(cb + off + j);
        st2(ca + off + j, csub(a.x + b.x, ql), csub(a.y + b.y, ql));
    }
}
// c0 += Delta * m + round-fix on the limbs below the dropped one: the scaling encryption applies to its message (weird_m_stuff,
// bfv_encryption.cuh:193-212), so Dec(c + plain m) = m_c + m mod t.  Message coefficients are taken mod t.  grid (x, r-1, batch)
NTT_KERNEL void k_ct_add_plain(u64 *c, const u64 *m, size_t m_stride, unsigned n, unsigned r, unsigned batch, u64 t, const u64 *q,
                               const u64 *qi_div_t)
{
    (void)batch;
    const unsigned l = blockIdx.y;
    const size_t k = blockIdx.z;
    const u64 ql = q[l], qdt = qi_div_t[l], tfix = (t + 1) >> 1;
    u64 *c0 = c + k * 2 * r * n + (size_t)l * n;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 a = ld2(c0 + j), mv = ld2(m + k * m_stride + j);
        const u64 m0 = mv.x % t, m1 = mv.y % t;
        st2(c0 + j, csub(a.x + m0 * qdt + (m0 + tfix) / t, ql), csub(a.y + m1 * qdt + (m1 + tfix) / t, ql));
    }
}
// centred lift of a plaintext polynomial (coefficients reduced mod t first) to every limb below the dropped one:
// P[k][l][j] = m mod q_l for m <= t/2, m - t mod q_l otherwise.  grid (x, items)
NTT_KERNEL void k_plain_lift(const u64 *m, size_t m_stride, u64 *P, unsigned n, unsigned rp, unsigned items, u64 t, const u64 *q)
{
    (void)items;
    const size_t k = blockIdx.y;
    NTT_PAIR_STRIDE(j, n) {
        const ulonglong2 mv = ld2(m + k * m_stride + j);
        const u64 m0 = mv.x % t, m1 = mv.y % t;
        for (unsigned l = 0; l < rp; l++) {
            const u64 ql = q[l];
            st2(P + (k * rp + l) * n + j, m0 <= t / 2 ? m0 : ql - (t - m0), m1 <= t / 2 ? m1 : ql - (t - m1));
        }
    }
}

// ---- compact wire format for ciphertexts (the reference has none beyond a text dump, decryption_test.cu:329-344) ----
// packed ciphertext = for half in {0,1}, for limb l < r-1: the n coefficients as n * qbit_l bits, coefficient j at bit offset
// j * qbit_l of that limb's bit string, little-endian bits in little-endian 64-bit words (n is a multiple of 64, so every limb is
// a whole number of words: n / 64 * qbit_l).  The padding limb is not stored.  One thread = 64 coefficients = qbit_l words.
// grid (x, r-1, 2 * batch); word_off[l] = first word of limb l inside one half; half_words = words per half.
NTT_KERNEL void k_ct_pack(const u64 *c, u64 *packed, unsigned n, unsigned r, unsigned batch, const u32 *qbit, const u32 *word_off, u32 half_words)
{
    (void)batch;
    const unsigned l = blockIdx.y, qb = qbit[l];
    const u64 *src = c + (size_t)blockIdx.z * r * n + (size_t)l * n;                 // blockIdx.z = item * 2 + half
    u64 *dst = packed + (size_t)blockIdx.z * half_words + word_off[l];
    for (u32 g = blockIdx.x * blockDim.x + threadIdx.x; g < n / 64; g += gridDim.x * blockDim.x) {
        const u64 *in = src + (size_t)g * 64;
        u64 *out = dst + (size_t)g * qb;
        u64 acc = 0;
        unsigned have = 0;
        for (unsigned k = 0; k < 64; k++) {
            const u64 v = in[k];
            acc |= v << have;
            if (have + qb >= 64) {
                *out++ = acc;
                acc = have ? v >> (64 - have) : 0;
                have = have + qb - 64;
            } else {
                have += qb;
            }
        }
    }
}
NTT_KERNEL void k_ct_unpack(const u64 *packed, u64 *c, unsigned n, unsigned r, unsigned batch, const u32 *qbit, const u32 *word_off, u32 half_words)
{
    (void)batch;
    const unsigned l = blockIdx.y, qb = qbit[l];
    const u64 mask = qb >= 64 ? ~0ull : ((1ull << qb) - 1);
    u64 *dst = c + (size_t)blockIdx.z * r * n + (size_t)l * n;
    const u64 *src = packed + (size_t)blockIdx.z * half_words + word_off[l];
    for (u32 g = blockIdx.x * blockDim.x + threadIdx.x; g < n / 64; g += gridDim.x * blockDim.x) {
        const u64 *in = src + (size_t)g * qb;
        u64 *out = dst + (size_t)g * 64;
        u64 cur = *in++;
        unsigned have = 64;                                                        // unread bits left in cur
        for (unsigned k = 0; k < 64; k++) {
            u64 v;
            if (have >= qb) {
                v = cur >> (64 - have);
                have -= qb;
            } else {
                v = have ? cur >> (64 - have) : 0;
                cur = *in++;
                v |= cur << have;
                have = 64 - (qb - have);
            }
            out[k] = v & mask;
        }
    }
}

}  // namespace nttb200
