/*
 * nttb200.h -- C ABI of libnttb200.so: B200-native 60-bit negacyclic NTT/INTT, pointwise RNS kernels,
 * Salsa20 sampling and the BFV keygen / encrypt / decrypt pipelines of ozgunozerk/NTT-Cuda.
 *
 * The reference has no FFI: its API is header inclusion.  Every entry below names the
 * reference host function or kernel it replaces (file:line relative to the reference's BFV_Scheme/).
 * The header-compatible shims in include/dropin/ forward the reference's own C++ signatures to these.
 *
 * Conventions: all pointers are plain device pointers to `unsigned long long` unless the name ends in
 * `_host`; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every function
 * returns 0 on success or a cudaError_t / NTTB200_E* code (nttb200_error_string()).  Calls are
 * asynchronous with respect to the host unless stated.  Nothing allocates on the hot path.
 * Threading: like the reference (global rng / nonce / __constant__ state), one host thread per context; the
 * Salsa20 key state mirrored for generate_random() is per process.
 */
#ifndef NTTB200_H
#define NTTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define NTTB200_API
#else
#define NTTB200_API __attribute__((visibility("default")))
#endif

typedef unsigned long long nttb200_u64;
typedef struct nttb200_ctx nttb200_ctx;

#define NTTB200_EINVAL 10001  /* unsupported n / limbs / null pointer */
#define NTTB200_ENOTMA 10002  /* cuTensorMapEncodeTiled unavailable or failed */
#define NTTB200_ENCCL 10003   /* NCCL not loadable, or a collective failed */
#define NTTB200_MAX_LIMBS 64  /* the reference caps at 16 (__constant__ tables, ntt_60bit.cuh:8-13) */

NTTB200_API int nttb200_version(void);
NTTB200_API const char *nttb200_error_string(int code);

/* ---------------------------------------------------------------------------------------------------
 * Contexts: per-limb constants + twiddle tables + Shoup companions in HBM (replaces the six __constant__
 * tables ntt_60bit.cuh:8-13 and the host-side table upload demo.cu:174-196).  n = 2^11 .. 2^17.
 * ------------------------------------------------------------------------------------------------- */

/* Tables generated from primitive 2n-th roots psi[i] mod q[i] exactly as parameter.h:5-12 fillTablePsi128
 * (psiinv = psi^(q-2), demo.cu:96-97). */
NTTB200_API int nttb200_ctx_create(nttb200_ctx **ctx, unsigned n, unsigned limbs, const nttb200_u64 *q,
                                   const nttb200_u64 *psi_roots);
/* Same, adopting reference-layout HOST tables psi[limbs][n], psiinv[limbs][n] produced by the caller. */
NTTB200_API int nttb200_ctx_create_from_tables(nttb200_ctx **ctx, unsigned n, unsigned limbs, const nttb200_u64 *q,
                                               const nttb200_u64 *psi_tables_host, const nttb200_u64 *psiinv_tables_host);
NTTB200_API void nttb200_ctx_destroy(nttb200_ctx *ctx);
/* Device pointers of the reference-layout tables psi[limbs][n], psiinv[limbs][n] and of q/mu/qbit arrays. */
NTTB200_API int nttb200_ctx_tables(const nttb200_ctx *ctx, const nttb200_u64 **psi, const nttb200_u64 **psiinv);
NTTB200_API int nttb200_ctx_consts(const nttb200_ctx *ctx, const nttb200_u64 **q, const nttb200_u64 **mu, const unsigned **qbit);
/* Synchronous device -> host copy helper (for hosts that own no CUDA runtime of their own, e.g. ctypes callers). */
NTTB200_API int nttb200_download(void *dst_host, const void *src_dev, size_t bytes);
NTTB200_API int nttb200_upload(void *dst_dev, const void *src_host, size_t bytes);
/* 1 = TMA tile movement (default), 0 = plain LDG/STG path (debug).  Env NTTB200_NO_TMA=1 sets 0 at creation. */
NTTB200_API int nttb200_ctx_set_tma(nttb200_ctx *ctx, int enable);

/* ---------------------------------------------------------------------------------------------------
 * NTT / INTT, fast path (Shoup/Harvey lazy butterflies on the context's companion tables).
 * a[num][n] in place; poly p uses limb p % division.
 * ------------------------------------------------------------------------------------------------- */
/* forwardNTT_batch, ntt_60bit.cuh:608-650 */
NTTB200_API int nttb200_forward_ntt_batch(const nttb200_ctx *ctx, nttb200_u64 *a, unsigned num, unsigned division, void *stream);
/* inverseNTT_batch, ntt_60bit.cuh:652-697 */
NTTB200_API int nttb200_inverse_ntt_batch(const nttb200_ctx *ctx, nttb200_u64 *a, unsigned num, unsigned division, void *stream);
/* Profiling hook: launches only the first (which = 0) or second (which = 1) of the transform's two kernels, in
 * execution order (forward: strided pass then contiguous pass; inverse: the mirror).  bench.py times each with events. */
NTTB200_API int nttb200_ntt_pass(const nttb200_ctx *ctx, nttb200_u64 *a, unsigned num, unsigned division, int inverse, int which,
                                 void *stream);
/* a <- a * b in Z_q[X]/(X^n + 1), polynomial p modulo limb p % division: full_poly_mul_device / half_poly_mul_device,
 * poly_arithmetic.cuh:296-310 (forwardNTTdouble, barrett, inverseNTT: 7 kernels, 7.5 HBM passes) as 4 launches / 4.5 passes:
 * the contiguous forward passes of both operands, the coefficient-wise product and the contiguous inverse pass are one kernel.
 * b is clobbered (the reference leaves NTT(b) there; here it holds b after its strided pass). */
NTTB200_API int nttb200_poly_mul_batch(const nttb200_ctx *ctx, nttb200_u64 *a, nttb200_u64 *b, unsigned num, unsigned division, void *stream);
/* a <- INTT(a (.) b) for two operands already in the NTT domain (canonical residues, the layout forward_ntt_batch leaves):
 * barrett_batch + inverseNTT_batch (poly_arithmetic.cuh:36, ntt_60bit.cuh:652) with the product fused into the first inverse
 * kernel.  b is only read. */
NTTB200_API int nttb200_ntt_domain_mul_inverse_batch(const nttb200_ctx *ctx, nttb200_u64 *a, const nttb200_u64 *b, unsigned num,
                                                     unsigned division, void *stream);
/* same through HOST buffers: chunked H2D -> transform -> D2H on internal streams; synchronous. */
NTTB200_API int nttb200_forward_ntt_batch_host(nttb200_ctx *ctx, const nttb200_u64 *in_host, nttb200_u64 *out_host, unsigned num,
                                               unsigned division);
NTTB200_API int nttb200_inverse_ntt_batch_host(nttb200_ctx *ctx, const nttb200_u64 *in_host, nttb200_u64 *out_host, unsigned num,
                                               unsigned division);
/* Compact host / wire format for residue polynomials (the reference's only serialisation is the text dump of
 * decryption_test.cu:329-344): polynomial p (limb p % division) is stored as n * qbit_limb BITS -- coefficient j at bit offset
 * j * qbit_limb, little-endian bits in little-endian 64-bit words -- polynomials back to back in the order of a[num][n].  55-bit
 * residues take 14 % fewer bytes than 64-bit words.  num must be a multiple of division.
 *   nttb200_polys_packed_words                    words of `num` packed polynomials
 *   nttb200_pack_polys / _unpack_polys            device <-> device conversion (canonical residues in, canonical residues out)
 *   nttb200_forward_/inverse_ntt_batch_host_packed  the host-buffer transforms with BOTH host arrays in the packed format: packed
 *       H2D -> unpack -> transform -> pack -> packed D2H, chunks overlapped on internal streams; synchronous. */
NTTB200_API int nttb200_polys_packed_words(const nttb200_ctx *ctx, unsigned num, unsigned division, size_t *words);
NTTB200_API int nttb200_pack_polys(nttb200_ctx *ctx, nttb200_u64 *packed, const nttb200_u64 *a, unsigned num, unsigned division, void *stream);
NTTB200_API int nttb200_unpack_polys(nttb200_ctx *ctx, nttb200_u64 *a, const nttb200_u64 *packed, unsigned num, unsigned division, void *stream);
NTTB200_API int nttb200_forward_ntt_batch_host_packed(nttb200_ctx *ctx, const nttb200_u64 *in_packed_host, nttb200_u64 *out_packed_host,
                                                      unsigned num, unsigned division);
NTTB200_API int nttb200_inverse_ntt_batch_host_packed(nttb200_ctx *ctx, const nttb200_u64 *in_packed_host, nttb200_u64 *out_packed_host,
                                                      unsigned num, unsigned division);

/* ---------------------------------------------------------------------------------------------------
 * NTT / INTT, stateless reference-contract path: nothing but the reference's tables and (q, mu, qbit);
 * arithmetic is the reference's Barrett sequence operation for operation.  Used by include/dropin/.
 * ------------------------------------------------------------------------------------------------- */
/* forwardNTT_batch / inverseNTT_batch with the constants read from device arrays (the q_cons, mu_cons,
 * q_bit_cons symbols owned by the including translation unit). */
NTTB200_API int nttb200_ref_forward_ntt_batch(nttb200_u64 *a, unsigned n, const nttb200_u64 *psi_powers, unsigned num, unsigned division,
                                              const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream);
NTTB200_API int nttb200_ref_inverse_ntt_batch(nttb200_u64 *a, unsigned n, const nttb200_u64 *psiinv_powers, unsigned num, unsigned division,
                                              const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream);
/* forwardNTT ntt_60bit.cuh:314-348, inverseNTT :350-386 (one polynomial, explicit constants) */
NTTB200_API int nttb200_ref_forward_ntt(nttb200_u64 *a, unsigned n, void *stream, nttb200_u64 q, nttb200_u64 mu, int qbit,
                                        const nttb200_u64 *psi_powers);
NTTB200_API int nttb200_ref_inverse_ntt(nttb200_u64 *a, unsigned n, void *stream, nttb200_u64 q, nttb200_u64 mu, int qbit,
                                        const nttb200_u64 *psiinv_powers);

/* ---------------------------------------------------------------------------------------------------
 * Coefficient-wise kernels (poly_arithmetic.cuh).  Reference semantics quirk for quirk: poly_add*
 * subtract q only if the sum is > q, poly_sub never subtracts b, mod_t / fast_convert_t use a 32-bit mask.
 * ------------------------------------------------------------------------------------------------- */
/* barrett<<<n/256,256>>> poly_arithmetic.cuh:9 (a *= b) and the 3-operand form c = a*b */
NTTB200_API int nttb200_barrett(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, nttb200_u64 mu, int qbit, void *stream);
NTTB200_API int nttb200_barrett_3param(nttb200_u64 *c, const nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, nttb200_u64 mu,
                                       int qbit, void *stream);
/* barrett_batch :36, barrett_batch_3param :68 (limb = poly % division; constants from device arrays) */
NTTB200_API int nttb200_barrett_batch(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, unsigned polys, unsigned division,
                                      const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream);
NTTB200_API int nttb200_barrett_batch_3param(nttb200_u64 *c, const nttb200_u64 *a, const nttb200_u64 *b, unsigned n, unsigned polys,
                                             unsigned division, const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev,
                                             void *stream);
/* poly_mul_int :317 (barrett_int :100), poly_mul_int_t :322 (mod_t :128) */
NTTB200_API int nttb200_barrett_int(nttb200_u64 *a, nttb200_u64 b, unsigned n, nttb200_u64 q, nttb200_u64 mu, int qbit, void *stream);
NTTB200_API int nttb200_mod_t(nttb200_u64 *a, nttb200_u64 b, unsigned n, nttb200_u64 t, void *stream);
/* poly_add_device :312, poly_add_integer_device :345, poly_sub_device :327, poly_negate_device :340 */
NTTB200_API int nttb200_poly_add(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, void *stream);
NTTB200_API int nttb200_poly_add_integer(nttb200_u64 *a, nttb200_u64 b, unsigned n, nttb200_u64 q, void *stream);
NTTB200_API int nttb200_poly_sub(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, void *stream);
NTTB200_API int nttb200_poly_negate(nttb200_u64 *a, unsigned n, nttb200_u64 q, void *stream);
/* divide_and_round_q_last_inplace_loop :180 */
NTTB200_API int nttb200_divide_and_round_q_last_inplace_loop(nttb200_u64 *input_poly, const nttb200_u64 *rns_poly_minus1, unsigned n,
                                                             nttb200_u64 base_q_i, nttb200_u64 half_mod, nttb200_u64 inv_q_last_mod_q_i,
                                                             nttb200_u64 mu, int qbit, void *stream);
/* fast_convert_array_kernels :270 (result_poly[0:n] mod t, [n:2n] mod gamma; both on `stream`), dec_round :265 */
NTTB200_API int nttb200_fast_convert_array(const nttb200_u64 *input_poly, nttb200_u64 *result_poly, nttb200_u64 t, const nttb200_u64 *bcm_dev,
                                           unsigned q_amount, nttb200_u64 gamma, int gamma_bits, nttb200_u64 mu_gamma, unsigned n, void *stream);
NTTB200_API int nttb200_dec_round(const nttb200_u64 *input_poly, nttb200_u64 *result_poly, nttb200_u64 t, nttb200_u64 gamma,
                                  nttb200_u64 gamma_div_2, unsigned n, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Sampling (distributions.cuh, salsa_common.h): Salsa20/20 keystream + distribution converters.
 * ------------------------------------------------------------------------------------------------- */
/* generate_random_default :249 (key 32 x 0x01), generate_random :220 (key 0x4D in bytes 0..23, tail = process state) */
NTTB200_API int nttb200_generate_random_default(unsigned char *a, unsigned nbytes, void *stream);
NTTB200_API int nttb200_generate_random(unsigned char *a, unsigned nbytes, void *stream);
/* explicit key; `streams` keystreams of blocks_per_stream 64-byte blocks, stream s = nonce0 + s at out + s*stream_stride */
NTTB200_API int nttb200_salsa20_keystream(unsigned char *out, nttb200_u64 blocks_per_stream, nttb200_u64 streams, size_t stream_stride,
                                          const unsigned char key[32], nttb200_u64 nonce0, void *stream);
/* gaussian_dist :278, uniform_dist :285, ternary_dist :292 */
NTTB200_API int nttb200_gaussian_dist(const unsigned *in, nttb200_u64 *out, unsigned n, void *stream, nttb200_u64 q);
NTTB200_API int nttb200_uniform_dist(const nttb200_u64 *in, nttb200_u64 *out, unsigned n, void *stream, nttb200_u64 q);
NTTB200_API int nttb200_ternary_dist(const unsigned char *in, nttb200_u64 *out, unsigned n, void *stream, nttb200_u64 q);
/* ternary_dist_xq bfv_keygen.cuh:14, uniform_dist_xq :33, gaussian_dist_xq :47, poly_add_negate_xq :81,
 * convert_ternary_gaussian_x2 bfv_encryption.cuh:17 */
NTTB200_API int nttb200_ternary_dist_xq(const unsigned char *in, nttb200_u64 *sk, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream);
NTTB200_API int nttb200_uniform_dist_xq(const unsigned char *in, nttb200_u64 *pk, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream);
NTTB200_API int nttb200_gaussian_dist_xq(const unsigned char *in, nttb200_u64 *temp, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream);
NTTB200_API int nttb200_poly_add_negate_xq(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream);
NTTB200_API int nttb200_convert_ternary_gaussian_x2(const unsigned char *in, nttb200_u64 *c, nttb200_u64 *e, unsigned n, unsigned q_amount,
                                                    const nttb200_u64 *q_dev, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * BFV pipelines, context-based and BATCHED.  Layouts per item (the reference's):
 *   sk[r][n] (NTT domain) . pk[2][r][n] = [pk0 | pk1 = a] (NTT domain) . c[2][r][n] with limb r-1 of each half
 *   left as padding after encryption . m[n] (< t).  Item k samples from Salsa20 nonce nonce0 + k; nonce0 = 0,
 *   batch = 1 reproduces the reference call bit for bit.
 * ------------------------------------------------------------------------------------------------- */
typedef struct nttb200_bfv nttb200_bfv;
/* derives every constant of demo.cu:62-264 (qbit, mu, inv_q_last_mod_q, q_i/t, t*gamma mod q_i, punctured inverses,
 * base-change matrix, -q^-1 mod {t, gamma}, mu_gamma) and the twiddle tables */
NTTB200_API int nttb200_bfv_create(nttb200_bfv **bfv, unsigned n, unsigned limbs, const nttb200_u64 *q, const nttb200_u64 *psi_roots,
                                   nttb200_u64 t, nttb200_u64 gamma);
NTTB200_API void nttb200_bfv_destroy(nttb200_bfv *bfv);
NTTB200_API nttb200_ctx *nttb200_bfv_ctx(nttb200_bfv *bfv);
/* Creation validates what the scheme, as the reference implements it, silently assumes: t a power of two <= 2^32 (32-bit masks,
 * poly_arithmetic.cuh:139,222), every q_i = 1 (mod t) (floor(q_i/t) stands for floor(q/t) mod q_i, bfv_encryption.cuh:193-212; the Fermat
 * "inverse" mod t of demo.cu:109), gamma an odd prime < 2^62 different from every q_i (demo.cu:110) with gamma = 1 (mod t) (dec_round,
 * poly_arithmetic.cuh:253-263, omits the multiplication by gamma^-1 mod t: the reference's gamma is 1 mod 2^11 only, so with it
 * t <= 2048).  Anything else: NTTB200_EINVAL.
 *
 * Limits of the batched calls: batch <= 65535 items per call (nttb200_bfv_add / _pack / _unpack / _mul_plain: 32767), a grid dimension.
 *
 * Threading / streams: a context owns grow-only scratch (keystream, gaussian draws, transformed plaintexts) that every keygen /
 * encrypt / mul_plain call uses on the CALLER's stream: a context serves one stream (and one host thread) at a time -- use one
 * context per stream, or order calls on different streams with events.
 *
 * Randomness: Salsa20/20 keystream under the context's sampling key, item k of a call draws nonce nonce0 + k.  The DEFAULT key is the
 * reference's (32 x 0x01, distributions.cuh:249) and keygen / encryption read the same stream layout, exactly as in the reference, so
 * that nonce0 = 0 reproduces its outputs bit for bit: that default is for parity tests only -- anyone can regenerate such keys.  A
 * deployment sets its own 32-byte key with nttb200_bfv_set_sampling_key and keeps the nonce ranges of key generation and of
 * encryption disjoint (e.g. bit 63 set for encryption). */
NTTB200_API int nttb200_bfv_set_sampling_key(nttb200_bfv *bfv, const unsigned char key[32]);
/* A/B knob (default 1): encryption with a loaded key on a lazy-policy ring fuses `+ e`, the modulus switch and Delta*m into the store
 * of the last inverse NTT kernel (5 launches, c written once); 0 keeps the separate epilogue kernels (7 launches).  Same bits. */
NTTB200_API int nttb200_bfv_set_fused_epilogue(nttb200_bfv *bfv, int enable);
/* pre-sizes the internal keystream scratch so that later calls never allocate */
NTTB200_API int nttb200_bfv_reserve(nttb200_bfv *bfv, unsigned batch);
/* Loads a key pair into the context (device pointers; either may be NULL): private copies + Shoup companions.
 * nttb200_bfv_encrypt(pk = NULL) / nttb200_bfv_decrypt(sk = NULL) then use the loaded key through the FUSED
 * "NTT (.) key -> INTT" kernel (no canonicalisation of NTT(x), no pointwise pass, two HBM passes fewer); same results. */
NTTB200_API int nttb200_bfv_load_keys(nttb200_bfv *bfv, const nttb200_u64 *sk, const nttb200_u64 *pk, void *stream);
/* keygen_rns bfv_keygen.cuh:95 */
NTTB200_API int nttb200_bfv_keygen(nttb200_bfv *bfv, nttb200_u64 *sk, nttb200_u64 *pk, unsigned batch, nttb200_u64 nonce0, void *stream);
/* encryption_rns bfv_encryption.cuh:223; pk_per_item = 0: one public key for the batch; pk = NULL: the loaded key */
NTTB200_API int nttb200_bfv_encrypt(nttb200_bfv *bfv, nttb200_u64 *c, const nttb200_u64 *pk, int pk_per_item, const nttb200_u64 *m,
                                    unsigned batch, nttb200_u64 nonce0, void *stream);
/* decryption_rns bfv_decryption.cuh:76; m_out[batch][n]; c1 of every item is overwritten (as in the reference);
 * sk = NULL: the loaded key */
NTTB200_API int nttb200_bfv_decrypt(nttb200_bfv *bfv, nttb200_u64 *m_out, nttb200_u64 *c, const nttb200_u64 *sk, int sk_per_item,
                                    unsigned batch, void *stream);

/* Compact wire format for ciphertexts (the reference has none beyond the text dump of decryption_test.cu:329-344):
 * for half in {0,1}, for limb l < r-1: n coefficients as n * qbit_l bits, coefficient j at bit j * qbit_l of the limb's bit string,
 * little-endian bits in little-endian 64-bit words; the padding limb is dropped.  nttb200_bfv_packed_words() = 2 * n/64 * sum qbit_l
 * words per ciphertext (32768 x 16 limbs: 6.45 MB instead of 8 MiB).  unpack restores the reference layout with a zeroed padding limb. */
NTTB200_API size_t nttb200_bfv_packed_words(const nttb200_bfv *bfv);
NTTB200_API int nttb200_bfv_pack(nttb200_bfv *bfv, nttb200_u64 *packed, const nttb200_u64 *c, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_unpack(nttb200_bfv *bfv, nttb200_u64 *c, const nttb200_u64 *packed, unsigned batch, void *stream);

/* Keys in the same bit-packed format, ALL r limbs (NTT-domain residues below q_l): polys = r for sk[batch][r][n], 2r for
 * pk[batch][2][r][n]; nttb200_bfv_key_packed_words(bfv, polys) words per key. */
NTTB200_API size_t nttb200_bfv_key_packed_words(const nttb200_bfv *bfv, unsigned polys);
NTTB200_API int nttb200_bfv_pack_key(nttb200_bfv *bfv, nttb200_u64 *packed, const nttb200_u64 *key, unsigned polys, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_unpack_key(nttb200_bfv *bfv, nttb200_u64 *key, const nttb200_u64 *packed, unsigned polys, unsigned batch, void *stream);
/* device ciphertexts <-> packed HOST buffer (the disk / wire side of the format); synchronous */
NTTB200_API int nttb200_bfv_pack_host(nttb200_bfv *bfv, nttb200_u64 *packed_host, const nttb200_u64 *c, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_unpack_host(nttb200_bfv *bfv, nttb200_u64 *c, const nttb200_u64 *packed_host, unsigned batch, void *stream);
/* BFV through HOST buffers (what demo.cu:275-299 times end to end): m_host[batch][n] -> ciphertexts in c_host, and back.  packed != 0:
 * ciphertexts cross PCIe in the wire format above (batch * nttb200_bfv_packed_words() words), else as c[batch][2][r][n].  Loaded keys
 * (nttb200_bfv_load_keys).  H2D, kernels and D2H of successive chunks overlap on three internal streams; synchronous. */
NTTB200_API int nttb200_bfv_encrypt_host(nttb200_bfv *bfv, nttb200_u64 *c_host, int packed, const nttb200_u64 *m_host, unsigned batch,
                                         nttb200_u64 nonce0);
NTTB200_API int nttb200_bfv_decrypt_host(nttb200_bfv *bfv, nttb200_u64 *m_host, const nttb200_u64 *c_host, int packed, unsigned batch);

/* Homomorphic operations on ciphertexts in the reference layout (the reference stops at decryption):
 * c_a <- c_a + c_b  (Dec = m_a + m_b mod t), and  c <- c * p  for a plaintext polynomial p[n] or p[batch][n]
 * (Dec = m * p mod (X^n + 1, t); coefficients of p are taken mod t and lifted centred).  The padding limb is left alone.
 * nttb200_bfv_mul_plain keeps a grow-only device scratch for the transformed plaintext limbs: the first call (or a larger batch)
 * allocates, so warm it up before capturing it into a CUDA graph. */
NTTB200_API int nttb200_bfv_add(nttb200_bfv *bfv, nttb200_u64 *c_a, const nttb200_u64 *c_b, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_add_plain(nttb200_bfv *bfv, nttb200_u64 *c, const nttb200_u64 *m_poly, int plain_per_item, unsigned batch,
                                      void *stream);   /* Dec = m_c + m mod t (the Delta*m scaling of bfv_encryption.cuh:193-212) */
NTTB200_API int nttb200_bfv_mul_plain(nttb200_bfv *bfv, nttb200_u64 *c, const nttb200_u64 *p_poly, int plain_per_item, unsigned batch,
                                      void *stream);

/* Ciphertext x ciphertext multiplication and relinearisation (the paper's stated future work, Article.pdf p.29; the
 * natural caller of the batched NTT and of half_poly_mul_device, poly_arithmetic.cuh:303).  RNS variant of Halevi-Polyakov-Shoup
 * (CT-RSA 2019): base extension to an auxiliary base of r fresh NTT primes, tensor product in the NTT domain, round(t/Q .) in the
 * auxiliary base, conversion back; relinearisation by RNS digits.  Ciphertexts in the reference layout c[batch][2][r][n] (limb r-1 =
 * padding, left alone).  Parity is semantic (the reference has no such operation): Dec(nttb200_bfv_mul(c_a, c_b)) = m_a * m_b mod
 * (X^n + 1, t), and exact against the big-integer oracle (oracle/bfv_mul_oracle.py) up to the documented rounding slack.
 *   nttb200_bfv_relin_keygen   evk_i = (-(a_i s + e_i) + g_i s^2, a_i), i < r-1, from sk[r][n]; digit i samples nonce nonce0 + i
 *   nttb200_bfv_mul_tensor     y[batch][3][r-1][n]: the degree-2 ciphertext, Dec = y0 + y1 s + y2 s^2
 *   nttb200_bfv_relinearize    y -> c_out[batch][2][r][n]
 *   nttb200_bfv_mul            both; c_out may alias an input
 * Work buffers are grow-only (the first call at a batch size allocates).  batch <= 10000.  A batch of two or more runs as two halves
 * on the caller's stream and an internal one (joined before the call returns control of the stream; NTTB200_BFV_SPLIT=0: one stream).
 * Moduli of at most 58 bits (EINVAL otherwise: the lazy sums of the base conversions must not carry). */
NTTB200_API int nttb200_bfv_relin_keygen(nttb200_bfv *bfv, const nttb200_u64 *sk, nttb200_u64 nonce0, void *stream);
NTTB200_API int nttb200_bfv_relin_key(nttb200_bfv *bfv, const nttb200_u64 **evk_dev, size_t *words);   /* evk[r-1][2][r-1][n], NTT domain */
NTTB200_API int nttb200_bfv_mul_tensor(nttb200_bfv *bfv, nttb200_u64 *y, const nttb200_u64 *c_a, const nttb200_u64 *c_b, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_relinearize(nttb200_bfv *bfv, nttb200_u64 *c_out, const nttb200_u64 *y, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_mul(nttb200_bfv *bfv, nttb200_u64 *c_out, const nttb200_u64 *c_a, const nttb200_u64 *c_b, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_mul_aux_base(nttb200_bfv *bfv, nttb200_u64 *p_out, unsigned *count);
/* NTT-friendly primes (parameter generation for ring degrees the reference has no constants for, BASELINE config 5): `count` primes
 * below 2^bits with q = 1 (mod 2n), largest first, skipping `exclude`, and for each the primitive 2n-th root psi = g^((q-1)/2n) of
 * the smallest base g with psi^n = -1.  Host-only. */
NTTB200_API int nttb200_find_ntt_primes(unsigned bits, unsigned n, unsigned count, const nttb200_u64 *exclude, unsigned nexclude,
                                        nttb200_u64 *q_out, nttb200_u64 *psi_out);

/* Limb-sharded encryption needs no extra entry point: create a context for the sub-ring {owned limbs..., last limb}
 * and call nttb200_bfv_encrypt on it -- a limb of the ciphertext depends only on itself, the dropped last limb and the
 * nonce-addressed randomness (whose layout does not depend on the limb count), so the shard equals the same limbs of the
 * full ciphertext bit for bit (nttb200/distributed.py: encrypt_limb_sharded).
 *
 * Limb-sharded decryption across GPUs.  Each GPU holds limbs [first_limb, first_limb + limb_count) of
 * every ciphertext as a shard c_shard[batch][2][shard_half_limbs][n] (per half: the owned limbs first; shard_half_limbs =
 * limb_count for a compact shard, limb_count + 1 for the output of a sub-ring encryption, 0 = compact) and the same limbs
 * of the secret key.  _partial runs NTT / (.)sk / INTT / scaling on the shard and writes the partial base-conversion sums
 * partial[batch][2][n]; the caller all-reduces (SUM, 64-bit) `partial` over the GPUs -- the path's only collective -- and
 * _finish rounds to the plaintext m_out[batch][n].  Bit-identical to nttb200_bfv_decrypt on one GPU. */
NTTB200_API int nttb200_bfv_decrypt_partial(nttb200_bfv *bfv, nttb200_u64 *partial, nttb200_u64 *c_shard, const nttb200_u64 *sk_shard,
                                            int sk_per_item, unsigned first_limb, unsigned limb_count, unsigned shard_half_limbs,
                                            unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_decrypt_finish(nttb200_bfv *bfv, nttb200_u64 *m_out, const nttb200_u64 *partial_sum, unsigned batch, void *stream);

/* ---------------------------------------------------------------------------------------------------
 * Limb-sharded BFV across the GPUs of one box, one process per GPU (BASELINE configs 4-5).
 * The collectives live in the library: NCCL is bound at run time (the libnccl already loaded in the process, e.g.
 * PyTorch's, else libnccl.so.2), so libnttb200.so has no NCCL link dependency.
 *
 * Partition: the batch (a multiple of the world size) is cut into `world` item blocks; block j's plaintext and
 * dropped limb belong to rank j.  Unit (limb l < r-1, block j) has flat index l * world + j; rank g owns
 * [g * (r-1), (g+1) * (r-1)): r-1 tiles on every rank (15 limbs on 8 GPUs balance), and for one block the limbs of a
 * rank are contiguous.  A rank's shard buffer is, block after block, c[items][2][limb_count][n].
 * ------------------------------------------------------------------------------------------------- */
typedef struct nttb200_comm nttb200_comm;
typedef struct nttb200_shard_block {
    unsigned first_item, items;        /* the block's items */
    unsigned first_limb, limb_count;   /* limbs of this block held by the calling rank (limb_count may be 0) */
    size_t offset;                     /* word offset of the block's tile inside the rank's shard buffer */
} nttb200_shard_block;
/* host-only: fills blocks[world] for `rank`; rp = limbs - 1; *shard_words = size of the rank's shard buffer */
NTTB200_API int nttb200_shard_plan(unsigned rp, unsigned n, unsigned batch, unsigned world, unsigned rank, nttb200_shard_block *blocks,
                                   size_t *shard_words);
/* communicator: created from a 128-byte NCCL unique id (rank 0 calls nttb200_comm_unique_id and ships it to the others by its own
 * means), or adopted from an existing ncclComm_t (not destroyed by nttb200_comm_destroy).  world = 1 needs no NCCL at all. */
NTTB200_API int nttb200_comm_unique_id(unsigned char id[128]);
NTTB200_API int nttb200_comm_create(nttb200_comm **comm, const unsigned char id[128], int world, int rank);
NTTB200_API int nttb200_comm_adopt(nttb200_comm **comm, void *nccl_comm, int world, int rank);
/* profiling only: rank `rank` of a world of `world` with every collective skipped (one rank's compute share; results meaningless) */
NTTB200_API int nttb200_comm_fake(nttb200_comm **comm, int world, int rank);
NTTB200_API void nttb200_comm_destroy(nttb200_comm *comm);
NTTB200_API int nttb200_comm_world(const nttb200_comm *comm);
NTTB200_API int nttb200_comm_rank(const nttb200_comm *comm);
/* encryption_rns bfv_encryption.cuh:223, limb-sharded: every rank passes the same m[batch][n] and nonce0 and receives its tiles of
 * the ciphertexts in c_shard (bit-identical to the same limbs of nttb200_bfv_encrypt's output).  Uses the loaded public key.
 * Collective: one all-gather of the finished dropped limbs (and of the signed-byte gaussian draws), overlapped with the transforms. */
NTTB200_API int nttb200_bfv_encrypt_sharded(nttb200_bfv *bfv, nttb200_comm *comm, nttb200_u64 *c_shard, const nttb200_u64 *m, unsigned batch,
                                            nttb200_u64 nonce0, void *stream);
/* decryption_rns bfv_decryption.cuh:76, limb-sharded: c_shard is consumed; m_out[batch][n] is complete on every rank.  Uses the
 * loaded secret key.  The cross-limb sum (poly_arithmetic.cuh:217-251) is the path's one real exchange; block by block:
 *   mode 4 (default): a rank's partial base-conversion sums of a tile (packed to 10 bytes per coefficient) are pushed by the copy
 *           engines into that rank's slot of a buffer at the items' OWNER, mapped through CUDA IPC (no collective kernel takes SMs from
 *           the transforms); owners are visited in rotated order (one incoming stream per GPU at a time); a 4-byte all-reduce per
 *           round is the barrier; the owner sums the slots while rounding (dec_round :253-263) and the round's plaintext is
 *           all-gathered as 16-bit words -- the path's final gather -- under the next round's transforms.
 *   mode 2 / 3: the same with the sums stored into the owner's slot directly by the kernel that forms them (3: one round).
 *           With the default configuration (mode 4, chunks 0) a SMALL call -- a block below 2^24 coefficient-limbs, e.g. 64 ciphertexts
 *           on 8 GPUs -- runs as mode 3: it is latency-bound and one round of direct stores is the shortest chain.
 *           Modes 2-4 fall back to mode 0 on all ranks when CUDA IPC is unavailable.
 *   mode 0: ncclReduce of the sums to the owner, each block in `chunks` pieces, on a second stream next to the transforms.
 *   mode 1: `chunks` ncclReduceScatter calls, rounding, one ncclAllGather. */
NTTB200_API int nttb200_bfv_decrypt_sharded(nttb200_bfv *bfv, nttb200_comm *comm, nttb200_u64 *m_out, nttb200_u64 *c_shard, unsigned batch,
                                            void *stream);
/* Building blocks for callers that run their own collectives: one (limb window, item block) tile of the plan through the fused
 * NTT (.) sk -> INTT kernel up to the partial sums (packed = 1: partial[batch][n + n/4] = gamma sums then four 16-bit t sums per word,
 * needs (r-1)(t-1) < 2^16; packed = 0: partial[batch][2][n]); SUM-reduce (64-bit) over the windows, then _finish_tile rounds
 * (out16 = 1: unsigned short m_out[batch][n], t <= 2^16). */
NTTB200_API int nttb200_bfv_decrypt_partial_tile(nttb200_bfv *bfv, nttb200_u64 *partial, int packed, nttb200_u64 *c_tile, unsigned first_limb,
                                                 unsigned limb_count, unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_decrypt_finish_tile(nttb200_bfv *bfv, void *m_out, int out16, const nttb200_u64 *partial_sum, int packed,
                                                unsigned batch, void *stream);
/* modes as above; chunks = 0: the measured best of the mode.  Env NTTB200_SHARD_MODE / NTTB200_SHARD_CHUNKS set the defaults of new contexts. */
NTTB200_API int nttb200_bfv_shard_config(nttb200_bfv *bfv, int mode, unsigned chunks);
/* device-local conversion between the reference layout c[batch][2][r][n] and the calling rank's shard */
NTTB200_API int nttb200_bfv_shard_from_full(nttb200_bfv *bfv, unsigned world, unsigned rank, nttb200_u64 *c_shard, const nttb200_u64 *c_full,
                                            unsigned batch, void *stream);
NTTB200_API int nttb200_bfv_shard_to_full(nttb200_bfv *bfv, unsigned world, unsigned rank, nttb200_u64 *c_full, const nttb200_u64 *c_shard,
                                          unsigned batch, void *stream);

/* The reference's single-item calls, stateless (tables and constant arrays are the caller's device buffers;
 * unused reference parameters are dropped).  Scratch: `in` as in the reference; the first n (keygen) / 2n (encrypt)
 * 32-bit words of `temp` / `e` receive the signed gaussian draws instead of r*n residues. */
NTTB200_API int nttb200_ref_keygen_rns(unsigned char *in, unsigned q_amount, unsigned n, nttb200_u64 *secret_key, nttb200_u64 *public_key,
                                       nttb200_u64 *temp, const nttb200_u64 *psi_table, const nttb200_u64 *psiinv_table,
                                       const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream);
NTTB200_API int nttb200_ref_encryption_rns(nttb200_u64 *c, const nttb200_u64 *public_key, unsigned char *in, nttb200_u64 *e, unsigned n,
                                           const nttb200_u64 *psi_table, const nttb200_u64 *psiinv_table, const nttb200_u64 *m_poly,
                                           const nttb200_u64 *qi_div_t_dev, nttb200_u64 t, unsigned q_amount, const nttb200_u64 *q_dev,
                                           const nttb200_u64 *mu_dev, const unsigned *qbit_dev, const nttb200_u64 *inv_q_last_mod_q_dev,
                                           void *stream);
/* q_amount = limbs after the drop; the plaintext lands at c + n*(q_amount-1) as in the reference (demo.cu:299) */
NTTB200_API int nttb200_ref_decryption_rns(nttb200_u64 *c, const nttb200_u64 *secret_key, const nttb200_u64 *psi_table,
                                           const nttb200_u64 *psiinv_table, unsigned n, unsigned q_amount,
                                           const nttb200_u64 *base_change_matrix_dev, nttb200_u64 t, nttb200_u64 gamma, nttb200_u64 mu_gamma,
                                           int gamma_bits, nttb200_u64 neg_inv_t, nttb200_u64 neg_inv_gamma, nttb200_u64 gamma_div_2,
                                           const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev,
                                           const nttb200_u64 *inv_punctured_q_dev, const nttb200_u64 *prod_t_gamma_mod_q_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NTTB200_H */
