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
/* same through HOST buffers: chunked H2D -> transform -> D2H on internal streams; synchronous. */
NTTB200_API int nttb200_forward_ntt_batch_host(nttb200_ctx *ctx, const nttb200_u64 *in_host, nttb200_u64 *out_host, unsigned num,
                                               unsigned division);
NTTB200_API int nttb200_inverse_ntt_batch_host(nttb200_ctx *ctx, const nttb200_u64 *in_host, nttb200_u64 *out_host, unsigned num,
                                               unsigned division);

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

#ifdef __cplusplus
}
#endif
#endif /* NTTB200_H */
