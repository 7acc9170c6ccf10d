// pointwise.cu -- C ABI of the coefficient-wise kernels and the stand-alone sampling entry points
// (drop-in targets for poly_arithmetic.cuh:265-353 and distributions.cuh:220-297).
#include "internal.h"

#include <atomic>
#include "bfv_kernels.cuh"

#include <cstring>

using namespace nttb200;

namespace nttb200 {
// grid-stride launch geometry: enough CTAs to fill the machine (multiple of the SM count), never more than the work
dim3 grid_for(size_t total, int threads)
{
    static std::atomic<int> sms{0};      // one process drives GPUs of one kind: the SM count is read once (racing readers store the same value)
    if (!sms) {
        int dev = 0, v = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms = v;
    }
    size_t need = (total + threads - 1) / threads;
    size_t cap = (size_t)sms * 8;
    return dim3((unsigned)(need < cap ? (need ? need : 1) : cap));
}
// The reference keeps the Salsa20 key in a __constant__ symbol that generate_random_default (32 x 0x01, all 32 bytes)
// and generate_random (0x4D, but only the first 24 bytes are uploaded, distributions.cuh:235) overwrite; the tail a
// generate_random call sees therefore depends on history.  The library mirrors that state per process.
static unsigned char g_key_state[32] = {0};
SalsaKey key_from_bytes(const unsigned char *b)
{
    SalsaKey k;
    for (int i = 0; i < 8; i++) k.k[i] = (u32)b[4 * i] | ((u32)b[4 * i + 1] << 8) | ((u32)b[4 * i + 2] << 16) | ((u32)b[4 * i + 3] << 24);
    return k;
}
SalsaKey default_key()
{
    memset(g_key_state, 1, 32);
    return key_from_bytes(g_key_state);
}
}  // namespace nttb200

#define ST ((cudaStream_t)stream)
#define LAUNCH(kern, total, ...)                                  \
    do {                                                          \
        if ((total) == 0) return 0;                               \
        kern<<<grid_for((total), 256), 256, 0, ST>>>(__VA_ARGS__); \
        return (int)cudaGetLastError();                           \
    } while (0)

extern "C" {

int nttb200_barrett(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, nttb200_u64 mu, int qbit, void *stream)
{ LAUNCH(k_barrett, (size_t)n, a, a, b, (size_t)n, q, mu, qbit); }
int nttb200_barrett_3param(nttb200_u64 *c, const nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, nttb200_u64 mu, int qbit, void *stream)
{ LAUNCH(k_barrett, (size_t)n, c, a, b, (size_t)n, q, mu, qbit); }
int nttb200_barrett_int(nttb200_u64 *a, nttb200_u64 b, unsigned n, nttb200_u64 q, nttb200_u64 mu, int qbit, void *stream)
{ LAUNCH(k_barrett_int, (size_t)n, a, b, (size_t)n, q, mu, qbit); }
int nttb200_mod_t(nttb200_u64 *a, nttb200_u64 b, unsigned n, nttb200_u64 t, void *stream)
{ LAUNCH(k_mod_t, (size_t)n, a, b, (size_t)n, t); }
int nttb200_poly_add(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, void *stream)
{ LAUNCH(k_poly_add, (size_t)n, a, b, (size_t)n, q); }
int nttb200_poly_add_integer(nttb200_u64 *a, nttb200_u64 b, unsigned n, nttb200_u64 q, void *stream)
{ LAUNCH(k_poly_add_integer, (size_t)n, a, b, (size_t)n, q); }
int nttb200_poly_sub(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, nttb200_u64 q, void *stream)
{ LAUNCH(k_poly_sub, (size_t)n, a, b, (size_t)n, q); }
int nttb200_poly_negate(nttb200_u64 *a, unsigned n, nttb200_u64 q, void *stream)
{ LAUNCH(k_poly_negate, (size_t)n, a, (size_t)n, q); }
int nttb200_divide_and_round_q_last_inplace_loop(nttb200_u64 *input_poly, const nttb200_u64 *rns_poly_minus1, unsigned n, nttb200_u64 base_q_i,
                                                 nttb200_u64 half_mod, nttb200_u64 inv_q_last_mod_q_i, nttb200_u64 mu, int qbit, void *stream)
{ LAUNCH(k_divide_and_round_q_last_inplace_loop, (size_t)n, input_poly, rns_poly_minus1, (size_t)n, base_q_i, half_mod, inv_q_last_mod_q_i, mu, qbit); }

int nttb200_barrett_batch(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, unsigned polys, unsigned division, const nttb200_u64 *q_dev,
                          const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream)
{
    if (!division) return NTTB200_EINVAL;
    LimbArrays L{q_dev, mu_dev, qbit_dev, nullptr, nullptr, nullptr};
    LAUNCH(k_barrett_batch, (size_t)n * polys, a, a, b, n, (size_t)n * polys, division, L);
}
int nttb200_barrett_batch_3param(nttb200_u64 *c, const nttb200_u64 *a, const nttb200_u64 *b, unsigned n, unsigned polys, unsigned division,
                                 const nttb200_u64 *q_dev, const nttb200_u64 *mu_dev, const unsigned *qbit_dev, void *stream)
{
    if (!division) return NTTB200_EINVAL;
    LimbArrays L{q_dev, mu_dev, qbit_dev, nullptr, nullptr, nullptr};
    LAUNCH(k_barrett_batch, (size_t)n * polys, c, a, b, n, (size_t)n * polys, division, L);
}

// fast_convert_array_kernels poly_arithmetic.cuh:265-275 (both kernels, one stream) and dec_round :265-268
int nttb200_fast_convert_array(const nttb200_u64 *input_poly, nttb200_u64 *result_poly, nttb200_u64 t, const nttb200_u64 *bcm_dev, unsigned q_amount,
                               nttb200_u64 gamma, int gamma_bits, nttb200_u64 mu_gamma, unsigned n, void *stream)
{
    if (n == 0) return 0;
    k_fast_convert_t<<<grid_for(n, 256), 256, 0, ST>>>(input_poly, result_poly, t, bcm_dev, q_amount, (size_t)n);
    k_fast_convert_gamma<<<grid_for(n, 256), 256, 0, ST>>>(input_poly, result_poly, gamma, bcm_dev, q_amount, gamma_bits, mu_gamma, (size_t)n);
    return (int)cudaGetLastError();
}
int nttb200_dec_round(const nttb200_u64 *input_poly, nttb200_u64 *result_poly, nttb200_u64 t, nttb200_u64 gamma, nttb200_u64 gamma_div_2, unsigned n,
                      void *stream)
{ LAUNCH(k_dec_round, (size_t)n, input_poly, result_poly, t, gamma, gamma_div_2, (size_t)n); }

// ---- sampling -------------------------------------------------------------------------------------------------------------
// generate_random_default distributions.cuh:249-276: floor(nbytes / 64) keystream blocks, key 32 x 0x01, nonce 0
int nttb200_generate_random_default(unsigned char *a, unsigned nbytes, void *stream)
{
    const u64 nblk = nbytes / 64;
    SalsaKey key = default_key();
    LAUNCH(k_salsa20_keystream, (size_t)nblk, a, nblk, (u64)1, (size_t)0, key, (u64)0);
}
// generate_random distributions.cuh:220-247: key 0x4D in bytes 0..23, bytes 24..31 as left by earlier calls
int nttb200_generate_random(unsigned char *a, unsigned nbytes, void *stream)
{
    const u64 nblk = nbytes / 64;
    memset(g_key_state, 77, 24);
    SalsaKey key = key_from_bytes(g_key_state);
    LAUNCH(k_salsa20_keystream, (size_t)nblk, a, nblk, (u64)1, (size_t)0, key, (u64)0);
}
// explicit key / nonce / several streams (batched sampling; stream s uses nonce0 + s)
int nttb200_salsa20_keystream(unsigned char *out, nttb200_u64 blocks_per_stream, nttb200_u64 streams, size_t stream_stride,
                              const unsigned char key[32], nttb200_u64 nonce0, void *stream)
{
    if (!key) return NTTB200_EINVAL;
    SalsaKey k = key_from_bytes(key);
    LAUNCH(k_salsa20_keystream, (size_t)(blocks_per_stream * streams), out, (u64)blocks_per_stream, (u64)streams, stream_stride, k, (u64)nonce0);
}
// gaussian_dist / uniform_dist / ternary_dist distributions.cuh:278-297
int nttb200_gaussian_dist(const unsigned *in, nttb200_u64 *out, unsigned n, void *stream, nttb200_u64 q)
{ LAUNCH(k_convert_gaussian, (size_t)n, in, out, (size_t)n, q); }
int nttb200_uniform_dist(const nttb200_u64 *in, nttb200_u64 *out, unsigned n, void *stream, nttb200_u64 q)
{ LAUNCH(k_convert_range, (size_t)n, in, out, (size_t)n, q); }
int nttb200_ternary_dist(const unsigned char *in, nttb200_u64 *out, unsigned n, void *stream, nttb200_u64 q)
{ LAUNCH(k_convert_ternary, (size_t)n, in, out, (size_t)n, q); }
// the "_xq" converters of bfv_keygen.cuh:14-79 and bfv_encryption.cuh:17-109 as stand-alone entry points
int nttb200_ternary_dist_xq(const unsigned char *in, nttb200_u64 *sk, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream)
{ LAUNCH(k_ternary_dist_xq, (size_t)n * q_amount, in, sk, n, (size_t)n * q_amount, q_dev); }
int nttb200_uniform_dist_xq(const unsigned char *in, nttb200_u64 *pk, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream)
{ LAUNCH(k_uniform_dist_xq, (size_t)n * q_amount, in, pk, n, (size_t)n * q_amount, q_dev); }
int nttb200_gaussian_dist_xq(const unsigned char *in, nttb200_u64 *temp, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream)
{ LAUNCH(k_gaussian_dist_xq, (size_t)n * q_amount, in, temp, n, (size_t)n * q_amount, q_dev); }
int nttb200_convert_ternary_gaussian_x2(const unsigned char *in, nttb200_u64 *c, nttb200_u64 *e, unsigned n, unsigned q_amount,
                                        const nttb200_u64 *q_dev, void *stream)
{ LAUNCH(k_convert_ternary_gaussian_x2, (size_t)n * q_amount, in, c, e, n, q_amount, q_dev); }
int nttb200_poly_add_negate_xq(nttb200_u64 *a, const nttb200_u64 *b, unsigned n, unsigned q_amount, const nttb200_u64 *q_dev, void *stream)
{
    LimbArrays L{q_dev, nullptr, nullptr, nullptr, nullptr, nullptr};
    LAUNCH(k_poly_add_negate_xq, (size_t)n * q_amount, a, b, n, (size_t)n * q_amount, L);
}

}  // extern "C"
