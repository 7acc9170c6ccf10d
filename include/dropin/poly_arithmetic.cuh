// poly_arithmetic.cuh (drop-in) -- host wrappers of BFV_Scheme/poly_arithmetic.cuh forwarding to libnttb200.so, plus
// the one kernel reference drivers launch themselves: barrett<<<N/256,256>>>(a, b, q, mu, qbit) (60bit_ntt_test.cu:76).
#pragma once
#include <cstdlib>

#include "cuda_runtime.h"
#include "device_launch_parameters.h"

#include "ntt_60bit.cuh"
#include "uint128.h"

// a[i] = a[i] * b[i] mod q with the reference's Barrett sequence (shift by qbit-2, multiply by mu, shift by qbit+2,
// multiply by q, subtract, one conditional correction) on mul.hi.u64 / mul.lo.u64
__global__ void barrett(unsigned long long a[], const unsigned long long b[], unsigned long long q, unsigned long long mu, int qbit)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    const unsigned long long x = a[i], y = b[i];
    const unsigned long long lo = x * y, hi = __umul64hi(x, y);
    const int s1 = qbit - 2, s2 = qbit + 2;
    const unsigned long long e = (lo >> s1) | (hi << (64 - s1));
    const unsigned long long plo = e * mu, phi = __umul64hi(e, mu);
    const unsigned long long est = s2 >= 64 ? phi : ((plo >> s2) | (phi << (64 - s2)));
    unsigned long long r = lo - est * q;
    if (r >= q) r -= q;
    a[i] = r;
}

__host__ inline void dec_round(unsigned long long *input_poly, unsigned long long *result_poly, unsigned long long t, unsigned long long gamma,
                               unsigned long long gamma_div_2, unsigned n, cudaStream_t &stream)
{
    nttb200_dec_round(input_poly, result_poly, t, gamma, gamma_div_2, n, stream);
}
// both conversions are issued on stream2 (the reference splits them over stream 0 and stream2)
__host__ inline void fast_convert_array_kernels(unsigned long long *input_poly, unsigned long long *result_poly, unsigned long long t,
                                                unsigned long long *base_change_matrix_device, unsigned q_amount, unsigned long long gamma,
                                                int gamma_bit_length, unsigned long long mu_gamma, cudaStream_t &stream1, cudaStream_t &stream2,
                                                unsigned n)
{
    (void)stream1;
    nttb200_fast_convert_array(input_poly, result_poly, t, base_change_matrix_device, q_amount, gamma, gamma_bit_length, mu_gamma, n, stream2);
}
__host__ inline void half_poly_mul_device(unsigned long long *device_a, unsigned long long *device_b, unsigned n, cudaStream_t &stream,
                                          unsigned long long q, unsigned long long mu, int bit_length, unsigned long long *psi_powers,
                                          unsigned long long *psiinv_powers)
{
    nttb200_ref_forward_ntt(device_a, n, stream, q, mu, bit_length, psi_powers);
    nttb200_barrett(device_a, device_b, n, q, mu, bit_length, stream);
    nttb200_ref_inverse_ntt(device_a, n, stream, q, mu, bit_length, psiinv_powers);
}
// NTT(a), NTT(b) on their streams, then a *= b on stream2 once stream1's transform has finished
__host__ inline void full_poly_mul_device(unsigned long long *device_a, unsigned long long *device_b, unsigned n, cudaStream_t &stream1,
                                          cudaStream_t &stream2, unsigned long long q, unsigned long long mu, int bit_length,
                                          unsigned long long *psi_powers)
{
    forwardNTTdouble(device_a, device_b, n, stream1, stream2, q, mu, bit_length, psi_powers);
    cudaEvent_t done;
    cudaEventCreateWithFlags(&done, cudaEventDisableTiming);
    cudaEventRecord(done, stream1);
    cudaStreamWaitEvent(stream2, done, 0);
    cudaEventDestroy(done);
    nttb200_barrett(device_a, device_b, n, q, mu, bit_length, stream2);
}
// returns a malloc'ed host array the caller frees; the copy back is complete on return
__host__ inline unsigned long long *full_poly_mul(unsigned long long *host_a, unsigned long long *host_b, unsigned long long *device_a,
                                                  unsigned long long *device_b, unsigned n, cudaStream_t &stream1, cudaStream_t &stream2,
                                                  unsigned long long q, unsigned long long mu, int bit_length, unsigned long long *psi_powers,
                                                  unsigned long long *psiinv_powers)
{
    const size_t bytes = sizeof(unsigned long long) * n;
    unsigned long long *result = (unsigned long long *)malloc(bytes);
    cudaMemcpyAsync(device_a, host_a, bytes, cudaMemcpyHostToDevice, stream1);
    cudaMemcpyAsync(device_b, host_b, bytes, cudaMemcpyHostToDevice, stream2);
    full_poly_mul_device(device_a, device_b, n, stream1, stream2, q, mu, bit_length, psi_powers);
    nttb200_ref_inverse_ntt(device_a, n, stream2, q, mu, bit_length, psiinv_powers);
    cudaMemcpyAsync(result, device_a, bytes, cudaMemcpyDeviceToHost, stream2);
    cudaStreamSynchronize(stream2);
    return result;
}
__host__ inline void poly_add_device(unsigned long long *device_a, const unsigned long long *device_b, unsigned n, cudaStream_t &stream,
                                     unsigned long long q)
{
    nttb200_poly_add(device_a, device_b, n, q, stream);
}
__host__ inline void poly_mul_int(unsigned long long *device_a, const unsigned long long b, unsigned n, cudaStream_t &stream, unsigned long long q,
                                  unsigned long long mu, int bit_length)
{
    nttb200_barrett_int(device_a, b, n, q, mu, bit_length, stream);
}
__host__ inline void poly_mul_int_t(unsigned long long *device_a, const unsigned long long b, unsigned n, cudaStream_t &stream, unsigned long long t)
{
    nttb200_mod_t(device_a, b, n, t, stream);
}
__host__ inline void poly_sub_device(unsigned long long *device_a, const unsigned long long *device_b, unsigned n, cudaStream_t &stream,
                                     unsigned long long q)
{
    nttb200_poly_sub(device_a, device_b, n, q, stream);
}
__host__ inline void poly_negate_device(unsigned long long *device_a, unsigned n, cudaStream_t &stream, unsigned long long q)
{
    nttb200_poly_negate(device_a, n, q, stream);
}
__host__ inline void poly_add_integer_device(unsigned long long *device_a, unsigned long long b, unsigned n, cudaStream_t &stream,
                                             unsigned long long q)
{
    nttb200_poly_add_integer(device_a, b, n, q, stream);
}
__host__ inline void poly_add_integer_device_default(unsigned long long *device_a, unsigned long long b, unsigned n, unsigned long long q)
{
    nttb200_poly_add_integer(device_a, b, n, q, 0);
}
