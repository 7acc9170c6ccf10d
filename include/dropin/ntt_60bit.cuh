// ntt_60bit.cuh (drop-in) -- the reference's NTT call surface (BFV_Scheme/ntt_60bit.cuh) forwarding to libnttb200.so.
//
// The including translation unit still OWNS the six per-limb __constant__ tables, because unmodified drivers fill
// them with cudaMemcpyToSymbol (demo.cu:72,81,125,127,172,245); their device addresses are handed to the library on
// every call.  The tables hold 64 entries instead of 16: that lifts the reference's 16-limb cap and makes the committed
// 16-prime demo -- which copies 8*r bytes into the 4-byte-per-entry q_bit_cons (demo.cu:72) -- fit.
// All functions are asynchronous, return void and discard the library's status, like the reference (no error checks).
#pragma once
#include "cuda_runtime.h"
#include "device_launch_parameters.h"

#include "nttb200.h"
#include "uint128.h"

#define NTTB200_DROPIN_LIMBS 64
__constant__ unsigned long long q_cons[NTTB200_DROPIN_LIMBS];
__constant__ unsigned q_bit_cons[NTTB200_DROPIN_LIMBS];
__constant__ unsigned long long mu_cons[NTTB200_DROPIN_LIMBS];
__constant__ unsigned long long inv_q_last_mod_q_cons[NTTB200_DROPIN_LIMBS];
__constant__ unsigned long long inv_punctured_q_cons[NTTB200_DROPIN_LIMBS];
__constant__ unsigned long long prod_t_gamma_mod_q_cons[NTTB200_DROPIN_LIMBS];

namespace nttb200_dropin {
struct ConstAddrs {
    const unsigned long long *q, *mu, *inv_q_last_mod_q, *inv_punctured_q, *prod_t_gamma_mod_q;
    const unsigned *qbit;
};
// device addresses of this translation unit's constant tables (resolved once)
static inline const ConstAddrs &const_addrs()
{
    static ConstAddrs a = [] {
        ConstAddrs r{};
        cudaGetSymbolAddress((void **)&r.q, q_cons);
        cudaGetSymbolAddress((void **)&r.qbit, q_bit_cons);
        cudaGetSymbolAddress((void **)&r.mu, mu_cons);
        cudaGetSymbolAddress((void **)&r.inv_q_last_mod_q, inv_q_last_mod_q_cons);
        cudaGetSymbolAddress((void **)&r.inv_punctured_q, inv_punctured_q_cons);
        cudaGetSymbolAddress((void **)&r.prod_t_gamma_mod_q, prod_t_gamma_mod_q_cons);
        return r;
    }();
    return a;
}
}  // namespace nttb200_dropin

// two transforms with the same parameters on two streams
__host__ inline void forwardNTTdouble(unsigned long long *device_a, unsigned long long *device_b, unsigned n, cudaStream_t &stream1,
                                      cudaStream_t &stream2, unsigned long long q, unsigned long long mu, int bit_length,
                                      unsigned long long *psi_powers)
{
    nttb200_ref_forward_ntt(device_a, n, stream1, q, mu, bit_length, psi_powers);
    nttb200_ref_forward_ntt(device_b, n, stream2, q, mu, bit_length, psi_powers);
}
__host__ inline void forwardNTT(unsigned long long *device_a, unsigned n, cudaStream_t &stream1, unsigned long long q, unsigned long long mu,
                                int bit_length, unsigned long long *psi_powers)
{
    nttb200_ref_forward_ntt(device_a, n, stream1, q, mu, bit_length, psi_powers);
}
__host__ inline void inverseNTT(unsigned long long *device_a, unsigned n, cudaStream_t &stream1, unsigned long long q, unsigned long long mu,
                                int bit_length, unsigned long long *psiinv_powers)
{
    nttb200_ref_inverse_ntt(device_a, n, stream1, q, mu, bit_length, psiinv_powers);
}
// batch transforms run on the legacy default stream, limb = polynomial index % division
__host__ inline void forwardNTT_batch(unsigned long long *device_a, unsigned n, unsigned long long *psi_powers, unsigned num, unsigned division)
{
    const auto &c = nttb200_dropin::const_addrs();
    nttb200_ref_forward_ntt_batch(device_a, n, psi_powers, num, division, c.q, c.mu, c.qbit, 0);
}
__host__ inline void inverseNTT_batch(unsigned long long *device_a, unsigned n, unsigned long long *psiinv_powers, unsigned num, unsigned division)
{
    const auto &c = nttb200_dropin::const_addrs();
    nttb200_ref_inverse_ntt_batch(device_a, n, psiinv_powers, num, division, c.q, c.mu, c.qbit, 0);
}
