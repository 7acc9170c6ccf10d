// bfv_encryption.cuh (drop-in) -- encryption_rns with the reference's signature (BFV_Scheme/bfv_encryption.cuh:223).
#pragma once
#include <cstdio>
#include <vector>
using std::vector;

#include "distributions.cuh"
#include "ntt_60bit.cuh"
#include "poly_arithmetic.cuh"
#include "salsa_common.h"
#include "uint128.h"

#ifndef small_block
#define small_block 128
#endif

// u, streams, q, q_bit_lengths, mu_array and inv_q_last_mod_q are accepted and unused, as in the reference (the
// constants come from q_cons / mu_cons / q_bit_cons / inv_q_last_mod_q_cons).  q_array_device must hold the same
// moduli as q_cons (it does in every reference driver).  c = [c0 | c1], limb q_amount-1 of each half is padding.
inline void encryption_rns(unsigned long long *c, unsigned long long *public_key, unsigned char *in, unsigned long long **u, unsigned long long *e,
                           unsigned n, cudaStream_t streams[], unsigned long long *q, vector<unsigned> q_bit_lengths,
                           vector<unsigned long long> mu_array, vector<unsigned long long> inv_q_last_mod_q, unsigned long long *psi_table_device,
                           unsigned long long *psiinv_table_device, unsigned long long *m_poly_device, unsigned long long *qi_div_t_rns_array_device,
                           unsigned long long *q_array_device, unsigned t, int q_amount)
{
    (void)u; (void)streams; (void)q; (void)q_bit_lengths; (void)mu_array; (void)inv_q_last_mod_q; (void)q_array_device;
    const auto &k = nttb200_dropin::const_addrs();
    nttb200_ref_encryption_rns(c, public_key, in, e, n, psi_table_device, psiinv_table_device, m_poly_device, qi_div_t_rns_array_device, t,
                               (unsigned)q_amount, k.q, k.mu, k.qbit, k.inv_q_last_mod_q, 0);
}
