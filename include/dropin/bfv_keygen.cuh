// bfv_keygen.cuh (drop-in) -- keygen_rns with the reference's signature (BFV_Scheme/bfv_keygen.cuh:95).
#pragma once
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

// q, streams, mu_array and q_bit_lengths are accepted and unused, as in the reference: the moduli come from q_cons.
// Outputs: secret_key[q_amount][n] (NTT domain), public_key = [NTT(-(a*s + e)) | a].  Runs on the legacy default stream.
inline void keygen_rns(unsigned char in[], int q_amount, unsigned long long *q, unsigned n, unsigned long long *secret_key,
                       unsigned long long *public_key, cudaStream_t *streams, unsigned long long *temp, vector<unsigned long long> mu_array,
                       vector<unsigned> q_bit_lengths, unsigned long long *psi_table_device, unsigned long long *psiinv_table_device)
{
    (void)q; (void)streams; (void)mu_array; (void)q_bit_lengths;
    const auto &c = nttb200_dropin::const_addrs();
    nttb200_ref_keygen_rns(in, (unsigned)q_amount, n, secret_key, public_key, temp, psi_table_device, psiinv_table_device, c.q, c.mu, c.qbit, 0);
}
