// bfv_decryption.cuh (drop-in) -- decryption_rns with the reference's signature (BFV_Scheme/bfv_decryption.cuh:76).
#pragma once
#include <stdio.h>
#include <vector>
using std::vector;

#include "ntt_60bit.cuh"
#include "poly_arithmetic.cuh"
#include "uint128.h"

#ifndef small_block
#define small_block 128
#endif

// c holds 2*(q_amount+1)*n words: c0 | padding | c1 | padding (see the reference's comment above decryption_rns);
// the n plaintext coefficients are written to c + n*(q_amount-1), where the reference leaves them (demo.cu:299).
// One stream, no allocation, no stream creation (the reference creates q_amount streams and mallocs per call).
// q, q_bit_lengths, mu_array, inv_punctured_q, output_base and prod_t_gamma_mod_q are accepted and unused, as in the
// reference; the per-limb constants come from the __constant__ tables.
inline void decryption_rns(unsigned long long *c, unsigned long long *secret_key, unsigned long long *q, vector<unsigned> &q_bit_lengths,
                           vector<unsigned long long> &mu_array, unsigned long long *psi_table_device, unsigned long long *psiinv_table_device, int n,
                           unsigned q_amount, vector<unsigned long long> &inv_punctured_q, unsigned long long *base_change_matrix_device,
                           unsigned long long t, unsigned long long gamma, unsigned long long mu_gamma, vector<unsigned long long> &output_base,
                           vector<unsigned> &output_base_bit_lengths, vector<unsigned long long> &neg_inv_q_mod_t_gamma,
                           unsigned long long gamma_div_2, vector<unsigned long long> prod_t_gamma_mod_q)
{
    (void)q; (void)q_bit_lengths; (void)mu_array; (void)inv_punctured_q; (void)output_base; (void)prod_t_gamma_mod_q;
    const auto &k = nttb200_dropin::const_addrs();
    nttb200_ref_decryption_rns(c, secret_key, psi_table_device, psiinv_table_device, (unsigned)n, q_amount, base_change_matrix_device, t, gamma,
                               mu_gamma, (int)output_base_bit_lengths[1], neg_inv_q_mod_t_gamma[0], neg_inv_q_mod_t_gamma[1], gamma_div_2, k.q,
                               k.mu, k.qbit, k.inv_punctured_q, k.prod_t_gamma_mod_q, 0);
}
