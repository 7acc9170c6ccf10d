// bfv_internal.h -- BFV context and pipeline descriptors shared by bfv.cu and sharded.cu.
#pragma once
#include "internal.h"
#include "bfv_kernels.cuh"
#include "epi.cuh"

struct nttb200_bfv {
    nttb200_ctx *ctx = nullptr;
    unsigned n = 0, r = 0;            // all limbs; rp = r - 1 after the modulus switch
    u64 t = 0, gamma = 0, mu_gamma = 0, gamma_div_2 = 0, neg_inv_t = 0, neg_inv_gamma = 0;
    int gamma_bits = 0;
    // device constant arrays
    u64 *inv_q_last_mod_q = nullptr, *qi_div_t = nullptr, *prod_t_gamma_mod_q = nullptr, *inv_punctured_q = nullptr, *bcm = nullptr;
    // keys loaded into the context (nttb200_bfv_load_keys): private copies + Shoup companions, enabling the fused
    // "NTT (.) key -> INTT" kernel; used when encrypt / decrypt are called with a NULL key pointer
    u64 *sk_l = nullptr, *sk_ls = nullptr, *pk_l = nullptr, *pk_ls = nullptr;
    // grow-only scratch: keystream and gaussian draws
    unsigned char *ks = nullptr; size_t ks_bytes = 0;
    int *es = nullptr; size_t es_count = 0;
    u64 *pt = nullptr; size_t pt_count = 0;          // lifted + transformed plaintexts of nttb200_bfv_mul_plain
    // fused-epilogue encryption path: u bytes [items][n], gaussian draws as signed bytes [items][2][n], per-limb epilogue constants
    unsigned char *ub = nullptr; size_t ub_bytes = 0;
    signed char *es8 = nullptr; size_t es8_bytes = 0;
    nttb200::EncEpiLimb *enc_epi = nullptr;
    unsigned tsh = 0;
    unsigned char salsa_key[32];                      // sampling key of the context API (default: the reference's 32 x 0x01)
    struct nttb200_shard_state *shard = nullptr;      // scratch + streams of the multi-GPU entry points (sharded.cu)
    unsigned *word_off = nullptr; unsigned half_words = 0;   // compact wire format: first word of each limb inside a half, words per half
    unsigned *key_word_off = nullptr; unsigned key_half_words = 0;   // same for keys (all r limbs)
    struct nttb200_host_state *host = nullptr;        // staging + streams of the host-buffer entry points (bfv_host.cu)
    struct nttb200_mul_state *mul = nullptr;          // auxiliary base, constants, relinearisation key, work buffers (bfv_mul.cu)
    // batched calls run their two halves on two streams (fork / join with events): the HBM-bound element-wise kernels of one half
    // overlap the issue-bound transforms of the other
    cudaStream_t st2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int split = 1;                                    // 0: never split (A/B knob, NTTB200_BFV_SPLIT=0)
    bool enc_lazy = false, dec_fast = false, all_exact = false;
    bool epi_ok = false;                              // every limb qualifies for the epilogue fused into the last inverse kernel
    bool no_fused_epilogue = true;                    // default: separate (HBM-bound) epilogue kernels; false = epilogue in the last inverse kernel's store (A/B, slower)
};
void nttb200_shard_state_destroy(struct nttb200_shard_state *s);
void nttb200_host_state_destroy(struct nttb200_host_state *s);
void nttb200_mul_state_destroy(struct nttb200_mul_state *s);


namespace nttb200 {
struct Pipe {                 // everything one pipeline run needs, independent of the front end
    unsigned n, logn, r;
    LimbArrays L;
    const u64 *qi_div_t;
    // NTT flavour
    int policy_fwd, policy_inv;
    const u64 *psi, *psiinv, *psi_s, *psiinv_s;
    const LimbConst *lc;
    int use_tma;
    cudaStream_t st;
    bool enc_lazy = false, dec_fast = false;   // host-verified: every limb qualifies for the lazy / Shoup-only epilogues
    bool all_exact = false;                    // host-verified: the reference's Barrett is exact for every limb (any exact product = its bits)
    SalsaKey key;                              // sampling key (reference: 32 x 0x01)
};


Pipe pipe_from_bfv(const nttb200_bfv *b, cudaStream_t st);
// restricts a pipeline to limbs [first, first + count): tables / constants start at the first owned limb, local limb = poly % count
Pipe pipe_limb_window(const Pipe &P, const nttb200_ctx *c, unsigned first);
int ensure_enc_scratch(nttb200_bfv *b, size_t ub_bytes, size_t es8_bytes);

// Building blocks of the fused-epilogue encryption (loaded public key, lazy-policy rings), for limbs [first, first + count) of
// `items` ciphertexts laid out c[item][2][slots][n] (limb `first + l` in slot l of each half; item stride 2 * slots * n):
//   enc_front        u (from ub[item][n]) -> strided forward pass -> fused contig forward (.) pk0 | pk1 -> contig inverse (both halves)
//   enc_finish_last  strided inverse pass of the DROPPED limb with `+ e`, rounding offset fused into its store: cl[item][2][n]
//   enc_finish_limbs strided inverse pass of limbs below the dropped one with mod-switch + Delta*m fused into its store
int enc_front(const nttb200_bfv *b, const Pipe &P, u64 *c, unsigned slots, unsigned first, unsigned count, unsigned items, const unsigned char *ub);
int enc_finish_last(const nttb200_bfv *b, const Pipe &P, u64 *cl, size_t cl_item_stride, size_t cl_half_stride, const signed char *es8, unsigned items);
int enc_finish_limbs(const nttb200_bfv *b, const Pipe &P, u64 *c, unsigned slots, unsigned first, unsigned count, unsigned items, const u64 *cl,
                     size_t cl_item_stride, size_t cl_half_stride, const signed char *es8, const u64 *m, size_t m_stride);
// decryption of limbs [first, first + count) up to the cross-limb sums (loaded secret key: fused kernel); c_shard[item][2][slots][n]
int dec_partial(const nttb200_bfv *b, const Pipe &P, u64 *partial, int packed, u64 *c_shard, unsigned slots, unsigned first, unsigned count,
                unsigned items);
int dec_transforms(const nttb200_bfv *b, const Pipe &P, u64 *c_shard, unsigned slots, unsigned first, unsigned count, unsigned items);
int dec_partial_sums(const nttb200_bfv *b, const Pipe &P, u64 *partial, int packed, const u64 *c_shard, unsigned slots, unsigned first, unsigned count,
                     unsigned items);
int dec_finish(const nttb200_bfv *b, void *out, int out16, const u64 *partial_sum, int packed, unsigned items, cudaStream_t st, unsigned slots = 1,
               size_t slot_stride = 0);
int dec_expand16(const unsigned short *in, u64 *out, size_t total, cudaStream_t st, unsigned blocks = 1, size_t out_block_stride = 0);
int enc_sample(const nttb200_bfv *b, unsigned char *ub, signed char *es8, unsigned items, u64 nonce0, int want_u, int want_e, cudaStream_t st);
SalsaKey bfv_salsa_key(const nttb200_bfv *b);
dim3 grid_for(size_t total, int threads);
}  // namespace nttb200
