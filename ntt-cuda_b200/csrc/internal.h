// internal.h -- host-side structures shared by the translation units of libnttb200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <vector>

#include "../../include/nttb200.h"
#include "modarith.cuh"

// NTTB200_DEBUG=1 in the environment: every failing step reports its source line on stderr before the code is returned
int nttb200_trace_error(int code, const char *file, int line);
#define NTTB200_CHECK(expr)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) return nttb200_trace_error((int)e__, __FILE__, __LINE__);        \
    } while (0)

#define NTTB200_TRY(x) do { int r__ = (x); if (r__) return nttb200_trace_error(r__, __FILE__, __LINE__); } while (0)
#define KCHECK() do { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return nttb200_trace_error((int)e__, __FILE__, __LINE__); } while (0)

struct nttb200_ctx {
    unsigned n = 0, logn = 0, limbs = 0;
    int device = 0;
    int use_tma = 1;
    int lazy_ok = 1;      // every modulus < 2^58: the forward transform may run without intermediate corrections
    // reference-layout tables [limbs][n] and Shoup companions, device
    u64 *psi = nullptr, *psiinv = nullptr, *psi_s = nullptr, *psiinv_s = nullptr;
    nttb200::LimbConst *lc = nullptr;      // [limbs] device
    u64 *q_dev = nullptr, *mu_dev = nullptr;   // [limbs] device (reference-style constant arrays)
    unsigned *qbit_dev = nullptr;
    std::vector<u64> q, mu;
    std::vector<unsigned> qbit;
    // host-buffer (e2e) pipeline resources, created lazily
    static constexpr int kStages = 3;
    cudaStream_t streams[kStages] = {nullptr, nullptr, nullptr};
    u64 *stage_dev[kStages] = {nullptr, nullptr, nullptr};
    size_t stage_bytes = 0;
    // packed host / wire format (nttb200_pack_polys ...): word_off[l] = words of limbs < l of one group, device copy, packed staging
    std::vector<unsigned> word_off;
    unsigned *word_off_dev = nullptr;
    u64 *stage_packed[kStages] = {nullptr, nullptr, nullptr};
    size_t stage_packed_bytes = 0;
};

namespace nttb200 {

// dir: 0 = forward, 1 = inverse.  policy: 0 = Shoup (ctx tables), 1 = Barrett (reference constants).
struct NttArgsHost;
// policy: 0 = Shoup/Harvey (q < 2^62), 1 = reference Barrett (stateless), 2 = Shoup without intermediate corrections (q < 2^58)
enum { kPolicyShoup = 0, kPolicyBarrett = 1, kPolicyShoupLazy = 2 };
int launch_ntt(bool inverse, int policy, unsigned logn, const NttArgsHost &h, cudaStream_t stream);
int launch_ntt_pass(bool inverse, int policy, unsigned logn, const NttArgsHost &h, int which, cudaStream_t stream);

struct NttArgsHost {
    u64 *a;
    const u64 *tw, *tws;
    const LimbConst *lc;
    const u64 *qv, *muv;
    const unsigned *qbitv;
    u64 q, mu;
    unsigned qbit;
    unsigned num, division;
    int use_tma;
    unsigned group_polys = 0;   // 0 = one contiguous [num][n] array
    size_t group_stride = 0;    // elements between groups
    const unsigned char *gen_src = nullptr;   // forward strided pass: generate the input from keystream bytes (NttArgs::gen_src)
    size_t gen_stride = 0;
};

int launch_fused_mul(bool lazy, unsigned logn, const NttArgsHost &h, const u64 *twi, const u64 *twis, const u64 *key, const u64 *key_s,
                     size_t key_item_stride, size_t key_half_stride, unsigned r, unsigned in_off, unsigned out_off0, unsigned out_off1,
                     unsigned items, int nout, cudaStream_t st);

int launch_polymul(bool lazy, unsigned logn, const NttArgsHost &ha, const u64 *twi, const u64 *twis, const u64 *b, unsigned b_group_polys,
                   size_t b_group_stride, bool fwd, u64 *out, cudaStream_t st);

struct EpiArgs;
int launch_strided_inv_epi(unsigned logn, const NttArgsHost &h, int mode, const EpiArgs &E, cudaStream_t st);

int get_tma_default();

}  // namespace nttb200
