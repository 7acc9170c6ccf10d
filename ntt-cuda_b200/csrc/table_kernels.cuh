// table_kernels.cuh -- twiddle tables and Shoup companions generated ON THE DEVICE.
//
// Replaces the host-side table fill every reference driver runs before its first transform: fillTablePsi128
// (parameter.h:5-12) = n independent bit-serial modpow128 calls per limb and direction (helper.h:8-28, uint128.h:278-341:
// seconds at n = 2^17 x 16 limbs), followed by one cudaMemcpy per limb (demo.cu:188-196).
// Same values, same layout: table[i] = root^bitrev_logn(i).  Construction: bitrev(m + j) = bitrev(m) + bitrev(j) for
// m = 2^k > j, hence table[m + j] = table[m] * table[j] with table[m] = root^(n / 2m): one modular multiplication per
// entry, log2(n) dependent levels, one CTA per (limb, direction).
#pragma once
#include "modarith.cuh"

namespace nttb200 {

__host__ __device__ __forceinline__ u64 mulmod_slow(u64 a, u64 b, u64 q) { return (u64)((unsigned __int128)a * b % q); }
__host__ __device__ __forceinline__ u64 companion_of(u64 w, u64 q) { return (u64)((((unsigned __int128)w) << 64) / q); }

// grid (limbs, 2): y = 0 builds psi / psi_s from roots[limb], y = 1 builds psiinv / psiinv_s from roots_inv[limb]
NTT_KERNEL void k_build_tables(u64 *psi, u64 *psi_s, u64 *psiinv, u64 *psiinv_s, const u64 *q_arr, const u64 *roots, const u64 *roots_inv,
                               unsigned logn)
{
    NTT_SHARED u64 pw[2][32];            // root^(2^j) and its companion, j < logn
    const unsigned limb = blockIdx.x, inv = blockIdx.y, n = 1u << logn;
    const u64 q = q_arr[limb];
    u64 *t = (inv ? psiinv : psi) + (size_t)limb * n, *ts = (inv ? psiinv_s : psi_s) + (size_t)limb * n;
    if (threadIdx.x == 0) {
        u64 r = (inv ? roots_inv : roots)[limb] % q;
        for (unsigned j = 0; j < logn; j++) {
            pw[0][j] = r;
            pw[1][j] = companion_of(r, q);
            r = mulmod_slow(r, r, q);
        }
        t[0] = 1;
    }
    __syncthreads();
    for (unsigned k = 0; k < logn; k++) {
        const unsigned m = 1u << k;
        const u64 b = pw[0][logn - 1 - k], bs = pw[1][logn - 1 - k];       // root^(n / 2m)
        for (unsigned j = threadIdx.x; j < m; j += blockDim.x) t[m + j] = csub(shoup_mul(t[j], b, bs, q), q);
        __syncthreads();
    }
    for (unsigned i = threadIdx.x; i < n; i += blockDim.x) ts[i] = companion_of(t[i], q);
}

// companions of caller-supplied tables (nttb200_ctx_create_from_tables): tab_s[i] = floor(tab[i] * 2^64 / q[limb])
NTT_KERNEL void k_build_companions(const u64 *tab, u64 *tab_s, const u64 *q_arr, unsigned logn, unsigned limbs, unsigned polys)
{
    const size_t total = (size_t)polys << logn;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
        tab_s[i] = companion_of(tab[i], q_arr[(i >> logn) % limbs]);       // [..][limbs][n] arrays: limb = polynomial index mod limbs
}

}  // namespace nttb200
