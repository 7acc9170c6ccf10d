// launch_util.h -- helpers shared by the launch translation units (tensor-map encoding, prefetch distance).
#pragma once
#include "internal.h"
#include <cuda.h>
#include <cstdlib>
#include <utility>

namespace nttb200 {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
unsigned pf_dist_for(int dev, int resident);
int make_tmap_strided(CUtensorMap *m, u64 *a, unsigned logn, unsigned k1, unsigned group_polys, size_t group_stride, unsigned groups);
int make_tmap_contig(CUtensorMap *m, u64 *a, unsigned logn, unsigned group_polys, size_t group_stride, unsigned groups);
unsigned sched_k1(unsigned logn);

// NTTB200_PDL=0: plain stream serialisation (A/B)
inline bool use_pdl()
{
    static int v = -1;
    if (v < 0) { const char *e = getenv("NTTB200_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
// kernel<<<grid, block, smem, st>>>(args...) with programmatic stream serialisation: for kernels that begin with NTT_PDL_ENTER().
// Only a TINY launch (at most kPdlMaxCtas CTAs: less than half the SMs) that follows another tiny launch on the same stream asks for
// it: there the launch gap is the whole cost -- single-item BFV calls at n = 8192 went from 54 / 49 / 34 us to 41 / 40 / 24 us per
// keygen / encrypt / decrypt, one N = 4096 transform from 6.2 to 4.3 us -- while next to kernels that fill the GPU the early-resident
// successor, waiting in griddepcontrol.wait, costs the running kernel more than the gap it hides (single-item decryption at
// 32768 x 16 limbs, 120-240 CTAs per launch: 37 -> 47 us).  Every other launch serialises exactly as a plain <<<>>> launch does.
constexpr unsigned long long kPdlMaxCtas = 64;
inline bool pdl_allow(cudaStream_t st, unsigned long long ctas)
{
    static thread_local cudaStream_t last_st = nullptr;
    static thread_local bool last_small = false;
    const bool small = ctas <= kPdlMaxCtas;
    const bool allow = use_pdl() && small && last_small && last_st == st;
    last_st = st; last_small = small;
    return allow;
}
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl_allow(st, (unsigned long long)grid.x * grid.y * grid.z) ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
}  // namespace nttb200
#define NTT_LAUNCH_PDL(...)                                                                                                  \
    do {                                                                                                                     \
        const cudaError_t e__ = nttb200::launch_pdl(__VA_ARGS__);                                                             \
        if (e__ != cudaSuccess) return nttb200_trace_error((int)e__, __FILE__, __LINE__);                                    \
    } while (0)
