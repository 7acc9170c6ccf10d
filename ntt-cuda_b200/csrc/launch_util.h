// launch_util.h -- helpers shared by the launch translation units (tensor-map encoding, prefetch distance).
#pragma once
#include "internal.h"
#include <cuda.h>

namespace nttb200 {
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode();
unsigned pf_dist_for(int dev, int resident);
int make_tmap_strided(CUtensorMap *m, u64 *a, unsigned logn, unsigned k1, unsigned group_polys, size_t group_stride, unsigned groups);
int make_tmap_contig(CUtensorMap *m, u64 *a, unsigned logn, unsigned group_polys, size_t group_stride, unsigned groups);
unsigned sched_k1(unsigned logn);
}  // namespace nttb200
