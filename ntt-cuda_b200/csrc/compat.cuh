// compat.cuh -- lets the kernel sources in this directory compile twice:
//   * with nvcc for sm_100a (the product), and
//   * with g++ under -DNTTB200_EMU for the CPU thread-per-CUDA-thread emulator in csrc/emu/ that the
//     CPU test-suite uses to check index maths, swizzles and butterfly schedules without a GPU.
// Nothing here is reachable from the product path when NTTB200_EMU is not defined.
#pragma once
#include <stdint.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#ifdef NTTB200_EMU
// ---------------------------------------------------------------------------------------------
#include <cmath>
#include <cstring>
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define NTT_RESTRICT
struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
extern thread_local emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern thread_local unsigned char *emu_dyn_smem;
void __syncthreads();
void __syncwarp();
static inline u64 __umul64hi(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) >> 64); }
template <class T> static inline T __ldg(const T *p) { return *p; }
struct ulonglong2 { u64 x, y; };
static inline ulonglong2 make_ulonglong2(u64 x, u64 y) { ulonglong2 r; r.x = x; r.y = y; return r; }
struct uint4 { u32 x, y, z, w; };
float emu_normcdfinvf(float x);   // double-precision stand-in; GPU parity for this op is pinned on the GPU
#define normcdfinvf emu_normcdfinvf
#define NTT_DYN_SMEM(name) unsigned char *name = emu_dyn_smem
#define NTT_KERNEL static
#define NTT_SHARED static          /* the emulator runs one CTA at a time: a function-local static is CTA-shared */
#define NTT_UNROLL _Pragma("GCC unroll 32")
#define NTT_PDL_ENTER()
#else
// ---------------------------------------------------------------------------------------------
#include <cuda_runtime.h>
#define NTT_RESTRICT __restrict__
#define NTT_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define NTT_KERNEL static __global__   /* header-defined kernels: internal linkage per translation unit */
#define NTT_SHARED __shared__
#define NTT_UNROLL _Pragma("unroll")
// Programmatic dependent launch (launch_pdl in launch_util.h sets the attribute): the NEXT kernel of the stream may be scheduled while
// this one runs, and this one may have been scheduled while its predecessor was still running -- so the first statement of the kernel
// releases its dependents and then waits for the predecessor's completion and memory flush.  Stream order is unchanged; without the
// launch attribute both instructions are no-ops.
#define NTT_PDL_ENTER() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")
#endif
