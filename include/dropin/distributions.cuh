// distributions.cuh (drop-in) -- Salsa20/20 sampling entry points of BFV_Scheme/distributions.cuh over libnttb200.so.
// The key lives inside the library (the reference keeps it in a __constant__ symbol of the including file); the
// history-dependent tail of generate_random's key (only 24 of 32 bytes are uploaded there) is mirrored, see nttb200.h.
#pragma once
#include <cuda_runtime.h>
#include "device_launch_parameters.h"

#include "nttb200.h"
#include "salsa_common.h"

inline void generate_random(unsigned char *a, unsigned n, cudaStream_t &stream) { nttb200_generate_random(a, n, stream); }
inline void generate_random_default(unsigned char *a, unsigned n) { nttb200_generate_random_default(a, n, 0); }
inline void gaussian_dist(unsigned *in, unsigned long long *out, unsigned n, cudaStream_t &stream, unsigned long long q)
{
    nttb200_gaussian_dist(in, out, n, stream, q);
}
inline void uniform_dist(unsigned long long *in, unsigned long long *out, unsigned n, cudaStream_t &stream, unsigned long long q)
{
    nttb200_uniform_dist(in, out, n, stream, q);
}
inline void ternary_dist(unsigned char *in, unsigned long long *out, unsigned n, cudaStream_t &stream, unsigned long long q)
{
    nttb200_ternary_dist(in, out, n, stream, q);
}
