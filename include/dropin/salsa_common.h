// salsa_common.h (drop-in) -- the macros reference drivers use (keygen_test.cu:17,55: convertBlockSize).
#ifndef __SALSA_COMMON_H__
#define __SALSA_COMMON_H__

#if defined(__CUDACC__)
#define MY_ALIGN(n) __align__(n)
#elif defined(__GNUC__)
#define MY_ALIGN(n) __attribute__((aligned(n)))
#else
#define MY_ALIGN(n)
#endif

#ifndef UINT64_MAX
#define UINT64_MAX (18446744073709551615ULL)
#endif
#define ROUNDS 20
#define THREADS_PER_BLOCK (128)
#define XSALSA20_CRYPTO_KEYBYTES 32
#define XSALSA20_CRYPTO_NONCEBYTES 24
#define XSALSA20_BLOCKSZ 64
#define convertBlockSize 64
#define dstdev 3.2
#define dmean 0

#endif
