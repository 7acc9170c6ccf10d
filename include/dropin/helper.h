// helper.h (drop-in) -- host helpers with the reference's names and meaning (BFV_Scheme/helper.h).
#pragma once
#include <cstdlib>
#include <random>

#include "uint128.h"

// a^b mod m by square-and-multiply; reductions go through operator% and inherit its x == y quirk
inline unsigned long long modpow128(unsigned long long a, unsigned long long b, unsigned long long mod)
{
    unsigned long long res = (b & 1) ? a : 1;
    while (b != 0) {
        b >>= 1;
        a = (host64x2(a, a) % mod).low;
        if (b & 1) res = (host64x2(res, a) % mod).low;
    }
    return res;
}
inline unsigned modpow64(unsigned a, unsigned b, unsigned mod)
{
    unsigned res = (b & 1) ? a : 1;
    while (b != 0) {
        b >>= 1;
        a = (unsigned)(((unsigned long long)a * a) % mod);
        if (b & 1) res = (unsigned)(((unsigned long long)a * res) % mod);
    }
    return res;
}
// Fermat inverse a^(q-2); the reference also applies it to the non-prime t (demo.cu:109)
inline unsigned long long modinv128(unsigned long long a, unsigned long long q) { return modpow128(a, q - 2, q); }

inline unsigned long long bitReverse(unsigned long long a, int bit_length)
{
    unsigned long long res = 0;
    for (int i = 0; i < bit_length; i++, a >>= 1) res = (res << 1) | (a & 1);
    return res;
}

static std::random_device dev;
static std::mt19937_64 rng(dev());

inline void randomArray128(unsigned long long a[], int n, unsigned long long q)
{
    std::uniform_int_distribution<unsigned long long> randnum(0, q - 1);
    for (int i = 0; i < n; i++) a[i] = randnum(rng);
}
inline void randomArray64(unsigned a[], int n, unsigned q)
{
    std::uniform_int_distribution<unsigned> randnum(0, q);
    for (int i = 0; i < n; i++) a[i] = randnum(rng);
}

// schoolbook negacyclic product, caller frees (the reference's O(n^2) check)
inline unsigned long long *refPolyMul128(unsigned long long a[], unsigned long long b[], unsigned long long m, int n)
{
    unsigned long long *d = (unsigned long long *)malloc(sizeof(unsigned long long) * n);
    for (int k = 0; k < n; k++) {
        unsigned __int128 pos = 0, neg = 0;
        for (int i = 0; i <= k; i++) pos = (pos + (unsigned __int128)a[i] * b[k - i] % m) % m;
        for (int i = k + 1; i < n; i++) neg = (neg + (unsigned __int128)a[i] * b[n + k - i] % m) % m;
        d[k] = (unsigned long long)((pos + m - neg) % m);
    }
    return d;
}
inline unsigned *refPolyMul64(unsigned a[], unsigned b[], unsigned m, int n)
{
    unsigned *d = (unsigned *)malloc(sizeof(unsigned) * n);
    for (int k = 0; k < n; k++) {
        unsigned long long pos = 0, neg = 0;
        for (int i = 0; i <= k; i++) pos = (pos + (unsigned long long)a[i] * b[k - i] % m) % m;
        for (int i = k + 1; i < n; i++) neg = (neg + (unsigned long long)a[i] * b[n + k - i] % m) % m;
        d[k] = (unsigned)((pos + m - neg) % m);
    }
    return d;
}
