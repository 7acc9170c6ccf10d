// parameter.h (drop-in) -- same moduli, same roots, same table layout as the reference's BFV_Scheme/parameter.h:
// psiTable[i] = psi^bitrev(i) mod q.  Tables are filled by walking the exponents in natural order (n modular
// multiplications) instead of n independent bit-serial modpow calls.
#pragma once
#include <cmath>

#include "helper.h"

inline void fillTablePsi128(unsigned long long psi, unsigned long long q, unsigned long long psiinv, unsigned long long psiTable[],
                            unsigned long long psiinvTable[], unsigned int n)
{
    const int bits = (int)std::log2((double)n);
    unsigned long long p = 1, pi = 1;
    for (unsigned int e = 0; e < n; e++) {
        const unsigned long long i = bitReverse(e, bits);
        psiTable[i] = p;
        psiinvTable[i] = pi;
        p = (unsigned long long)((unsigned __int128)p * psi % q);
        pi = (unsigned long long)((unsigned __int128)pi * psiinv % q);
    }
}
inline void fillTablePsi128Forward(unsigned long long psi, unsigned long long q, unsigned long long psiTable[], unsigned int n)
{
    const int bits = (int)std::log2((double)n);
    unsigned long long p = 1;
    for (unsigned int e = 0; e < n; e++) {
        psiTable[bitReverse(e, bits)] = p;
        p = (unsigned long long)((unsigned __int128)p * psi % q);
    }
}
inline void fillTablePsi64(unsigned psi, unsigned q, unsigned psiinv, unsigned psiTable[], unsigned psiinvTable[], unsigned int n)
{
    const int bits = (int)std::log2((double)n);
    unsigned long long p = 1, pi = 1;
    for (unsigned int e = 0; e < n; e++) {
        const unsigned long long i = bitReverse(e, bits);
        psiTable[i] = (unsigned)p;
        psiinvTable[i] = (unsigned)pi;
        p = p * psi % q;
        pi = pi * psiinv % q;
    }
}

namespace nttb200_dropin {
struct ParamRow { unsigned long long n, q, psi, psiinv, ninv; unsigned q_bit; };
// one prime per ring degree (64-bit-word design) and the 30-bit archive set
static const ParamRow kParams[] = {
    {2048, 137438691329ull, 22157790ull, 88431458764ull, 137371582593ull, 37},
    {4096, 33538049ull, 2386ull, 26102329ull, 33529861ull, 25},
    {8192, 8796092858369ull, 1734247217ull, 5727406356888ull, 8795019116565ull, 43},
    {16384, 281474976546817ull, 23720796222ull, 129310633907832ull, 281457796677643ull, 48},
    {32768, 36028797017456641ull, 1155186985540ull, 31335194304461613ull, 36027697505828911ull, 55},
};
static const ParamRow kParams30[] = {
    {2048, 536608769ull, 284166ull, 208001377ull, 536346753ull, 29},
    {4096, 33538049ull, 2386ull, 26102329ull, 33529861ull, 25},
    {8192, 8716289ull, 1089ull, 8196033ull, 8715225ull, 24},
    {16384, 13664257ull, 273ull, 8959348ull, 13663423ull, 24},
    {32768, 19070977ull, 377ull, 16642842ull, 19070395ull, 25},
    {65536, 13631489ull, 13ull, 12582913ull, 13631281ull, 24},
};
}  // namespace nttb200_dropin

// unknown n leaves the outputs untouched, like the reference
inline void getParams(unsigned long long &q, unsigned long long &psi, unsigned long long &psiinv, unsigned long long &ninv, unsigned int &q_bit,
                      unsigned long long n)
{
    for (const auto &r : nttb200_dropin::kParams)
        if (r.n == n) { q = r.q; psi = r.psi; psiinv = r.psiinv; ninv = r.ninv; q_bit = r.q_bit; }
}
inline void getParams30(unsigned &q, unsigned &psi, unsigned &psiinv, unsigned &ninv, unsigned &q_bit, unsigned n)
{
    for (const auto &r : nttb200_dropin::kParams30)
        if (r.n == n) { q = (unsigned)r.q; psi = (unsigned)r.psi; psiinv = (unsigned)r.psiinv; ninv = (unsigned)r.ninv; q_bit = r.q_bit; }
}
