// bfv_batched.cpp -- the native, batched way into libnttb200.so from C++: what demo.cu does for ONE message
// (parameter set-up by hand, keygen_rns -> encryption_rns -> decryption_rns, demo.cu:62-311), here for a batch, plus the operations the
// reference does not have (homomorphic add, plaintext multiply, compact wire format).  Plain C ABI + the CUDA runtime for the buffers.
//
//   g++ -std=c++17 -I include -I /usr/local/cuda/include examples/bfv_batched.cpp -L ntt-cuda_b200/nttb200 -lnttb200 \
//       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/ntt-cuda_b200/nttb200 -o bfv_batched
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#include "nttb200.h"

#define CK(x) do { int e__ = (int)(x); if (e__) { fprintf(stderr, "%s failed: %d (%s)\n", #x, e__, nttb200_error_string(e__)); return 1; } } while (0)

int main()
{
    // demo.cu:43-44: n = 8192, three primes and their primitive 2n-th roots; t and gamma of demo.cu:28, :93
    const unsigned n = 8192, r = 3, batch = 16;
    const nttb200_u64 q[r] = {274877562881ull, 274877202433ull, 274877153281ull}, psi[r] = {71485851ull, 33872056ull, 22399294ull};
    const nttb200_u64 t = 1024, gamma = 2305843009213683713ull;
    nttb200_bfv *bfv = nullptr;
    CK(nttb200_bfv_create(&bfv, n, r, q, psi, t, gamma));
    const size_t rn = (size_t)r * n;
    nttb200_u64 *sk, *pk, *c, *c2, *m, *out, *packed;
    CK(cudaMalloc(&sk, rn * 8)); CK(cudaMalloc(&pk, 2 * rn * 8));
    CK(cudaMalloc(&c, batch * 2 * rn * 8)); CK(cudaMalloc(&c2, batch * 2 * rn * 8));
    CK(cudaMalloc(&m, (size_t)batch * n * 8)); CK(cudaMalloc(&out, (size_t)batch * n * 8));
    const size_t words = nttb200_bfv_packed_words(bfv);
    CK(cudaMalloc(&packed, batch * words * 8));
    std::vector<nttb200_u64> hm((size_t)batch * n), hout((size_t)batch * n);
    for (size_t i = 0; i < hm.size(); i++) hm[i] = (i * 2654435761u >> 7) % t;
    CK(cudaMemcpy(m, hm.data(), hm.size() * 8, cudaMemcpyHostToDevice));

    CK(nttb200_bfv_keygen(bfv, sk, pk, 1, /*nonce0*/ 0, nullptr));
    CK(nttb200_bfv_load_keys(bfv, sk, pk, nullptr));                          // fused key kernels from here on (NULL key pointers)
    CK(nttb200_bfv_encrypt(bfv, c, nullptr, 0, m, batch, /*nonce0*/ 1, nullptr));
    CK(nttb200_bfv_pack(bfv, packed, c, batch, nullptr));                      // what would go over the wire
    CK(nttb200_bfv_unpack(bfv, c2, packed, batch, nullptr));
    CK(nttb200_bfv_add(bfv, c2, c, batch, nullptr));                           // Enc(m) + Enc(m)
    CK(nttb200_bfv_decrypt(bfv, out, c2, nullptr, 0, batch, nullptr));
    CK(cudaMemcpy(hout.data(), out, hout.size() * 8, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (size_t i = 0; i < hm.size(); i++) bad += hout[i] != (2 * hm[i]) % t;
    printf("batch %u, n %u, %u limbs: packed ciphertext %zu bytes (reference layout %zu), Dec(unpack(pack(Enc(m))) + Enc(m)) == 2m mod t: %s\n", batch, n, r,
           words * 8, 2 * rn * 8, bad ? "WRONG" : "ok");
    nttb200_bfv_destroy(bfv);
    return bad ? 1 : 0;
}
