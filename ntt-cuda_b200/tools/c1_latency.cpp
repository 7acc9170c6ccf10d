// c1_latency.cpp -- BASELINE config 1 measured the way oracle/ref/ref_dump.cu measures the reference: ONE polynomial, N = 4096, the 58-bit
// prime of parameter.h:43-47, C++ host, back-to-back calls timed with CUDA events after warm-up.
//   stateless   nttb200_ref_forward_ntt / _inverse_ntt (what include/dropin/ntt_60bit.cuh forwards forwardNTT / inverseNTT to)
//   context     nttb200_forward_ntt_batch / _inverse_ntt_batch with num = 1 (Shoup arithmetic on context tables)
//   polymul     forwardNTT x 2, barrett, inverseNTT (60bit_ntt_test.cu:70-80), plain launches and replayed from a CUDA graph
// Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "nttb200.h"

typedef unsigned long long u64;
typedef unsigned __int128 u128;
#define CK(x) do { int r__ = (int)(x); if (r__) { fprintf(stderr, "error %d at %s:%d\n", r__, __FILE__, __LINE__); return 1; } } while (0)

template <class F> static float timed(cudaStream_t st, int iters, F f)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 20; i++) f();
    cudaStreamSynchronize(st);
    cudaEventRecord(e0, st);
    for (int i = 0; i < iters; i++) f();
    cudaEventRecord(e1, st);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    return 1e3f * ms / iters;
}

int main(int argc, char **argv)
{
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    const unsigned n = 4096;
    const u64 q = 288230376135196673ull, psi = 60193018759093ull;
    const int qbit = 58;
    const u64 mu = (u64)(((u128)1 << (2 * qbit)) / q);
    nttb200_ctx *ctx = nullptr;
    CK(nttb200_ctx_create(&ctx, n, 1, &q, &psi));
    const u64 *tab = nullptr, *tabinv = nullptr;
    CK(nttb200_ctx_tables(ctx, &tab, &tabinv));
    std::vector<u64> a(n);
    u64 x = 0x9E3779B97F4A7C15ull;
    for (unsigned i = 0; i < n; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; a[i] = x % q; }
    u64 *d, *d2;
    cudaMalloc(&d, 8 * n); cudaMalloc(&d2, 8 * n);
    cudaMemcpy(d, a.data(), 8 * n, cudaMemcpyHostToDevice); cudaMemcpy(d2, a.data(), 8 * n, cudaMemcpyHostToDevice);
    cudaStream_t st;
    cudaStreamCreate(&st);
    const float sf = timed(st, iters, [&] { nttb200_ref_forward_ntt(d, n, st, q, mu, qbit, tab); });
    const float si = timed(st, iters, [&] { nttb200_ref_inverse_ntt(d, n, st, q, mu, qbit, tabinv); });
    const float cf = timed(st, iters, [&] { nttb200_forward_ntt_batch(ctx, d, 1, 1, st); });
    const float ci = timed(st, iters, [&] { nttb200_inverse_ntt_batch(ctx, d, 1, 1, st); });
    auto polymul = [&] {
        nttb200_ref_forward_ntt(d, n, st, q, mu, qbit, tab);
        nttb200_ref_forward_ntt(d2, n, st, q, mu, qbit, tab);
        nttb200_barrett(d, d2, n, q, mu, qbit, st);
        nttb200_ref_inverse_ntt(d, n, st, q, mu, qbit, tabinv);
    };
    const float pm = timed(st, iters, polymul);
    cudaGraph_t graph; cudaGraphExec_t exec;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    polymul();
    CK(cudaStreamEndCapture(st, &graph));
    CK(cudaGraphInstantiate(&exec, graph, 0));
    const float pg = timed(st, iters, [&] { cudaGraphLaunch(exec, st); });
    // correctness of what was timed: forward then inverse restores the input (stateless path)
    cudaMemcpy(d, a.data(), 8 * n, cudaMemcpyHostToDevice);
    nttb200_ref_forward_ntt(d, n, st, q, mu, qbit, tab);
    nttb200_ref_inverse_ntt(d, n, st, q, mu, qbit, tabinv);
    cudaStreamSynchronize(st);
    std::vector<u64> back(n);
    cudaMemcpy(back.data(), d, 8 * n, cudaMemcpyDeviceToHost);
    printf("{\"n\": %u, \"iters\": %d, \"stateless_fwd_us\": %.3f, \"stateless_inv_us\": %.3f, \"context_fwd_us\": %.3f, \"context_inv_us\": %.3f, "
           "\"stateless_polymul_us\": %.3f, \"graph_polymul_us\": %.3f, \"roundtrip_ok\": %s, \"err\": \"%s\"}\n",
           n, iters, sf, si, cf, ci, pm, pg, back == a ? "true" : "false", cudaGetErrorString(cudaGetLastError()));
    nttb200_ctx_destroy(ctx);
    return back == a ? 0 : 1;
}
