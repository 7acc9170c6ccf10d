// bfly_ubench.cu -- register-only replay of the NTT kernels' radix-16 round (ct_stages<4,1> / gs_stages<4,1>) to
// measure what the integer pipes sustain on the REAL butterfly instruction stream (no tile traffic, no shared memory).
// Each thread keeps 16 coefficients, runs ITERS rounds of 32 butterflies with twiddles from a small L1-resident table
// (index varies per iteration so nothing is hoisted), at several occupancies.  Prints one JSON object: SMSP cycles per
// warp-butterfly (lower is better; the issue-slot floor is "instructions per butterfly").
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include "../csrc/ntt_kernels.cuh"
using namespace nttb200;

// twiddles from registers instead of the L1-resident table: isolates the arithmetic from the LDG stream
template <class P> struct RegTw : P {
    typename P::Tw r0, r1;
    __device__ __forceinline__ void init(const NttArgs &A, u32 limb, u32 n) { P::init(A, limb, n); P::load2(2 * (threadIdx.x & 31), r0, r1); }
    __device__ __forceinline__ typename P::Tw load(u32) const { return r0; }
    __device__ __forceinline__ void load2(u32, typename P::Tw &t0, typename P::Tw &t1) const { t0 = r0; t1 = r1; }
};

static const char *g_only = nullptr;

template <class P, bool INV, int MINB>
__global__ void __launch_bounds__(256, MINB) kern(u64 *data, NttArgs A, int iters, long long *cyc)
{
    P pol;
    pol.init(A, 0, 4096);
    u64 v[16];
    u64 *g = data + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16;
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = g[i];
    long long t0 = clock64();
    u32 twbase = 1 + (threadIdx.x & 15);
    for (int it = 0; it < iters; it++) {
        if constexpr (!INV) ct_stages<4, 1>(v, twbase, pol);
        else gs_stages<4, 1, false>(v, twbase, pol);
        twbase = (twbase * 5 + 3) & 127;
        if (twbase == 0) twbase = 1;
    }
    long long t1 = clock64();
#pragma unroll
    for (int i = 0; i < 16; i++) g[i] = v[i];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class P, bool INV, int MINB>
static void run(const char *name, int sms, const NttArgs &A, u64 *data, bool last)
{
    if (g_only && strcmp(g_only, name) != 0) return;
    const int iters = 512;
    int blocks = sms * MINB;
    long long *cyc;
    cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<P, INV, MINB><<<blocks, 256>>>(data, A, iters, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<P, INV, MINB><<<blocks, 256>>>(data, A, iters, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(blocks);
    cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto c : h) if (c > mx) mx = c;
    // warp-butterflies per SMSP: MINB CTAs * 8 warps / 4 SMSPs * iters * 32
    double wb = (double)MINB * 8 / 4 * iters * 32;
    double total_bfly = (double)blocks * 256 * iters * 32;
    printf("  \"%s\": {\"ctas_per_sm\": %d, \"cycles_per_warp_bfly\": %.2f, \"Gbfly_s_wall\": %.1f, \"ms\": %.4f, \"err\": \"%s\"}%s\n", name, MINB,
           (double)mx / wb, total_bfly / (ms * 1e6), ms, cudaGetErrorString(cudaGetLastError()), last ? "" : ",");
    cudaFree(cyc);
}

int main(int argc, char **argv)
{
    if (argc > 1) g_only = argv[1];
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int s = p.multiProcessorCount;
    const u64 q = 36028797017456641ull;      // demo.cu:35 (55-bit)
    const int n = 4096;
    std::vector<u64> tw(n), tws(n);
    u64 x = 0x9E3779B97F4A7C15ull;
    for (int i = 0; i < n; i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        tw[i] = x % q;
        tws[i] = (u64)(((unsigned __int128)tw[i] << 64) / q);
    }
    LimbConst lc = {};
    lc.q = q; lc.twoq = 2 * q; lc.negq = 0 - q; lc.ratio = (u64)((((unsigned __int128)1) << 64) / q);
    lc.ninv = tw[5]; lc.ninv_s = tws[5]; lc.w1ninv = tw[6]; lc.w1ninv_s = tws[6]; lc.qbit = 55; lc.pad = 0u - 0x43300000u * (u32)lc.negq;
    u64 *dtw, *dtws, *data; LimbConst *dlc;
    size_t elems = (size_t)s * 8 * 256 * 16;
    cudaMalloc(&dtw, n * 8); cudaMalloc(&dtws, n * 8); cudaMalloc(&dlc, sizeof(lc)); cudaMalloc(&data, elems * 8);
    cudaMemcpy(dtw, tw.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dtws, tws.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dlc, &lc, sizeof(lc), cudaMemcpyHostToDevice);
    std::vector<u64> init(elems);
    for (size_t i = 0; i < elems; i++) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; init[i] = x % q; }
    cudaMemcpy(data, init.data(), elems * 8, cudaMemcpyHostToDevice);
    NttArgs A = {};
    A.tw = dtw; A.tws = dtws; A.lc = dlc;
    printf("{\n  \"device\": \"%s\", \"sms\": %d,\n", p.name, s);
#define RUN3(P, INV, nm)                                    \
    run<P, INV, 2>(nm "_2cta", s, A, data, false);          \
    run<P, INV, 3>(nm "_3cta", s, A, data, false);          \
    run<P, INV, 4>(nm "_4cta", s, A, data, false);
    RUN3(ShoupPolicy, false, "fwd_shoup_exact")
    RUN3(ShoupLazyPolicy, false, "fwd_lazy_approx")
    RUN3(RegTw<ShoupLazyPolicy>, false, "fwd_lazy_approx_regtw")
    RUN3(RegTw<ShoupLazyInvPolicy>, true, "inv_lazy_approx_regtw")
    RUN3(ShoupLazyInvPolicy, true, "inv_lazy_approx")
    run<ShoupPolicy, true, 3>("inv_shoup_exact_3cta", s, A, data, true);
    printf("}\n");
    return 0;
}
