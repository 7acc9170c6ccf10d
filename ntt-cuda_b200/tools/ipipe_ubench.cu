// ipipe_ubench.cu -- measures the integer-pipe peaks the NTT roofline is quoted against (SURVEY.md 8d:
// "integer-pipe peak is not in MEASURED_PEAKS.json -- measure it").  Prints one JSON object.
// Each kernel runs ILP independent dependency chains per thread, 1024 threads x 2 CTAs per SM, and reports
// thread-instructions per clock per SM from clock64() deltas (max over CTAs) and from wall time.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../csrc/modarith.cuh"
#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(1024) k(u64 *out, u32 a0, u32 b0, long long *cyc)
{
    u32 x[ILP], y[ILP];
    u64 w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = a0 + threadIdx.x * 7 + i; y[i] = b0 + i * 3; w[i] = ((u64)x[i] << 32) | y[i]; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(b0));                 // IMAD
            if (OP == 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));               // IMAD.WIDE.U32
            if (OP == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(b0));                 // IMAD.HI.U32
            if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));                                  // IADD3
            if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(b0));              // LOP3
            if (OP == 5) asm volatile("mul.hi.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % ILP] | 1));              // mul.hi.u64
            if (OP == 6) asm volatile("mul.lo.u64 %0, %0, %1;" : "+l"(w[i]) : "l"(w[(i + 1) % ILP] | 1));              // mul.lo.u64
            if (OP == 7) { // Shoup modmul + Harvey CT butterfly on (w[i], w[i^1]) as compiled from C
                u64 q = 0x7fffffd8001ull | ((u64)b0 << 40), tw = w[i] | 1, tws = w[(i + 3) % ILP];
                u64 X = w[i], Y = w[(i + 1) % ILP];
                u64 x2 = X >= 2 * q ? X - 2 * q : X;
                u64 T = Y * tw - __umul64hi(Y, tws) * q;
                w[i] = x2 + T; w[(i + 1) % ILP] = x2 - T + 2 * q;
            }
            if (OP == 8) { // lazy butterfly: exact mul.hi + fused 2-product low chain, no conditional subtraction
                u64 q = 0x7fffffd8001ull | ((u64)b0 << 40), tw = w[i] | 1, tws = w[(i + 3) % ILP];
                u64 X = w[i], Y = w[(i + 1) % ILP];
                u64 T = nttb200::shoup_mul_n(Y, tw, tws, 0 - q);
                w[i] = X + T; w[(i + 1) % ILP] = X - T + 2 * q;
            }
            if (OP == 10) { // 1:1 mix of independent IMAD (fma pipe) and LOP3 (alu pipe): do the two half-rate pipes co-issue?
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(b0));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(a0), "r"(b0));
            }
            if (OP == 11) { // 1:1 mix of IMAD.WIDE and 64-bit add (IADD3 + IADD3.X)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y[i]));
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0), "r"(b0));
            }
            if (OP == 12) { // 64-bit add as add.cc / addc (IADD3 with carry-out predicate + IADD3.X)
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0), "r"(b0));
            }
            if (OP == 13) { // 64-bit a + b - c as nvcc compiles it (3-input IADD3 with two carries + IADD3.X)
                w[i] = w[i] + w[(i + 1) % ILP] - w[(i + 2) % ILP];
            }
            if (OP == 14) { // carry-free 64-bit add: IMAD.WIDE(lo, 1, acc) then a plain 32-bit add into the high word
                u32 tl = y[i], th = x[i];
                asm volatile("{\n\t.reg .u32 l, h;\n\tmad.wide.u32 %0, %1, 1, %0;\n\tmov.b64 {l, h}, %0;\n\tadd.u32 h, h, %2;\n\tmov.b64 %0, {l, h};\n\t}"
                             : "+l"(w[i]) : "r"(tl), "r"(th));
            }
            if (OP == 15) { // INDEPENDENT streams: IMAD.WIDE chain on w[i], 64-bit add.cc/addc chain on (x[i], y[i])
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a0), "r"(b0));
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0), "r"(b0));
            }
            if (OP == 16) { // INDEPENDENT streams: IMAD.WIDE chain + two plain 32-bit adds (no carry)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a0), "r"(b0));
                asm volatile("add.u32 %0, %0, %2;\n\tadd.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(a0), "r"(b0));
            }
            if (OP == 17) { // INDEPENDENT streams: IMAD (lo) + IMAD.WIDE: both on the fma pipe
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a0), "r"(b0));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a0), "r"(b0));
            }
            if (OP == 9) { // same with the approximate quotient (3 wide multiplies)
                u64 q = 0x7fffffd8001ull | ((u64)b0 << 40), tw = w[i] | 1, tws = w[(i + 3) % ILP];
                u64 X = w[i], Y = w[(i + 1) % ILP];
                u64 T = nttb200::shoup_mul_a(Y, tw, tws, 0 - q);
                w[i] = X + T; w[(i + 1) % ILP] = X - T + 4 * q;
            }
        }
    }
    long long t1 = clock64();
    u64 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += x[i] + y[i] + w[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char *name, int sms, int clk_khz, bool last)
{
    int blocks = sms;   // one 1024-thread CTA per SM so the clock64() window covers all resident work
    u64 *out; long long *cyc;
    cudaMalloc(&out, (size_t)blocks * 1024 * 8);
    cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 1024>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<blocks, 1024>>>(out, 3, 5, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long *h = (long long *)malloc(blocks * 8);
    cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < blocks; i++) if (h[i] > mx) mx = h[i];
    double ops_per_sm = 1024.0 * ILP * (double)ITERS;           // thread-ops per SM
    double per_clk = ops_per_sm / (double)mx;
    double total_ops = ops_per_sm * sms;
    printf("  \"%s\": {\"per_clk_per_sm\": %.2f, \"Gops_wall\": %.1f, \"ms\": %.4f, \"cycles\": %lld}%s\n", name, per_clk,
           total_ops / (ms * 1e6), ms, mx, last ? "" : ",");
    (void)clk_khz;
    cudaFree(out); cudaFree(cyc); free(h);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_attr\": %d,\n", p.name, p.multiProcessorCount, clk);
    int s = p.multiProcessorCount;
    run<0>("imad_lo32", s, clk, false);
    run<1>("imad_wide_u32", s, clk, false);
    run<2>("imad_hi_u32", s, clk, false);
    run<3>("iadd3", s, clk, false);
    run<4>("lop3", s, clk, false);
    run<5>("mul_hi_u64", s, clk, false);
    run<6>("mul_lo_u64", s, clk, false);
    run<7>("shoup_ct_butterfly_c", s, clk, false);
    run<8>("shoup_ct_butterfly_lazy_ptx", s, clk, false);
    run<9>("shoup_ct_butterfly_lazy_approx_ptx", s, clk, false);
    run<10>("mix_imad_lop3_pairs", s, clk, false);
    run<11>("mix_imadwide_add64_triples", s, clk, false);
    run<12>("add64_cc_pairs", s, clk, false);
    run<13>("add64_3input_c", s, clk, false);
    run<14>("add64_via_imad_wide", s, clk, false);
    run<15>("indep_imadwide_plus_add64cc", s, clk, false);
    run<16>("indep_imadwide_plus_2add32", s, clk, false);
    run<17>("indep_imadwide_plus_imad", s, clk, true);
    printf("}\n");
    return 0;
}
