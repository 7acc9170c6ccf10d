// ipipe2_ubench.cu -- second-generation integer-pipe micro-benchmark.  The first one (ipipe_ubench.cu) used loop-invariant
// operands, which ptxas folds (two adds into one 3-input IADD3, strength-reduced chains), so its mixed-stream numbers are
// not trustworthy.  Here every instruction consumes a value produced by a neighbouring chain in the same iteration
// (x[i] op= x[(i+1) % ILP]), which cannot be folded or hoisted; the SASS of every kernel was inspected (see
// profiles/r01_ipipe2_sass_counts.txt).  Reports warp-instructions issued per cycle per SM sub-partition (IPC) for
// pure and mixed streams: the question is whether the fma-heavy pipe (IMAD*) and the alu pipe (IADD3/LOP3) overlap.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ILP 8
#define ITERS 2048

template <int OP>
__global__ void __launch_bounds__(256) k(u32 *out, const u32 *in, long long *cyc, u32 cparam)
{
    double d[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) d[i] = 1.0 + 1e-9 * in[threadIdx.x + i];
    u32 x[ILP], y[ILP], z[ILP];
    u64 w[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) w[i] = in[threadIdx.x + i] * 0x100000001ull;
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = in[threadIdx.x * ILP + i]; y[i] = in[4096 + threadIdx.x * ILP + i]; z[i] = in[8192 + threadIdx.x * ILP + i]; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            const int j = (i + 1) % ILP;
            if (OP == 0) {          // IMAD (32-bit): 1 fma-heavy
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[j]), "r"(x[j]));
            } else if (OP == 1) {   // IADD3 without carry: 1 alu
                x[i] = x[i] + x[j] + y[j];          // 3-input IADD3 (a 2-input add may be emitted as IMAD.IADD)
            } else if (OP == 2) {   // LOP3: 1 alu
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[j]), "r"(y[i]));
            } else if (OP == 3) {   // 64-bit add: IADD3 (carry out) + IADD3.X (carry in): 2 alu
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(x[j]), "r"(y[j]));
            } else if (OP == 4) {   // IMAD.WIDE.U32 with 64-bit accumulator: 1 fma-heavy
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[j]), "r"(z[i]));
            } else if (OP == 5) {   // 1 IMAD + 1 IADD3 (no carry), independent registers: do the pipes overlap?
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(z[i]), "r"(x[j]));
                y[i] = y[i] + y[j] + z[j];
            } else if (OP == 6) {   // 1 IMAD + 1 LOP3
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(z[i]), "r"(x[j]));
                asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(y[i]) : "r"(y[j]), "r"(z[j]));
            } else if (OP == 7) {   // 1 IMAD.WIDE + 64-bit add (2 alu) on independent registers
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[j]), "r"(z[i]));
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(x[j]), "r"(y[j]));
            } else if (OP == 8) {   // 2 IMAD + 1 IADD3: fma-bound mix
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(z[i]), "r"(x[j]));
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(z[i]), "r"(y[j]));
                z[i] = z[i] + z[j] + x[j];
            } else if (OP == 9) {   // FFMA (fp32, both fma pipes): reference for a full-rate instruction
                float a = __uint_as_float(x[i]), b = __uint_as_float(x[j]), c = __uint_as_float(y[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
                x[i] = __float_as_uint(a);
            } else if (OP == 10) {  // 1 FFMA + 1 IADD3
                float a = __uint_as_float(x[i]), b = __uint_as_float(x[j]), c = __uint_as_float(z[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
                x[i] = __float_as_uint(a);
                y[i] = y[i] + y[j] + z[j];
            } else if (OP == 11) {  // 1 FFMA + 1 IMAD
                float a = __uint_as_float(x[i]), b = __uint_as_float(x[j]), c = __uint_as_float(z[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
                x[i] = __float_as_uint(a);
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(z[i]), "r"(y[j]));
            } else if (OP == 12) {  // IADD3 three inputs, two carry outs + IADD3.X two carry ins (x - t + c as the butterflies do)
                u64 a = ((u64)y[i] << 32) | x[i], b = ((u64)y[j] << 32) | x[j], c = ((u64)z[j] << 32) | z[i];
                a = a - b + c;
                x[i] = (u32)a; y[i] = (u32)(a >> 32);
            } else if (OP == 13) {  // DFMA: is the fp64 pipe independent of the integer issue limit?
                double a = __hiloint2double(y[i], x[i]), b = __hiloint2double(y[j], x[j]);
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a) : "d"(b));
                x[i] = __double2loint(a); y[i] = __double2hiint(a);
            } else if (OP == 14) {  // 1 DFMA + 1 IMAD
                double a = __hiloint2double(y[i], x[i]), b = __hiloint2double(y[j], x[j]);
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a) : "d"(b));
                x[i] = __double2loint(a); y[i] = __double2hiint(a);
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(z[i]) : "r"(z[j]), "r"(z[j]));
            } else if (OP == 15) {  // IMAD.HI.U32
                asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[j]), "r"(x[j]));
            } else if (OP == 16) {  // IMAD used as an adder (x*1 + y) next to IADD3: "IMAD.IADD"
                asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(x[i]) : "r"(x[j]));
            } else if (OP == 20) {  // IMAD.WIDE.U32 with RZ addend
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[i]), "r"((u32)(w[i] >> 32)));
            } else if (OP == 21) {  // IMAD (lo) with RZ addend
                x[i] = x[i] * x[j];
            } else if (OP == 22) {  // DFMA on double registers (no moves)
                asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(d[i]) : "d"(d[j]));
            } else if (OP == 23) {  // 1 IMAD.WIDE (RZ) + 1 IMAD (lo), as in the low-product chains
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[i]), "r"((u32)(w[i] >> 32)));
                asm volatile("mad.lo.u32 %0, %0, %0, %1;" : "+r"(x[i]) : "r"(z[i]));
            } else if (OP == 24) {  // 1 IMAD.WIDE (RZ) + 2 IADD3 (3-input)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[i]), "r"((u32)(w[i] >> 32)));
                y[i] = y[i] + x[i] + z[i];
                x[i] = x[i] + y[i] + z[i];
            } else if (OP == 25) {  // 1 DFMA + 1 IMAD.WIDE (RZ)
                asm volatile("fma.rn.f64 %0, %0, %0, %0;" : "+d"(d[i]));
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[i]), "r"((u32)(w[i] >> 32)));
            } else if (OP == 26) {  // IMAD.WIDE.U32 with a constant-bank multiplicand and RZ addend (one register read)
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)(w[i] >> 32)), "r"(cparam));
            } else if (OP == 27) {  // IMAD (lo) with a constant-bank multiplicand
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(x[i]) : "r"(x[j]), "r"(cparam));
            } else if (OP == 28) {  // I2F.F64.U32 (u32 -> double)
                d[i] = __uint2double_rn(x[i]); x[i] = (u32)__double2loint(d[i]) + (u32)__double2hiint(d[i]);
            } else if (OP == 29) {  // I2F.F64.U32 + IMAD.WIDE on independent registers
                d[i] = __uint2double_rn(x[i]); x[i] = (u32)__double2loint(d[i]) + (u32)__double2hiint(d[i]);
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w[i]) : "r"((u32)w[i]), "r"((u32)(w[i] >> 32)));
            } else if (OP == 17) {  // SHF.L.W (funnel shift): alu
                asm volatile("shf.l.wrap.b32 %0, %0, %1, 7;" : "+r"(x[i]) : "r"(x[j]));
            }
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc += x[i] + y[i] + z[i] + (u32)w[i] + (u32)(w[i] >> 32) + (u32)__double2loint(d[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static const char *g_only = nullptr;
template <int OP>
static void run(const char *name, int sms, int instr_per_elem, const u32 *in, bool last)
{
    if (g_only && strcmp(g_only, name) != 0) return;
    const int ctas = 4;                       // 4 x 256 threads = 8 warps per SM sub-partition
    int blocks = sms * ctas;
    u32 *out; long long *cyc;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    cudaMalloc(&cyc, blocks * 8);
    k<OP><<<blocks, 256>>>(out, in, cyc, 0x9E3779B1u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, in, cyc, 0x9E3779B1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(blocks);
    cudaMemcpy(h.data(), cyc, blocks * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (auto c : h) if (c > mx) mx = c;
    double warp_instr_per_smsp = (double)ctas * 8 / 4 * ILP * ITERS * instr_per_elem;
    (void)warp_instr_per_smsp;
    // SMSP cycles one warp-group costs (8 warps per SMSP share the issue port): divide the SASS instruction count of a group by this for IPC
    printf("  \"%s\": {\"smsp_cycles_per_warp_group\": %.3f, \"ms\": %.4f}%s\n", name, (double)mx / ((double)ctas * 8 / 4 * ILP * ITERS), ms, last ? "" : ",");
    cudaFree(out); cudaFree(cyc);
}

int main(int argc, char **argv)
{
    if (argc > 1) g_only = argv[1];
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int s = p.multiProcessorCount;
    std::vector<u32> h(12288);
    u64 x = 0x9E3779B97F4A7C15ull;
    for (auto &v : h) { x ^= x << 13; x ^= x >> 7; x ^= x << 17; v = (u32)x | 1u; }
    u32 *in; cudaMalloc(&in, h.size() * 4);
    cudaMemcpy(in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"warps_per_smsp\": 8,\n", p.name, s);
    run<0>("imad", s, 1, in, false);
    run<1>("iadd3", s, 1, in, false);
    run<2>("lop3", s, 1, in, false);
    run<3>("add64_carry_pair", s, 2, in, false);
    run<4>("imad_wide_acc", s, 1, in, false);
    run<15>("imad_hi", s, 1, in, false);
    run<16>("imad_as_add", s, 1, in, false);
    run<17>("shf", s, 1, in, false);
    run<12>("sub_add64_3input", s, 2, in, false);
    run<5>("imad+iadd3", s, 2, in, false);
    run<6>("imad+lop3", s, 2, in, false);
    run<7>("imadwide+add64", s, 3, in, false);
    run<8>("2imad+iadd3", s, 3, in, false);
    run<20>("imad_wide_rz", s, 1, in, false);
    run<21>("imad_lo_rz", s, 1, in, false);
    run<26>("imad_wide_rz_cbank", s, 1, in, false);
    run<27>("imad_lo_cbank", s, 1, in, false);
    run<23>("imadwide+imad", s, 2, in, false);
    run<24>("imadwide+2iadd3", s, 3, in, false);
    run<22>("dfma_regs", s, 1, in, false);
    run<25>("dfma+imadwide", s, 2, in, false);
    run<28>("i2f_f64_u32(+iadd3)", s, 2, in, false);
    run<29>("i2f_f64_u32(+iadd3)+imadwide", s, 3, in, false);
    run<9>("ffma", s, 1, in, false);
    run<10>("ffma+iadd3", s, 2, in, false);
    run<11>("ffma+imad", s, 2, in, false);
    run<13>("dfma", s, 1, in, false);
    run<14>("dfma+imad", s, 2, in, true);
    printf("}\n");
    return 0;
}
