// ref_dump.cu -- TEST INFRASTRUCTURE.  A small driver over the UNMODIFIED reference headers (compiled with
// -I /root/reference/BFV_Scheme by oracle/ref/Makefile into oracle/_ref/ref_dump; never linked into the product).
// It runs the reference's own kernels on the B200 for (a) GPU-vs-GPU parity dumps and (b) the "reference kernels
// rebuilt for B200" baseline timings of BASELINE.md section 2.
//
//   ref_dump ntt   <set> <num> <outdir>     forwardNTT_batch / inverseNTT_batch on `num` seeded polynomials
//   ref_dump bfv   <set> <outdir>           keygen_rns -> encryption_rns -> decryption_rns, every buffer dumped
//   ref_dump bench <set> <num> <iters>      CUDA-event timings (JSON on stdout)
//   ref_dump c1    <iters>                  single-polynomial latency, N = 4096, 58-bit prime (BASELINE config 1)
//
// Inputs are generated here with the same splitmix64 recipe as oracle/ntt_oracle.c:orc_fill_uniform, so no input files
// are needed.  Twiddle tables are produced by walking the exponents (identical values to fillTablePsi128, which the
// CPU tests check) because the reference's bit-serial table fill takes minutes at n = 32768 x 16 limbs.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
using std::vector;

#include "helper.h"
#include "parameter.h"
#include "poly_arithmetic.cuh"
#include "distributions.cuh"
#include "bfv_keygen.cuh"
#include "bfv_encryption.cuh"
#include "bfv_decryption.cuh"

typedef unsigned long long u64;

struct Set { const char *name; unsigned n; vector<u64> q, psi; };
static vector<Set> sets()
{
    return {
        {"32k_16q", 32768, {18014398506729473ull, 36028797017456641ull, 36028797014704129ull, 36028797014573057ull, 36028797014376449ull, 36028797013327873ull, 36028797013000193ull, 36028797012606977ull, 36028797010444289ull, 36028797009985537ull, 36028797005856769ull, 36028797005529089ull, 36028797005135873ull, 36028797003694081ull, 36028797003563009ull, 36028797001138177ull},
         {58232959302ull, 1155186985540ull, 631260524634ull, 1526647220035ull, 455957817523ull, 1650884166641ull, 10316746886ull, 768741990072ull, 3911086673862ull, 5947090524825ull, 47595902954ull, 2691682578057ull, 3903338373ull, 235185854118ull, 1769787302793ull, 3151164484090ull}},
        {"32k_9q", 32768, {36028797012606977ull, 36028797010444289ull, 36028797009985537ull, 36028797005856769ull, 36028797005529089ull, 36028797005135873ull, 36028797003694081ull, 36028797003563009ull, 36028797001138177ull},
         {768741990072ull, 3911086673862ull, 5947090524825ull, 47595902954ull, 2691682578057ull, 3903338373ull, 235185854118ull, 1769787302793ull, 3151164484090ull}},
        {"16k_5q", 16384, {1125899904679937ull, 1125899903991809ull, 1125899903827969ull, 1125899903795201ull, 1125899903500289ull},
         {184459094098ull, 125929543876ull, 13806300337ull, 10351677219ull, 68423600398ull}},
        {"8k_4q", 8192, {8796092858369ull, 8796092792833ull, 17592186028033ull, 17592185438209ull}, {1734247217ull, 304486499ull, 331339694ull, 9366611238ull}},
        {"8k_3q", 8192, {274877562881ull, 274877202433ull, 274877153281ull}, {71485851ull, 33872056ull, 22399294ull}},
        {"4k_3q", 4096, {68719403009ull, 68719230977ull, 137438822401ull}, {24250113ull, 29008497ull, 8625844ull}},
    };
}

static u64 splitmix64(u64 &s)
{
    u64 z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static void fill_uniform(u64 *a, size_t n, u64 q, u64 seed)
{
    int bits = 64; while (bits > 1 && !((q - 1) >> (bits - 1))) bits--;
    u64 mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1), s = seed;
    for (size_t i = 0; i < n; i++) { u64 v; do v = splitmix64(s) & mask; while (v >= q); a[i] = v; }
}
static void fast_tables(u64 psi, u64 q, u64 psiinv, u64 *t, u64 *ti, unsigned n)
{
    int bits = 0; while ((1u << bits) < n) bits++;
    u64 p = 1, pi = 1;
    for (unsigned e = 0; e < n; e++) {
        u64 i = bitReverse(e, bits);
        t[i] = p; ti[i] = pi;
        p = (u64)((unsigned __int128)p * psi % q); pi = (u64)((unsigned __int128)pi * psiinv % q);
    }
}
static void dump(const std::string &dir, const char *name, const void *dev, size_t bytes)
{
    vector<unsigned char> h(bytes);
    cudaMemcpy(h.data(), dev, bytes, cudaMemcpyDeviceToHost);
    FILE *f = fopen((dir + "/" + name).c_str(), "wb");
    if (!f) { perror("fopen"); exit(2); }
    fwrite(h.data(), 1, bytes, f);
    fclose(f);
}

struct Ring {     // everything demo.cu:62-272 sets up, for one parameter set
    unsigned n, r;
    vector<u64> q, mu_array, inv_q_last_mod_q, inv_punctured_q, prod_t_gamma_mod_q, neg_inv, output_base;
    vector<unsigned> q_bit_lengths, output_base_bit_lengths;
    u64 t = 1024, gamma = 2305843009213683713ull, gamma_div_2, mu_gamma;
    u64 *psi_dev, *psiinv_dev, *q_array_device, *qi_div_t_dev, *bcm_dev;
};
static Ring setup(const Set &S)
{
    Ring R; R.n = S.n; R.r = (unsigned)S.q.size(); R.q = S.q;
    unsigned r = R.r, n = R.n;
    if (r > 8) { fprintf(stderr, "note: %u limbs: demo.cu:72 copies 8*r bytes into the 64-byte q_bit_cons; using a 4*r-byte copy here\n", r); }
    vector<unsigned> qb(16, 0);
    for (unsigned i = 0; i < r; i++) { R.q_bit_lengths.push_back((unsigned)(log2((double)S.q[i]) + 1)); qb[i] = R.q_bit_lengths[i]; }
    cudaMemcpyToSymbol(q_bit_cons, qb.data(), sizeof(unsigned) * 16);
    vector<u64> tmp(16, 0);
    for (unsigned i = 0; i + 1 < r; i++) { R.inv_q_last_mod_q.push_back(modinv128(S.q[r - 1] % S.q[i], S.q[i])); tmp[i] = R.inv_q_last_mod_q[i]; }
    cudaMemcpyToSymbol(inv_q_last_mod_q_cons, tmp.data(), 8 * 16);
    vector<u64> qdt(r);
    for (unsigned i = 0; i < r; i++) qdt[i] = S.q[i] / R.t;
    R.gamma_div_2 = R.gamma >> 1;
    R.output_base = {R.t, R.gamma}; R.output_base_bit_lengths = {10, 61};
    u64 mult_t = 1, mult_g = 1;
    for (unsigned i = 0; i + 1 < r; i++) { mult_t = (host64x2(mult_t, S.q[i]) % R.t).low; mult_g = (host64x2(mult_g, S.q[i]) % R.gamma).low; }
    R.neg_inv = {R.t - modinv128(mult_t, R.t), R.gamma - modinv128(mult_g, R.gamma)};
    uint128_t ptg = host64x2(R.t, R.gamma);
    std::fill(tmp.begin(), tmp.end(), 0);
    for (unsigned i = 0; i + 1 < r; i++) { R.prod_t_gamma_mod_q.push_back((ptg % S.q[i]).low); tmp[i] = R.prod_t_gamma_mod_q[i]; }
    cudaMemcpyToSymbol(prod_t_gamma_mod_q_cons, tmp.data(), 8 * 16);
    std::fill(tmp.begin(), tmp.end(), 0);
    for (unsigned i = 0; i < r; i++) tmp[i] = S.q[i];
    cudaMemcpyToSymbol(q_cons, tmp.data(), 8 * 16);
    cudaMalloc(&R.q_array_device, 8 * r); cudaMemcpy(R.q_array_device, S.q.data(), 8 * r, cudaMemcpyHostToDevice);
    cudaMalloc(&R.qi_div_t_dev, 8 * r); cudaMemcpy(R.qi_div_t_dev, qdt.data(), 8 * r, cudaMemcpyHostToDevice);
    std::fill(tmp.begin(), tmp.end(), 0);
    for (unsigned i = 0; i < r; i++) {
        uint128_t mu1 = uint128_t::exp2(2 * R.q_bit_lengths[i]);
        R.mu_array.push_back((mu1 / S.q[i]).low); tmp[i] = R.mu_array[i];
    }
    cudaMemcpyToSymbol(mu_cons, tmp.data(), 8 * 16);
    vector<u64> psi((size_t)r * n), psiinv((size_t)r * n);
    for (unsigned i = 0; i < r; i++) fast_tables(S.psi[i], S.q[i], modinv128(S.psi[i], S.q[i]), &psi[(size_t)i * n], &psiinv[(size_t)i * n], n);
    cudaMalloc(&R.psi_dev, 8 * (size_t)r * n); cudaMalloc(&R.psiinv_dev, 8 * (size_t)r * n);
    cudaMemcpy(R.psi_dev, psi.data(), 8 * (size_t)r * n, cudaMemcpyHostToDevice);
    cudaMemcpy(R.psiinv_dev, psiinv.data(), 8 * (size_t)r * n, cudaMemcpyHostToDevice);
    unsigned rp = r - 1;
    { uint128_t mu1 = uint128_t::exp2(2 * 61); R.mu_gamma = (mu1 / R.gamma).low; }
    std::fill(tmp.begin(), tmp.end(), 0);
    for (unsigned i = 0; i < rp; i++) {
        uint128_t t1 = 1;
        for (unsigned j = 0; j < rp; j++) if (i != j) t1 = host64x2(t1.low, S.q[j]) % S.q[i];
        R.inv_punctured_q.push_back(modinv128(t1.low, S.q[i])); tmp[i] = R.inv_punctured_q[i];
    }
    cudaMemcpyToSymbol(inv_punctured_q_cons, tmp.data(), 8 * 16);
    vector<u64> bcm(2 * rp);
    for (int i = 0; i < 2; i++)
        for (unsigned j = 0; j < rp; j++) {
            uint128_t t1 = 1;
            for (unsigned k = 0; k < rp; k++) if (j != k) t1 = host64x2(t1.low, S.q[k]) % R.output_base[i];
            bcm[i * rp + j] = t1.low;
        }
    cudaMalloc(&R.bcm_dev, 8 * 2 * rp); cudaMemcpy(R.bcm_dev, bcm.data(), 8 * 2 * rp, cudaMemcpyHostToDevice);
    return R;
}

int main(int argc, char **argv)
{
    if (argc < 3) { fprintf(stderr, "usage: see source header\n"); return 2; }
    std::string mode = argv[1], setname = argv[2];
    if (mode == "c1") {
        // BASELINE config 1: ONE polynomial, N = 4096, the 58-bit prime of parameter.h:43-47 -- latency of the reference's single
        // forwardNTT / inverseNTT calls (ntt_60bit.cuh:314, :350), CUDA events over a back-to-back loop after warm-up, plus the
        // forwardNTT -> barrett -> inverseNTT flow of 60bit_ntt_test.cu:70-80.
        const unsigned n = 4096; const int iters = atoi(argv[2]) > 0 ? atoi(argv[2]) : 200;
        const u64 q = 288230376135196673ull, psi = 60193018759093ull, psiinv = 236271020333049746ull;
        const int qbit = 58;
        const u64 mu = (uint128_t::exp2(2 * qbit) / q).low;
        vector<u64> t(n), ti(n), a(n);
        fast_tables(psi, q, psiinv, t.data(), ti.data(), n);
        fill_uniform(a.data(), n, q, 0xC1);
        u64 *d, *d2, *td, *tid;
        cudaMalloc(&d, 8 * n); cudaMalloc(&d2, 8 * n); cudaMalloc(&td, 8 * n); cudaMalloc(&tid, 8 * n);
        cudaMemcpy(d, a.data(), 8 * n, cudaMemcpyHostToDevice); cudaMemcpy(d2, a.data(), 8 * n, cudaMemcpyHostToDevice);
        cudaMemcpy(td, t.data(), 8 * n, cudaMemcpyHostToDevice); cudaMemcpy(tid, ti.data(), 8 * n, cudaMemcpyHostToDevice);
        cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float f_ms, i_ms, mul_ms;
        for (int i = 0; i < 10; i++) { forwardNTT(d, n, s1, q, mu, qbit, td); inverseNTT(d, n, s1, q, mu, qbit, tid); }
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1); for (int i = 0; i < iters; i++) forwardNTT(d, n, s1, q, mu, qbit, td); cudaEventRecord(e1, s1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&f_ms, e0, e1);
        cudaEventRecord(e0, s1); for (int i = 0; i < iters; i++) inverseNTT(d, n, s1, q, mu, qbit, tid); cudaEventRecord(e1, s1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&i_ms, e0, e1);
        cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        for (int i = 0; i < iters; i++) {
            forwardNTT(d, n, s1, q, mu, qbit, td); forwardNTT(d2, n, s1, q, mu, qbit, td);
            barrett<<<n / 256, 256, 0, s1>>>(d, d2, q, mu, qbit);
            inverseNTT(d, n, s1, q, mu, qbit, tid);
        }
        cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&mul_ms, e0, e1);
        vector<u64> back(n);
        cudaMemcpy(d, a.data(), 8 * n, cudaMemcpyHostToDevice);
        forwardNTT(d, n, s1, q, mu, qbit, td); inverseNTT(d, n, s1, q, mu, qbit, tid);
        cudaDeviceSynchronize();
        cudaMemcpy(back.data(), d, 8 * n, cudaMemcpyDeviceToHost);
        printf("{\"c1\": true, \"n\": %u, \"iters\": %d, \"fwd_us\": %.3f, \"inv_us\": %.3f, \"polymul_us\": %.3f, \"roundtrip_ok\": %s, \"err\": \"%s\"}\n",
               n, iters, 1e3 * f_ms / iters, 1e3 * i_ms / iters, 1e3 * mul_ms / iters, back == a ? "true" : "false", cudaGetErrorString(cudaGetLastError()));
        return 0;
    }
    const Set *S = nullptr;
    static vector<Set> all = sets();
    for (auto &s : all) if (setname == s.name) S = &s;
    if (!S) { fprintf(stderr, "unknown set %s\n", setname.c_str()); return 2; }
    Ring R = setup(*S);
    unsigned n = R.n, r = R.r;
    if (mode == "ntt") {
        unsigned num = atoi(argv[3]); std::string dir = argv[4];
        vector<u64> a((size_t)num * n);
        for (unsigned p = 0; p < num; p++) fill_uniform(&a[(size_t)p * n], n, S->q[p % r], 0x5EED0000ull + p);
        u64 *d; cudaMalloc(&d, 8 * a.size());
        cudaMemcpy(d, a.data(), 8 * a.size(), cudaMemcpyHostToDevice);
        forwardNTT_batch(d, n, R.psi_dev, num, r);
        dump(dir, "ref_fwd.bin", d, 8 * a.size());
        inverseNTT_batch(d, n, R.psiinv_dev, num, r);
        dump(dir, "ref_inv.bin", d, 8 * a.size());
        printf("ntt %s num=%u ok\n", S->name, num);
    } else if (mode == "bfv") {
        std::string dir = argv[3];
        size_t rn = (size_t)r * n;
        unsigned char *in; cudaMalloc(&in, 9 * rn + 4 * n);
        u64 *sk, *pk, *temp, *c, *e, *m_dev;
        cudaMalloc(&sk, 8 * rn); cudaMalloc(&pk, 16 * rn); cudaMalloc(&temp, 8 * rn); cudaMalloc(&c, 16 * rn); cudaMalloc(&e, 16 * rn);
        cudaMemset(c, 0, 16 * rn);
        vector<u64> m(n); fill_uniform(m.data(), n, R.t, 0xC0FFEE);
        cudaMalloc(&m_dev, 8 * n); cudaMemcpy(m_dev, m.data(), 8 * n, cudaMemcpyHostToDevice);
        cudaStream_t *streams = (cudaStream_t *)malloc(sizeof(cudaStream_t) * r * 2);
        for (unsigned i = 0; i < r * 2; i++) cudaStreamCreate(&streams[i]);
        u64 **u = (u64 **)malloc(sizeof(u64 *) * r);
        keygen_rns(in, r, (u64 *)S->q.data(), n, sk, pk, streams, temp, R.mu_array, R.q_bit_lengths, R.psi_dev, R.psiinv_dev);
        cudaDeviceSynchronize();
        dump(dir, "ref_keygen_in.bin", in, 9 * rn + 4 * n); dump(dir, "ref_sk.bin", sk, 8 * rn); dump(dir, "ref_pk.bin", pk, 16 * rn);
        dump(dir, "ref_temp.bin", temp, 8 * rn);
        encryption_rns(c, pk, in, u, e, n, streams, (u64 *)S->q.data(), R.q_bit_lengths, R.mu_array, R.inv_q_last_mod_q, R.psi_dev, R.psiinv_dev, m_dev,
                       R.qi_div_t_dev, R.q_array_device, (unsigned)R.t, r);
        cudaDeviceSynchronize();
        dump(dir, "ref_c.bin", c, 16 * rn); dump(dir, "ref_e.bin", e, 16 * rn);
        decryption_rns(c, sk, (u64 *)S->q.data(), R.q_bit_lengths, R.mu_array, R.psi_dev, R.psiinv_dev, n, r - 1, R.inv_punctured_q, R.bcm_dev, R.t,
                       R.gamma, R.mu_gamma, R.output_base, R.output_base_bit_lengths, R.neg_inv, R.gamma_div_2, R.prod_t_gamma_mod_q);
        cudaDeviceSynchronize();
        dump(dir, "ref_plain.bin", c + (size_t)n * (r - 2), 8 * n);
        vector<u64> plain(n); cudaMemcpy(plain.data(), c + (size_t)n * (r - 2), 8 * n, cudaMemcpyDeviceToHost);
        printf("bfv %s roundtrip %s err=%s\n", S->name, plain == m ? "ok" : "MISMATCH", cudaGetErrorString(cudaGetLastError()));
    } else if (mode == "bench") {
        unsigned num = atoi(argv[3]); int iters = atoi(argv[4]);
        vector<u64> a((size_t)num * n);
        for (unsigned p = 0; p < num; p++) fill_uniform(&a[(size_t)p * n], n, S->q[p % r], 0x5EED0000ull + p % 64);
        u64 *d; cudaMalloc(&d, 8 * a.size());
        cudaMemcpy(d, a.data(), 8 * a.size(), cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float fwd_ms, inv_ms;
        for (int i = 0; i < 3; i++) forwardNTT_batch(d, n, R.psi_dev, num, r);
        cudaEventRecord(e0); for (int i = 0; i < iters; i++) forwardNTT_batch(d, n, R.psi_dev, num, r); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&fwd_ms, e0, e1);
        for (int i = 0; i < 3; i++) inverseNTT_batch(d, n, R.psiinv_dev, num, r);
        cudaEventRecord(e0); for (int i = 0; i < iters; i++) inverseNTT_batch(d, n, R.psiinv_dev, num, r); cudaEventRecord(e1);
        cudaEventSynchronize(e1); cudaEventElapsedTime(&inv_ms, e0, e1);
        // BFV single-item latencies, looped (the reference has no batch API)
        size_t rn = (size_t)r * n;
        unsigned char *in; cudaMalloc(&in, 9 * rn + 4 * n);
        u64 *sk, *pk, *temp, *c, *e, *m_dev, *c_keep;
        cudaMalloc(&sk, 8 * rn); cudaMalloc(&pk, 16 * rn); cudaMalloc(&temp, 8 * rn); cudaMalloc(&c, 16 * rn); cudaMalloc(&e, 16 * rn); cudaMalloc(&c_keep, 16 * rn);
        vector<u64> m(n); fill_uniform(m.data(), n, R.t, 0xC0FFEE);
        cudaMalloc(&m_dev, 8 * n); cudaMemcpy(m_dev, m.data(), 8 * n, cudaMemcpyHostToDevice);
        cudaStream_t *streams = (cudaStream_t *)malloc(sizeof(cudaStream_t) * r * 2);
        for (unsigned i = 0; i < r * 2; i++) cudaStreamCreate(&streams[i]);
        u64 **u = (u64 **)malloc(sizeof(u64 *) * r);
        float kg_ms, enc_ms, dec_ms;
        auto KG = [&] { keygen_rns(in, r, (u64 *)S->q.data(), n, sk, pk, streams, temp, R.mu_array, R.q_bit_lengths, R.psi_dev, R.psiinv_dev); };
        auto ENC = [&] { encryption_rns(c, pk, in, u, e, n, streams, (u64 *)S->q.data(), R.q_bit_lengths, R.mu_array, R.inv_q_last_mod_q, R.psi_dev, R.psiinv_dev, m_dev, R.qi_div_t_dev, R.q_array_device, (unsigned)R.t, r); };
        for (int i = 0; i < 3; i++) KG();
        cudaDeviceSynchronize();
        cudaEventRecord(e0); for (int i = 0; i < iters; i++) KG(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&kg_ms, e0, e1);
        for (int i = 0; i < 3; i++) ENC();
        cudaDeviceSynchronize();
        cudaEventRecord(e0); for (int i = 0; i < iters; i++) ENC(); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&enc_ms, e0, e1);
        cudaMemcpy(c_keep, c, 16 * rn, cudaMemcpyDeviceToDevice);
        // decryption_rns creates r-1 streams per call and never destroys them: keep the loop short
        int dit = iters < 20 ? iters : 20;
        double dec_total = 0;
        for (int i = 0; i < dit + 2; i++) {
            cudaMemcpy(c, c_keep, 16 * rn, cudaMemcpyDeviceToDevice);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            decryption_rns(c, sk, (u64 *)S->q.data(), R.q_bit_lengths, R.mu_array, R.psi_dev, R.psiinv_dev, n, r - 1, R.inv_punctured_q, R.bcm_dev, R.t,
                           R.gamma, R.mu_gamma, R.output_base, R.output_base_bit_lengths, R.neg_inv, R.gamma_div_2, R.prod_t_gamma_mod_q);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (i >= 2) dec_total += ms;
        }
        dec_ms = (float)dec_total;
        printf("{\"set\": \"%s\", \"n\": %u, \"limbs\": %u, \"num\": %u, \"iters\": %d, \"fwd_ms_per_batch\": %.5f, \"inv_ms_per_batch\": %.5f, "
               "\"fwd_ntt_per_s\": %.1f, \"inv_ntt_per_s\": %.1f, \"keygen_us\": %.2f, \"encrypt_us\": %.2f, \"decrypt_us\": %.2f, \"err\": \"%s\"}\n",
               S->name, n, r, num, iters, fwd_ms / iters, inv_ms / iters, num * 1e3 * iters / fwd_ms, num * 1e3 * iters / inv_ms, 1e3 * kg_ms / iters,
               1e3 * enc_ms / iters, 1e3 * dec_ms / dit, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
