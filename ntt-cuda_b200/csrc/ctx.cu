// ctx.cu -- contexts: reference-layout twiddle tables, Shoup companions and per-limb constants in HBM, plus the
// host-buffer (end-to-end) transform pipeline.  Replaces the parameter/table set-up every reference driver
// repeats by hand (demo.cu:62-196: q_bit/mu computation, fillTablePsi128 per limb, cudaMemcpy per limb,
// cudaMemcpyToSymbol into six 16-entry __constant__ tables).
#include "internal.h"
#include "table_kernels.cuh"
#include "bfv_kernels.cuh"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using namespace nttb200;
typedef unsigned __int128 u128;

static u64 h_modpow(u64 a, u64 e, u64 m)
{
    u64 r = 1 % m;
    a %= m;
    while (e) { if (e & 1) r = (u64)((u128)r * a % m); a = (u64)((u128)a * a % m); e >>= 1; }
    return r;
}
static u64 h_shoup(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }

// Allocates the device state, fills the per-limb constants and produces the tables: from the roots on the device
// (psi_h == nullptr) or from caller-supplied host tables (companions still computed on the device).
static int ctx_finish(nttb200_ctx *c, const u64 *roots, const u64 *psi_h, const u64 *psiinv_h)
{
    const size_t tot = (size_t)c->limbs * c->n;
    std::vector<LimbConst> lc(c->limbs);
    std::vector<u64> roots_v(c->limbs, 0), rootsinv_v(c->limbs, 0);
    c->mu.resize(c->limbs); c->qbit.resize(c->limbs);
    for (unsigned l = 0; l < c->limbs; l++) {
        const u64 q = c->q[l];
        const unsigned qbit = (unsigned)(log2((double)q) + 1);            // demo.cu:69
        const u64 mu = 2 * qbit < 128 ? (u64)(((u128)1 << (2 * qbit)) / q) : 0;   // demo.cu:157-165
        c->mu[l] = mu; c->qbit[l] = qbit;
        const u64 ninv = h_modpow(c->n % q, q - 2, q);
        u64 w1;                                                          // psiinv[1] = psiinv^(n/2)
        if (roots) {
            roots_v[l] = roots[l] % q;
            rootsinv_v[l] = h_modpow(roots_v[l], q - 2, q);              // demo.cu:96-97
            w1 = h_modpow(rootsinv_v[l], c->n / 2, q);
        } else {
            w1 = psiinv_h[l * (size_t)c->n + 1];
        }
        const u64 w1n = (u64)((u128)w1 * ninv % q);
        LimbConst &k = lc[l];
        k.q = q; k.twoq = 2 * q; k.mu = mu; k.qbit = qbit; k.pad = 0;
        k.ratio = (u64)((((u128)1) << 64) / q); k.negq = 0 - q;
        if (qbit > 57) c->lazy_ok = 0;   // 69 q < 2^64
        k.ninv = ninv; k.ninv_s = h_shoup(ninv, q);
        k.w1ninv = w1n; k.w1ninv_s = h_shoup(w1n, q);
    }
    NTTB200_CHECK(cudaGetDevice(&c->device));
    NTTB200_CHECK(cudaMalloc(&c->psi, tot * 8));
    NTTB200_CHECK(cudaMalloc(&c->psiinv, tot * 8));
    NTTB200_CHECK(cudaMalloc(&c->psi_s, tot * 8));
    NTTB200_CHECK(cudaMalloc(&c->psiinv_s, tot * 8));
    NTTB200_CHECK(cudaMalloc(&c->lc, sizeof(LimbConst) * c->limbs));
    NTTB200_CHECK(cudaMalloc(&c->q_dev, 8 * c->limbs));
    NTTB200_CHECK(cudaMalloc(&c->mu_dev, 8 * c->limbs));
    NTTB200_CHECK(cudaMalloc(&c->qbit_dev, 4 * c->limbs));
    NTTB200_CHECK(cudaMemcpy(c->lc, lc.data(), sizeof(LimbConst) * c->limbs, cudaMemcpyHostToDevice));
    NTTB200_CHECK(cudaMemcpy(c->q_dev, c->q.data(), 8 * c->limbs, cudaMemcpyHostToDevice));
    NTTB200_CHECK(cudaMemcpy(c->mu_dev, c->mu.data(), 8 * c->limbs, cudaMemcpyHostToDevice));
    NTTB200_CHECK(cudaMemcpy(c->qbit_dev, c->qbit.data(), 4 * c->limbs, cudaMemcpyHostToDevice));
    if (roots) {
        u64 *rd = nullptr;                                               // [roots | inverse roots]
        NTTB200_CHECK(cudaMalloc(&rd, 16 * c->limbs));
        NTTB200_CHECK(cudaMemcpy(rd, roots_v.data(), 8 * c->limbs, cudaMemcpyHostToDevice));
        NTTB200_CHECK(cudaMemcpy(rd + c->limbs, rootsinv_v.data(), 8 * c->limbs, cudaMemcpyHostToDevice));
        k_build_tables<<<dim3(c->limbs, 2), 1024>>>(c->psi, c->psi_s, c->psiinv, c->psiinv_s, c->q_dev, rd, rd + c->limbs, c->logn);
        cudaError_t e = cudaDeviceSynchronize();
        cudaFree(rd);
        if (e != cudaSuccess) return (int)e;
    } else {
        NTTB200_CHECK(cudaMemcpy(c->psi, psi_h, tot * 8, cudaMemcpyHostToDevice));
        NTTB200_CHECK(cudaMemcpy(c->psiinv, psiinv_h, tot * 8, cudaMemcpyHostToDevice));
        k_build_companions<<<1184, 256>>>(c->psi, c->psi_s, c->q_dev, c->logn, c->limbs, c->limbs);
        k_build_companions<<<1184, 256>>>(c->psiinv, c->psiinv_s, c->q_dev, c->logn, c->limbs, c->limbs);
        NTTB200_CHECK(cudaDeviceSynchronize());
    }
#ifdef NTT_TWZ   /* experiment: the companion tables become interleaved {w, companion} tables (only the plain transform entries are valid in this build) */
    for (int which = 0; which < 2; which++) {
        u64 **sp = which ? &c->psiinv_s : &c->psi_s;
        const u64 *wsrc = which ? c->psiinv : c->psi;
        std::vector<u64> hw(tot), hs(tot), hz(2 * tot);
        NTTB200_CHECK(cudaMemcpy(hw.data(), wsrc, tot * 8, cudaMemcpyDeviceToHost));
        NTTB200_CHECK(cudaMemcpy(hs.data(), *sp, tot * 8, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < tot; i++) { hz[2 * i] = hw[i]; hz[2 * i + 1] = hs[i]; }
        NTTB200_CHECK(cudaFree(*sp));
        NTTB200_CHECK(cudaMalloc(sp, 2 * tot * 8));
        NTTB200_CHECK(cudaMemcpy(*sp, hz.data(), 2 * tot * 8, cudaMemcpyHostToDevice));
    }
#endif
    return 0;
}

static int ctx_alloc(nttb200_ctx **out, unsigned n, unsigned limbs, const u64 *q)
{
    if (!out || !q || limbs == 0 || limbs > NTTB200_MAX_LIMBS || (n & (n - 1)) || n < 2048 || n > 131072) return NTTB200_EINVAL;
    for (unsigned l = 0; l < limbs; l++)
        if (q[l] < 3 || (q[l] >> 62) || ((q[l] - 1) % (2ull * n))) return NTTB200_EINVAL;   // q < 2^62, q = 1 mod 2n
    nttb200_ctx *c = new nttb200_ctx();
    c->n = n; c->limbs = limbs;
    while ((1u << c->logn) < n) c->logn++;
    c->q.assign(q, q + limbs);
    c->use_tma = get_tma_default();
    *out = c;
    return 0;
}

int nttb200_trace_error(int code, const char *file, int line)
{
    static int on = -1;
    if (on < 0) { const char *e = getenv("NTTB200_DEBUG"); on = (e && e[0] == '1') ? 1 : 0; }
    if (on) fprintf(stderr, "nttb200: error %d (%s) at %s:%d\n", code, code < 10000 ? cudaGetErrorString((cudaError_t)code) : "nttb200", file, line);
    return code;
}

extern "C" {

int nttb200_version(void) { return 200; }

const char *nttb200_error_string(int code)
{
    if (code == 0) return "success";
    if (code == NTTB200_EINVAL) return "nttb200: invalid argument (unsupported n / limbs / modulus, or null pointer)";
    if (code == NTTB200_ENOTMA) return "nttb200: cuTensorMapEncodeTiled unavailable or failed";
    if (code == NTTB200_ENCCL) return "nttb200: NCCL unavailable (libnccl.so.2 not loadable) or a collective failed (NTTB200_DEBUG=1 prints the NCCL error)";
    return cudaGetErrorString((cudaError_t)code);
}

int nttb200_ctx_create(nttb200_ctx **ctx, unsigned n, unsigned limbs, const nttb200_u64 *q, const nttb200_u64 *psi_roots)
{
    if (!psi_roots) return NTTB200_EINVAL;
    int r = ctx_alloc(ctx, n, limbs, q);
    if (r) return r;
    nttb200_ctx *c = *ctx;
    for (unsigned l = 0; l < limbs; l++)                               // psi must be a primitive 2n-th root: psi^n = -1
        if (h_modpow(psi_roots[l] % q[l], n, q[l]) != q[l] - 1) { delete c; *ctx = nullptr; return NTTB200_EINVAL; }
    r = ctx_finish(c, psi_roots, nullptr, nullptr);                     // tables + companions generated on the device
    if (r) { nttb200_ctx_destroy(c); *ctx = nullptr; }
    return r;
}

int nttb200_ctx_create_from_tables(nttb200_ctx **ctx, unsigned n, unsigned limbs, const nttb200_u64 *q,
                                   const nttb200_u64 *psi_tables_host, const nttb200_u64 *psiinv_tables_host)
{
    if (!psi_tables_host || !psiinv_tables_host) return NTTB200_EINVAL;
    int r = ctx_alloc(ctx, n, limbs, q);
    if (r) return r;
    r = ctx_finish(*ctx, nullptr, psi_tables_host, psiinv_tables_host);
    if (r) { nttb200_ctx_destroy(*ctx); *ctx = nullptr; }
    return r;
}

void nttb200_ctx_destroy(nttb200_ctx *c)
{
    if (!c) return;
    cudaFree(c->psi); cudaFree(c->psiinv); cudaFree(c->psi_s); cudaFree(c->psiinv_s);
    cudaFree(c->lc); cudaFree(c->q_dev); cudaFree(c->mu_dev); cudaFree(c->qbit_dev);
    cudaFree(c->word_off_dev);
    for (int i = 0; i < nttb200_ctx::kStages; i++) {
        if (c->stage_packed[i]) cudaFree(c->stage_packed[i]);
        if (c->stage_dev[i]) cudaFree(c->stage_dev[i]);
        if (c->streams[i]) cudaStreamDestroy(c->streams[i]);
    }
    delete c;
}

int nttb200_ctx_tables(const nttb200_ctx *c, const nttb200_u64 **psi, const nttb200_u64 **psiinv)
{
    if (!c) return NTTB200_EINVAL;
    if (psi) *psi = c->psi;
    if (psiinv) *psiinv = c->psiinv;
    return 0;
}
int nttb200_ctx_consts(const nttb200_ctx *c, const nttb200_u64 **q, const nttb200_u64 **mu, const unsigned **qbit)
{
    if (!c) return NTTB200_EINVAL;
    if (q) *q = c->q_dev;
    if (mu) *mu = c->mu_dev;
    if (qbit) *qbit = c->qbit_dev;
    return 0;
}
int nttb200_download(void *dst_host, const void *src_dev, size_t bytes)
{
    NTTB200_CHECK(cudaMemcpy(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost));
    return 0;
}
int nttb200_upload(void *dst_dev, const void *src_host, size_t bytes)
{
    NTTB200_CHECK(cudaMemcpy(dst_dev, src_host, bytes, cudaMemcpyHostToDevice));
    return 0;
}
int nttb200_ctx_set_tma(nttb200_ctx *c, int enable)
{
    if (!c) return NTTB200_EINVAL;
    c->use_tma = enable ? 1 : 0;
    return 0;
}

// ---- host-buffer transforms: H2D / kernels / D2H of successive chunks overlap on kStages streams ------------------
static int host_pipeline(nttb200_ctx *c, bool inverse, const u64 *in, u64 *out, unsigned num, unsigned division)
{
    if (!c || !in || !out || division == 0 || division > c->limbs) return NTTB200_EINVAL;
    if (num == 0) return 0;
    // chunk = a multiple of `division` polynomials, about 32 MiB (NTTB200_E2E_CHUNK_MB overrides, for tuning)
    static size_t chunk_mb = 0;
    if (!chunk_mb) { const char *e = getenv("NTTB200_E2E_CHUNK_MB"); chunk_mb = e && atoi(e) > 0 ? (size_t)atoi(e) : 32; }
    size_t per = (chunk_mb << 20) / ((size_t)c->n * 8);
    per = per / division * division;
    if (per == 0) per = division;
    const size_t bytes = per * c->n * 8;
    if (c->stage_bytes < bytes) {
        for (int i = 0; i < nttb200_ctx::kStages; i++) {
            if (c->stage_dev[i]) { cudaFree(c->stage_dev[i]); c->stage_dev[i] = nullptr; }
            NTTB200_CHECK(cudaMalloc(&c->stage_dev[i], bytes));
            if (!c->streams[i]) NTTB200_CHECK(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
        }
        c->stage_bytes = bytes;
    }
    // chunk schedule: full chunks in the middle, a geometric ramp (per/4, per/2) at both ends so that the first H2D and the last
    // D2H -- the only copies with nothing to overlap -- are short
    std::vector<size_t> sizes;
    {
        size_t left = num;
        auto rounded = [&](size_t v) { v = v / division * division; return v ? v : (size_t)division; };
        const size_t ramp[2] = {rounded(per / 4), rounded(per / 2)};
        std::vector<size_t> head, tail;
        for (int i = 0; i < 2 && left > 2 * per; i++) { head.push_back(ramp[i]); tail.push_back(ramp[i]); left -= 2 * ramp[i]; }
        sizes = head;
        while (left) { const size_t cnt = left < per ? left : per; sizes.push_back(cnt); left -= cnt; }
        for (size_t i = tail.size(); i-- > 0;) sizes.push_back(tail[i]);
    }
    int k = 0;
    size_t p0 = 0;
    for (size_t ci = 0; ci < sizes.size(); p0 += sizes[ci], ci++, k = (k + 1) % nttb200_ctx::kStages) {
        const unsigned cnt = (unsigned)sizes[ci];
        cudaStream_t st = c->streams[k];
        u64 *d = c->stage_dev[k];
        NTTB200_CHECK(cudaMemcpyAsync(d, in + p0 * c->n, (size_t)cnt * c->n * 8, cudaMemcpyHostToDevice, st));
        int r = inverse ? nttb200_inverse_ntt_batch(c, d, cnt, division, st) : nttb200_forward_ntt_batch(c, d, cnt, division, st);
        if (r) return r;
        NTTB200_CHECK(cudaMemcpyAsync(out + p0 * c->n, d, (size_t)cnt * c->n * 8, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < nttb200_ctx::kStages; i++) NTTB200_CHECK(cudaStreamSynchronize(c->streams[i]));
    return 0;
}

// ---- packed host / wire format ---------------------------------------------------------------------------------------------------
static int packed_setup(nttb200_ctx *c)
{
    if (c->word_off_dev) return 0;
    c->word_off.assign(c->limbs + 1, 0);
    for (unsigned l = 0; l < c->limbs; l++) {
        const unsigned long long w = (unsigned long long)c->word_off[l] + (unsigned long long)(c->n / 64) * c->qbit[l];
        if (w >> 32) return NTTB200_EINVAL;
        c->word_off[l + 1] = (unsigned)w;
    }
    NTTB200_CHECK(cudaMalloc(&c->word_off_dev, 4 * (c->limbs + 1)));
    NTTB200_CHECK(cudaMemcpy(c->word_off_dev, c->word_off.data(), 4 * (c->limbs + 1), cudaMemcpyHostToDevice));
    return 0;
}
static int packed_convert(nttb200_ctx *c, bool pack, u64 *packed, u64 *a, unsigned num, unsigned division, cudaStream_t st)
{
    const unsigned groups = num / division;
    const unsigned x = (c->n / 64 + 127) / 128;
    for (unsigned g0 = 0; g0 < groups; g0 += 65535) {                 // grid.z limit
        const unsigned gz = groups - g0 < 65535 ? groups - g0 : 65535;
        u64 *pk = packed + (size_t)g0 * c->word_off[division], *aa = a + (size_t)g0 * division * c->n;
        if (pack) nttb200::k_ct_pack<<<dim3(x, division, gz), 128, 0, st>>>(aa, pk, c->n, division, gz, c->qbit_dev, c->word_off_dev, c->word_off[division]);
        else nttb200::k_ct_unpack<<<dim3(x, division, gz), 128, 0, st>>>(pk, aa, c->n, division, gz, c->qbit_dev, c->word_off_dev, c->word_off[division]);
    }
    NTTB200_CHECK(cudaGetLastError());
    return 0;
}
static int packed_args(nttb200_ctx *c, unsigned num, unsigned division)
{
    if (!c || division == 0 || division > c->limbs || num % division) return NTTB200_EINVAL;
    return packed_setup(c);
}
int nttb200_polys_packed_words(const nttb200_ctx *c, unsigned num, unsigned division, size_t *words)
{
    if (!c || !words || division == 0 || division > c->limbs || num % division) return NTTB200_EINVAL;
    size_t w = 0;
    for (unsigned l = 0; l < division; l++) w += (size_t)(c->n / 64) * c->qbit[l];
    *words = w * (num / division);
    return 0;
}
int nttb200_pack_polys(nttb200_ctx *c, nttb200_u64 *packed, const nttb200_u64 *a, unsigned num, unsigned division, void *stream)
{
    if (!packed || !a) return NTTB200_EINVAL;
    int r = packed_args(c, num, division);
    return r ? r : packed_convert(c, true, packed, const_cast<u64 *>(a), num, division, (cudaStream_t)stream);
}
int nttb200_unpack_polys(nttb200_ctx *c, nttb200_u64 *a, const nttb200_u64 *packed, unsigned num, unsigned division, void *stream)
{
    if (!packed || !a) return NTTB200_EINVAL;
    int r = packed_args(c, num, division);
    return r ? r : packed_convert(c, false, const_cast<u64 *>(packed), a, num, division, (cudaStream_t)stream);
}
// host_pipeline with both host arrays in the packed format: per chunk packed H2D -> unpack -> transform -> pack -> packed D2H
static int host_pipeline_packed(nttb200_ctx *c, bool inverse, const u64 *in, u64 *out, unsigned num, unsigned division)
{
    if (!in || !out) return NTTB200_EINVAL;
    int r = packed_args(c, num, division);
    if (r) return r;
    if (num == 0) return 0;
    static size_t chunk_mb = 0;
    if (!chunk_mb) { const char *e = getenv("NTTB200_E2E_CHUNK_MB"); chunk_mb = e && atoi(e) > 0 ? (size_t)atoi(e) : 32; }
    size_t per = (chunk_mb << 20) / ((size_t)c->n * 8) / division * division;          // polynomials per chunk, whole groups
    if (per == 0) per = division;
    const size_t gw = c->word_off[division];                                            // packed words per group
    const size_t bytes = per * c->n * 8, pbytes = per / division * gw * 8;
    if (c->stage_bytes < bytes || c->stage_packed_bytes < pbytes) {
        for (int i = 0; i < nttb200_ctx::kStages; i++) {
            if (c->stage_bytes < bytes) {
                if (c->stage_dev[i]) { cudaFree(c->stage_dev[i]); c->stage_dev[i] = nullptr; }
                NTTB200_CHECK(cudaMalloc(&c->stage_dev[i], bytes));
            }
            if (c->stage_packed[i]) { cudaFree(c->stage_packed[i]); c->stage_packed[i] = nullptr; }
            NTTB200_CHECK(cudaMalloc(&c->stage_packed[i], pbytes));
            if (!c->streams[i]) NTTB200_CHECK(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
        }
        if (c->stage_bytes < bytes) c->stage_bytes = bytes;
        c->stage_packed_bytes = pbytes;
    }
    int k = 0;
    for (size_t p0 = 0; p0 < num; p0 += per, k = (k + 1) % nttb200_ctx::kStages) {
        const unsigned cnt = (unsigned)(num - p0 < per ? num - p0 : per);
        const size_t w0 = p0 / division * gw, w = (size_t)cnt / division * gw;
        cudaStream_t st = c->streams[k];
        NTTB200_CHECK(cudaMemcpyAsync(c->stage_packed[k], in + w0, w * 8, cudaMemcpyHostToDevice, st));
        r = packed_convert(c, false, c->stage_packed[k], c->stage_dev[k], cnt, division, st);
        if (!r) r = inverse ? nttb200_inverse_ntt_batch(c, c->stage_dev[k], cnt, division, st) : nttb200_forward_ntt_batch(c, c->stage_dev[k], cnt, division, st);
        if (!r) r = packed_convert(c, true, c->stage_packed[k], c->stage_dev[k], cnt, division, st);
        if (r) return r;
        NTTB200_CHECK(cudaMemcpyAsync(out + w0, c->stage_packed[k], w * 8, cudaMemcpyDeviceToHost, st));
    }
    for (int i = 0; i < nttb200_ctx::kStages; i++) NTTB200_CHECK(cudaStreamSynchronize(c->streams[i]));
    return 0;
}
int nttb200_forward_ntt_batch_host_packed(nttb200_ctx *c, const nttb200_u64 *in, nttb200_u64 *out, unsigned num, unsigned division)
{
    return host_pipeline_packed(c, false, in, out, num, division);
}
int nttb200_inverse_ntt_batch_host_packed(nttb200_ctx *c, const nttb200_u64 *in, nttb200_u64 *out, unsigned num, unsigned division)
{
    return host_pipeline_packed(c, true, in, out, num, division);
}

int nttb200_forward_ntt_batch_host(nttb200_ctx *c, const nttb200_u64 *in, nttb200_u64 *out, unsigned num, unsigned division)
{
    return host_pipeline(c, false, in, out, num, division);
}
int nttb200_inverse_ntt_batch_host(nttb200_ctx *c, const nttb200_u64 *in, nttb200_u64 *out, unsigned num, unsigned division)
{
    return host_pipeline(c, true, in, out, num, division);
}

}  // extern "C"
