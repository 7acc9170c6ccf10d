// bfv_host.cu -- BFV encryption / decryption through HOST buffers (the end-to-end path a caller outside CUDA sees; demo.cu:275-299
// times keygen / encryption / decryption and copies the plaintext back) and the host side of the compact wire format.
//
// Three streams, double-buffered device staging: H2D of chunk i+1, the kernels of chunk i and D2H of chunk i-1 overlap.  The BFV
// context's sampling scratch serves one stream at a time, so every kernel runs on ONE compute stream; only the copies use their own.
// Ciphertexts cross PCIe in the compact wire format when `packed` is set (n * qbit_l bits per limb, padding limb dropped:
// 6.45 MB instead of 8 MiB per ciphertext at (32768, 16 limbs)), else in the reference layout c[2][r][n].
#include "bfv_internal.h"

#include <algorithm>

using namespace nttb200;

struct nttb200_host_state {
    cudaStream_t h2d = nullptr, comp = nullptr, d2h = nullptr;
    cudaEvent_t in_ready[2], in_free[2], out_ready[2], out_free[2];
    u64 *m[2] = {nullptr, nullptr}, *c[2] = {nullptr, nullptr}, *p[2] = {nullptr, nullptr};
    size_t chunk = 0;     // items per chunk the buffers are sized for
    bool ok = false;
};

void nttb200_host_state_destroy(nttb200_host_state *s)
{
    if (!s) return;
    if (s->ok) {
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(s->in_ready[i]); cudaEventDestroy(s->in_free[i]); cudaEventDestroy(s->out_ready[i]); cudaEventDestroy(s->out_free[i]);
        }
        cudaStreamDestroy(s->h2d); cudaStreamDestroy(s->comp); cudaStreamDestroy(s->d2h);
    }
    for (int i = 0; i < 2; i++) { cudaFree(s->m[i]); cudaFree(s->c[i]); cudaFree(s->p[i]); }
    delete s;
}

static int host_state(nttb200_bfv *b, nttb200_host_state **out, size_t chunk)
{
    if (!b->host) b->host = new nttb200_host_state();
    nttb200_host_state *s = b->host;
    if (!s->ok) {
        NTTB200_CHECK(cudaStreamCreateWithFlags(&s->h2d, cudaStreamNonBlocking));
        NTTB200_CHECK(cudaStreamCreateWithFlags(&s->comp, cudaStreamNonBlocking));
        NTTB200_CHECK(cudaStreamCreateWithFlags(&s->d2h, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            NTTB200_CHECK(cudaEventCreateWithFlags(&s->in_ready[i], cudaEventDisableTiming));
            NTTB200_CHECK(cudaEventCreateWithFlags(&s->in_free[i], cudaEventDisableTiming));
            NTTB200_CHECK(cudaEventCreateWithFlags(&s->out_ready[i], cudaEventDisableTiming));
            NTTB200_CHECK(cudaEventCreateWithFlags(&s->out_free[i], cudaEventDisableTiming));
        }
        s->ok = true;
    }
    if (s->chunk < chunk) {
        const size_t n = b->n, rn = (size_t)b->r * n, pw = 2 * (size_t)b->half_words;
        for (int i = 0; i < 2; i++) {
            cudaFree(s->m[i]); cudaFree(s->c[i]); cudaFree(s->p[i]);
            s->m[i] = s->c[i] = s->p[i] = nullptr;
            NTTB200_CHECK(cudaMalloc(&s->m[i], chunk * n * 8));
            NTTB200_CHECK(cudaMalloc(&s->c[i], chunk * 2 * rn * 8));
            NTTB200_CHECK(cudaMalloc(&s->p[i], chunk * pw * 8));
        }
        s->chunk = chunk;
    }
    *out = s;
    return 0;
}

static size_t pick_chunk(const nttb200_bfv *b, unsigned batch)
{
    // about 128 MiB of ciphertext per chunk, at least 1 item, at most the batch
    const size_t per = (size_t)2 * b->r * b->n * 8;
    size_t k = ((size_t)128 << 20) / per;
    if (k < 1) k = 1;
    if (k > batch) k = batch;
    return k;
}

extern "C" {

// m_host[batch][n] -> ciphertexts in c_host: packed != 0 -> batch * nttb200_bfv_packed_words() words, else c[batch][2][r][n].
// Uses the loaded public key.  Synchronous.  Pinned host buffers make the copies asynchronous (pageable ones still work).
int nttb200_bfv_encrypt_host(nttb200_bfv *b, nttb200_u64 *c_host, int packed, const nttb200_u64 *m_host, unsigned batch, nttb200_u64 nonce0)
{
    if (!b || !c_host || !m_host || !batch || !b->pk_l) return NTTB200_EINVAL;
    const size_t n = b->n, rn = (size_t)b->r * n, pw = 2 * (size_t)b->half_words, K = pick_chunk(b, batch);
    nttb200_host_state *s;
    NTTB200_TRY(host_state(b, &s, K));
    size_t done = 0;
    for (int i = 0; done < batch; i++, done += K) {
        const int k = i & 1;
        const unsigned cnt = (unsigned)std::min(K, (size_t)batch - done);
        if (i >= 2) NTTB200_CHECK(cudaStreamWaitEvent(s->h2d, s->in_free[k], 0));
        NTTB200_CHECK(cudaMemcpyAsync(s->m[k], m_host + done * n, cnt * n * 8, cudaMemcpyHostToDevice, s->h2d));
        NTTB200_CHECK(cudaEventRecord(s->in_ready[k], s->h2d));
        NTTB200_CHECK(cudaStreamWaitEvent(s->comp, s->in_ready[k], 0));
        if (i >= 2) NTTB200_CHECK(cudaStreamWaitEvent(s->comp, s->out_free[k], 0));
        NTTB200_TRY(nttb200_bfv_encrypt(b, s->c[k], nullptr, 0, s->m[k], cnt, nonce0 + done, s->comp));
        if (packed) NTTB200_TRY(nttb200_bfv_pack(b, s->p[k], s->c[k], cnt, s->comp));
        NTTB200_CHECK(cudaEventRecord(s->in_free[k], s->comp));
        NTTB200_CHECK(cudaEventRecord(s->out_ready[k], s->comp));
        NTTB200_CHECK(cudaStreamWaitEvent(s->d2h, s->out_ready[k], 0));
        if (packed) NTTB200_CHECK(cudaMemcpyAsync(c_host + done * pw, s->p[k], cnt * pw * 8, cudaMemcpyDeviceToHost, s->d2h));
        else NTTB200_CHECK(cudaMemcpyAsync(c_host + done * 2 * rn, s->c[k], cnt * 2 * rn * 8, cudaMemcpyDeviceToHost, s->d2h));
        NTTB200_CHECK(cudaEventRecord(s->out_free[k], s->d2h));
    }
    NTTB200_CHECK(cudaStreamSynchronize(s->d2h));
    NTTB200_CHECK(cudaStreamSynchronize(s->comp));
    return 0;
}

// ciphertexts in c_host (format as above) -> m_host[batch][n].  Uses the loaded secret key.  Synchronous.
int nttb200_bfv_decrypt_host(nttb200_bfv *b, nttb200_u64 *m_host, const nttb200_u64 *c_host, int packed, unsigned batch)
{
    if (!b || !c_host || !m_host || !batch || !b->sk_l) return NTTB200_EINVAL;
    const size_t n = b->n, rn = (size_t)b->r * n, pw = 2 * (size_t)b->half_words, K = pick_chunk(b, batch);
    nttb200_host_state *s;
    NTTB200_TRY(host_state(b, &s, K));
    size_t done = 0;
    for (int i = 0; done < batch; i++, done += K) {
        const int k = i & 1;
        const unsigned cnt = (unsigned)std::min(K, (size_t)batch - done);
        if (i >= 2) NTTB200_CHECK(cudaStreamWaitEvent(s->h2d, s->in_free[k], 0));
        if (packed) NTTB200_CHECK(cudaMemcpyAsync(s->p[k], c_host + done * pw, cnt * pw * 8, cudaMemcpyHostToDevice, s->h2d));
        else NTTB200_CHECK(cudaMemcpyAsync(s->c[k], c_host + done * 2 * rn, cnt * 2 * rn * 8, cudaMemcpyHostToDevice, s->h2d));
        NTTB200_CHECK(cudaEventRecord(s->in_ready[k], s->h2d));
        NTTB200_CHECK(cudaStreamWaitEvent(s->comp, s->in_ready[k], 0));
        if (i >= 2) NTTB200_CHECK(cudaStreamWaitEvent(s->comp, s->out_free[k], 0));
        if (packed) NTTB200_TRY(nttb200_bfv_unpack(b, s->c[k], s->p[k], cnt, s->comp));
        NTTB200_TRY(nttb200_bfv_decrypt(b, s->m[k], s->c[k], nullptr, 0, cnt, s->comp));
        NTTB200_CHECK(cudaEventRecord(s->in_free[k], s->comp));
        NTTB200_CHECK(cudaEventRecord(s->out_ready[k], s->comp));
        NTTB200_CHECK(cudaStreamWaitEvent(s->d2h, s->out_ready[k], 0));
        NTTB200_CHECK(cudaMemcpyAsync(m_host + done * n, s->m[k], cnt * n * 8, cudaMemcpyDeviceToHost, s->d2h));
        NTTB200_CHECK(cudaEventRecord(s->out_free[k], s->d2h));
    }
    NTTB200_CHECK(cudaStreamSynchronize(s->d2h));
    NTTB200_CHECK(cudaStreamSynchronize(s->comp));
    return 0;
}

// ---- key wire format: a secret key sk[r][n] or a public key pk[2][r][n] (NTT-domain residues below q_l) as n * qbit_l
// bits per limb, ALL r limbs, little-endian bits in little-endian words -- the ciphertext format of nttb200_bfv_pack without the dropped
// limb.  polys = r (secret key) or 2r (public key).  Device buffers; the _host variants stage through the device.
size_t nttb200_bfv_key_packed_words(const nttb200_bfv *b, unsigned polys)
{
    if (!b || (polys != b->r && polys != 2 * b->r)) return 0;
    return (size_t)(polys / b->r) * b->key_half_words;
}

}  // extern "C"
