// sharded.cu -- limb-sharded BFV encryption / decryption across the GPUs of one NVSwitch box (BASELINE configs 4-5).
//
// The reference is single-GPU.  Units (ciphertext, limb) are independent everywhere except
//   * encryption's modulus switch: every limb needs the DROPPED limb's rounded value (bfv_encryption.cuh:146-149), and
//   * decryption's base conversion: one sum over the limbs (fast_convert_array_kernel_t / _gamma, poly_arithmetic.cuh:217-251).
// Partition (nttb200_shard_plan): the batch is cut into `world` item blocks; unit (limb l, block j) has flat index l * world + j and
// rank g owns flat indices [g * rp, (g + 1) * rp) -- exactly rp (limb, block) tiles per rank whatever rp and world are (15 limbs on 8
// GPUs: no idle rank), and for a fixed block the limbs a rank owns are a contiguous range.  The dropped limb of block j is computed by
// rank j, once, and all-gathered (it is 1/r of the work: recomputing it on every rank, as round 1 did, costs 8/15 extra at 8 GPUs).
// Decryption: per block, partial base-conversion sums packed to 10 bytes per coefficient -> ncclReduce to the block's owner (or a
// chunked ncclReduceScatter) on a second stream while the next block's transforms run -> rounding on the owner -> ONE all-gather of
// 16-bit plaintext words.  Bit-identical to the single-GPU calls (tests/test_gpu_sharded.py, scripts/multigpu_check.py).
//
// NCCL is bound at run time (dlopen of the libnccl already in the process, e.g. PyTorch's, else the system one): libnttb200.so
// itself has no NCCL link dependency and single-GPU users never load it.
#include "bfv_internal.h"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace nttb200;

// ---- NCCL, bound at run time ---------------------------------------------------------------------------------------------------
namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommCount)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*CommUserRank)(const ncclComm_t, int *) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi &nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    // the library that created the caller's communicator must be the one we call: prefer what the process already holds
    const char *names[] = {getenv("NTTB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (int pass = 0; pass < 2 && !api.handle; pass++)
        for (const char *nm : names) {
            if (!nm) continue;
            api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL | (pass == 0 ? RTLD_NOLOAD : 0));
            if (api.handle) break;
        }
    if (!api.handle) return api;
#define NTTB200_SYM(f) *(void **)(&api.f) = dlsym(api.handle, "nccl" #f)
    NTTB200_SYM(GetUniqueId); NTTB200_SYM(CommInitRank); NTTB200_SYM(CommDestroy); NTTB200_SYM(CommCount); NTTB200_SYM(CommUserRank);
    NTTB200_SYM(AllGather); NTTB200_SYM(ReduceScatter); NTTB200_SYM(Reduce); NTTB200_SYM(Broadcast); NTTB200_SYM(AllReduce); NTTB200_SYM(GroupStart); NTTB200_SYM(GroupEnd);
    NTTB200_SYM(GetErrorString);
#undef NTTB200_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.ReduceScatter && api.Reduce && api.Broadcast && api.AllReduce && api.GroupStart &&
             api.GroupEnd;
    return api;
}
int nccl_fail(ncclResult_t r, const char *file, int line)
{
    const char *e = getenv("NTTB200_DEBUG");
    if (e && e[0] == '1') fprintf(stderr, "nttb200: NCCL error %d (%s) at %s:%d\n", (int)r, nccl().GetErrorString ? nccl().GetErrorString(r) : "?", file, line);
    return NTTB200_ENCCL;
}
}  // namespace
#define NCCLCHECK(x) do { ncclResult_t r__ = (x); if (r__ != ncclSuccess) return nccl_fail(r__, __FILE__, __LINE__); } while (0)
// collectives of the sharded calls: skipped for a fake communicator (compute-only profiling of one rank's share)
#define COLL(x) do { if (!comm->fake) NCCLCHECK(x); } while (0)
#define TRY(x) do { int r__ = (x); if (r__) return nttb200_trace_error(r__, __FILE__, __LINE__); } while (0)

struct nttb200_comm {
    ncclComm_t comm = nullptr;
    int world = 1, rank = 0;
    bool owned = false;
    bool fake = false;        // profiling only (nttb200_comm_fake): one rank's share of the work with every collective skipped
};

// scratch + second stream of the sharded entry points, owned by the BFV context (grow-only: warm up before graph capture)
struct nttb200_shard_state {
    cudaStream_t cs = nullptr;                    // collectives (and the owner's rounding) run here, next to the transforms
    std::vector<cudaEvent_t> ev;
    unsigned char *buf[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    size_t cap[6] = {0, 0, 0, 0, 0, 0};
    int mode = 4;                                 // decryption's exchange.  4 (default): partial sums into local scratch, pushed into the owner's
                                                  //    slot by the copy engines; 2 / 3: stored there directly by the kernel (3: one round);
                                                  //    0: per-piece ncclReduce to the owner; 1: chunked ncclReduceScatter.  2-4 need CUDA IPC and
                                                  //    fall back to 0 on all ranks without it.
    unsigned chunks = 0;                          // rounds / pieces per block; 0 = the measured best of the mode (4: 2, others: 4)
    // symmetric buffers: one cudaMalloc per rank, mapped into every other rank through CUDA IPC (peer stores / copy-engine pushes)
    struct Sym {
        unsigned char *local = nullptr;
        size_t bytes = 0;
        std::vector<unsigned char *> peer;        // peer[g] = rank g's buffer as mapped here (peer[rank] = local)
    } sym[4];
    int p2p_failed = 0;
    int *flag = nullptr;                          // device ints of the barrier / agreement all-reduces
    cudaStream_t st2 = nullptr;                   // second compute stream: independent tiles overlap their launch tails
};
enum { kSymSlots = 0, kSymCl, kSymEs, kSymUb };
enum { kBufUb = 0, kBufEs, kBufCl, kBufPartial, kBufRecv, kBufPlain };

static void sym_release(nttb200_shard_state::Sym &y)
{
    for (size_t g = 0; g < y.peer.size(); g++)
        if (y.peer[g] && y.peer[g] != y.local) cudaIpcCloseMemHandle(y.peer[g]);
    y.peer.clear();
    if (y.local) cudaFree(y.local);
    y.local = nullptr; y.bytes = 0;
}
void nttb200_shard_state_destroy(nttb200_shard_state *s)
{
    if (!s) return;
    for (auto &y : s->sym) sym_release(y);
    if (s->flag) cudaFree(s->flag);
    if (s->st2) cudaStreamDestroy(s->st2);
    for (auto e : s->ev) cudaEventDestroy(e);
    for (auto p : s->buf) if (p) cudaFree(p);
    if (s->cs) cudaStreamDestroy(s->cs);
    delete s;
}
static int shard_state(nttb200_bfv *b, nttb200_shard_state **out, size_t events)
{
    if (!b->shard) {
        b->shard = new nttb200_shard_state();
        if (const char *e = getenv("NTTB200_SHARD_MODE")) b->shard->mode = atoi(e);
        if (const char *e = getenv("NTTB200_SHARD_CHUNKS")) b->shard->chunks = (unsigned)atoi(e);
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);          // collectives / rounding of finished pieces go ahead of queued transforms
        NTTB200_CHECK(cudaStreamCreateWithPriority(&b->shard->cs, cudaStreamNonBlocking, hi));
        NTTB200_CHECK(cudaStreamCreateWithFlags(&b->shard->st2, cudaStreamNonBlocking));
    }
    nttb200_shard_state *s = b->shard;
    while (s->ev.size() < events) {
        cudaEvent_t e;
        NTTB200_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        s->ev.push_back(e);
    }
    *out = s;
    return 0;
}
static int shard_buf(nttb200_shard_state *s, int which, size_t bytes, void **p)
{
    if (s->cap[which] < bytes) {
        if (s->buf[which]) NTTB200_CHECK(cudaFree(s->buf[which]));
        s->buf[which] = nullptr; s->cap[which] = 0;
        NTTB200_CHECK(cudaMalloc(&s->buf[which], bytes));
        s->cap[which] = bytes;
    }
    *p = s->buf[which];
    return 0;
}

static void plan_blocks(unsigned rp, unsigned n, unsigned batch, unsigned world, unsigned rank, nttb200_shard_block *blk, size_t *words)
{
    const unsigned per = batch / world;
    size_t off = 0;
    for (unsigned j = 0; j < world; j++) {
        // limbs l with rank * rp <= l * world + j < (rank + 1) * rp
        const long lo_num = (long)rank * rp - (long)j, hi_num = (long)(rank + 1) * rp - 1 - (long)j;
        long lmin = lo_num <= 0 ? 0 : (lo_num + world - 1) / world;
        long lmax = hi_num < 0 ? -1 : hi_num / (long)world;
        if (lmax > (long)rp - 1) lmax = (long)rp - 1;
        const unsigned cnt = lmax >= lmin ? (unsigned)(lmax - lmin + 1) : 0u;
        blk[j].first_item = j * per; blk[j].items = per;
        blk[j].first_limb = cnt ? (unsigned)lmin : 0u; blk[j].limb_count = cnt;
        blk[j].offset = off;
        off += (size_t)per * 2 * cnt * n;
    }
    if (words) *words = off;
}

// Consecutive item blocks in which the rank holds the SAME limb window are contiguous in its shard buffer (and in every per-item
// array), so their transforms run as one launch set: at (16384, 9 limbs) on 8 GPUs a rank owns one limb of every block = ONE run of
// 4096 items instead of eight launch sets of 512; at (32768, 16 limbs) three runs.  Large runs are cut in two so that both compute
// streams have work.
struct ShardRun { unsigned j0, j1, first_limb, cnt, first_item, items; size_t offset; int stream; };
static std::vector<ShardRun> shard_runs(const std::vector<nttb200_shard_block> &blk, unsigned per, unsigned n)
{
    std::vector<ShardRun> runs;
    const unsigned G = (unsigned)blk.size();
    for (unsigned j = 0; j < G;) {
        if (!blk[j].limb_count) { j++; continue; }
        unsigned k = j + 1;
        while (k < G && blk[k].limb_count == blk[j].limb_count && blk[k].first_limb == blk[j].first_limb) k++;
        runs.push_back(ShardRun{j, k, blk[j].first_limb, blk[j].limb_count, j * per, (k - j) * per, blk[j].offset, 0});
        j = k;
    }
    // at least two parts of comparable size: cut the largest run until there are >= 2 (whole blocks when possible)
    while (runs.size() == 1 && runs[0].items >= 2) {
        ShardRun a = runs[0], c = runs[0];
        const unsigned blocks = a.j1 - a.j0;
        const unsigned cut_items = blocks >= 2 ? (blocks / 2) * per : a.items / 2;
        a.items = cut_items; a.j1 = a.j0 + (blocks >= 2 ? blocks / 2 : blocks);
        c.first_item = a.first_item + cut_items; c.items = runs[0].items - cut_items; c.j0 = blocks >= 2 ? a.j1 : a.j0;
        c.offset = a.offset + (size_t)cut_items * 2 * a.cnt * n;
        runs.clear(); runs.push_back(a); runs.push_back(c);
    }
    // alternate streams, largest first on the caller's stream
    for (size_t i = 0; i < runs.size(); i++) runs[i].stream = (int)(i & 1);
    return runs;
}

// Symmetric-buffer set-up (collective: every rank calls it with the same index and size): allocate this rank's buffer, exchange
// CUDA IPC handles through an NCCL all-gather, map every peer's buffer.  Any failure (IPC unsupported in this environment) disables
// the peer-to-peer paths on ALL ranks -- the outcome is agreed through an all-reduce so that no rank is left in a different protocol.
static int sym_setup(nttb200_shard_state *s, nttb200_comm *comm, int which, size_t bytes)
{
    const unsigned G = (unsigned)comm->world, g = (unsigned)comm->rank;
    nttb200_shard_state::Sym &y = s->sym[which];
    if (!s->flag) { NTTB200_CHECK(cudaMalloc(&s->flag, 4 * sizeof(int))); NTTB200_CHECK(cudaMemset(s->flag, 0, 4 * sizeof(int))); }
    if (s->p2p_failed) return 0;
    if (y.bytes >= bytes && y.peer.size() == G) return 0;
    NTTB200_CHECK(cudaStreamSynchronize(s->cs));
    sym_release(y);
    int bad = 0;
    if (cudaMalloc(&y.local, bytes) != cudaSuccess) { bad = 1; y.local = nullptr; cudaGetLastError(); }
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (!bad && cudaIpcGetMemHandle(&mine, y.local) != cudaSuccess) { bad = 1; cudaGetLastError(); }
    unsigned char *dev = nullptr;
    NTTB200_CHECK(cudaMalloc(&dev, (size_t)G * sizeof mine));
    NTTB200_CHECK(cudaMemcpy(dev + (size_t)g * sizeof mine, &mine, sizeof mine, cudaMemcpyHostToDevice));
    NCCLCHECK(nccl().AllGather(dev + (size_t)g * sizeof mine, dev, sizeof mine, ncclInt8, comm->comm, s->cs));
    NTTB200_CHECK(cudaStreamSynchronize(s->cs));
    std::vector<cudaIpcMemHandle_t> all(G);
    NTTB200_CHECK(cudaMemcpy(all.data(), dev, (size_t)G * sizeof mine, cudaMemcpyDeviceToHost));
    cudaFree(dev);
    y.peer.assign(G, nullptr);
    for (unsigned k = 0; k < G && !bad; k++) {
        if (k == g) { y.peer[k] = y.local; continue; }
        void *p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { bad = 1; cudaGetLastError(); break; }
        y.peer[k] = (unsigned char *)p;
    }
    int h = bad;                                      // agree on the outcome
    NTTB200_CHECK(cudaMemcpy(s->flag + 1, &h, sizeof h, cudaMemcpyHostToDevice));
    NCCLCHECK(nccl().AllReduce(s->flag + 1, s->flag + 1, 1, ncclInt32, ncclSum, comm->comm, s->cs));
    NTTB200_CHECK(cudaStreamSynchronize(s->cs));
    NTTB200_CHECK(cudaMemcpy(&h, s->flag + 1, sizeof h, cudaMemcpyDeviceToHost));
    if (h) {
        for (auto &z : s->sym) sym_release(z);
        s->p2p_failed = 1;
        if (const char *e = getenv("NTTB200_DEBUG")) if (e[0] == '1') fprintf(stderr, "nttb200: CUDA IPC unavailable on %d rank(s): the sharded calls fall back to NCCL collectives\n", h);
        return 0;
    }
    y.bytes = bytes;
    return 0;
}
// "every rank has reached this point of its stream": a 4-byte all-reduce
#define BARRIER() COLL(nccl().AllReduce(s->flag, s->flag, 1, ncclInt32, ncclSum, comm->comm, s->cs))

extern "C" {

int nttb200_shard_plan(unsigned rp, unsigned n, unsigned batch, unsigned world, unsigned rank, nttb200_shard_block *blocks, size_t *shard_words)
{
    if (!blocks || !rp || !world || rank >= world || !batch || batch % world) return NTTB200_EINVAL;
    plan_blocks(rp, n, batch, world, rank, blocks, shard_words);
    return 0;
}

int nttb200_comm_unique_id(unsigned char id[128])
{
    if (!id || !nccl().ok) return NTTB200_ENCCL;
    ncclUniqueId u;
    NCCLCHECK(nccl().GetUniqueId(&u));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return 0;
}
int nttb200_comm_create(nttb200_comm **out, const unsigned char id[128], int world, int rank)
{
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return NTTB200_EINVAL;
    nttb200_comm *c = new nttb200_comm();
    c->world = world; c->rank = rank;
    if (world > 1) {
        if (!nccl().ok) { delete c; return NTTB200_ENCCL; }
        ncclUniqueId u;
        memcpy(&u, id, 128);
        ncclResult_t r = nccl().CommInitRank(&c->comm, world, u, rank);
        if (r != ncclSuccess) { delete c; return nccl_fail(r, __FILE__, __LINE__); }
        c->owned = true;
    }
    *out = c;
    return 0;
}
int nttb200_comm_adopt(nttb200_comm **out, void *nccl_comm, int world, int rank)
{
    if (!out || world < 1 || rank < 0 || rank >= world || (world > 1 && !nccl_comm)) return NTTB200_EINVAL;
    if (world > 1 && !nccl().ok) return NTTB200_ENCCL;
    nttb200_comm *c = new nttb200_comm();
    c->comm = (ncclComm_t)nccl_comm; c->world = world; c->rank = rank; c->owned = false;
    if (world > 1 && nccl().CommCount && nccl().CommUserRank) {
        int w = 0, r = 0;
        if (nccl().CommCount(c->comm, &w) != ncclSuccess || nccl().CommUserRank(c->comm, &r) != ncclSuccess || w != world || r != rank) {
            delete c;
            return NTTB200_EINVAL;
        }
    }
    *out = c;
    return 0;
}
// profiling only: behaves as rank `rank` of `world` but skips every collective (results are meaningless; timings are one rank's compute)
int nttb200_comm_fake(nttb200_comm **out, int world, int rank)
{
    if (!out || world < 1 || rank < 0 || rank >= world) return NTTB200_EINVAL;
    nttb200_comm *c = new nttb200_comm();
    c->world = world; c->rank = rank; c->fake = true;
    *out = c;
    return 0;
}
void nttb200_comm_destroy(nttb200_comm *c)
{
    if (!c) return;
    if (c->owned && c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}
int nttb200_comm_world(const nttb200_comm *c) { return c ? c->world : 0; }
int nttb200_comm_rank(const nttb200_comm *c) { return c ? c->rank : -1; }

int nttb200_bfv_shard_config(nttb200_bfv *b, int mode, unsigned chunks)
{
    if (!b || mode < 0 || mode > 4) return NTTB200_EINVAL;
    nttb200_shard_state *s;
    TRY(shard_state(b, &s, 0));
    s->mode = mode;
    s->chunks = chunks;
    return 0;
}

// reference layout c[batch][2][r][n] (this device) <-> the calling rank's shard (nttb200_shard_plan), device-local copies
static int shard_copy(nttb200_bfv *b, unsigned world, unsigned rank, nttb200_u64 *c_shard, nttb200_u64 *c_full, unsigned batch, bool to_shard, cudaStream_t st)
{
    if (!b || !c_shard || !c_full || !world || rank >= world || !batch || batch % world) return NTTB200_EINVAL;
    const unsigned n = b->n, r = b->r;
    std::vector<nttb200_shard_block> blk(world);
    plan_blocks(r - 1, n, batch, world, rank, blk.data(), nullptr);
    for (unsigned j = 0; j < world; j++) {
        const unsigned cnt = blk[j].limb_count;
        if (!cnt) continue;
        for (unsigned h = 0; h < 2; h++) {
            u64 *full = c_full + ((size_t)blk[j].first_item * 2 + h) * r * n + (size_t)blk[j].first_limb * n;
            u64 *sh = c_shard + blk[j].offset + (size_t)h * cnt * n;
            const size_t fp = (size_t)2 * r * n * 8, sp = (size_t)2 * cnt * n * 8, w = (size_t)cnt * n * 8;
            if (to_shard) NTTB200_CHECK(cudaMemcpy2DAsync(sh, sp, full, fp, w, blk[j].items, cudaMemcpyDeviceToDevice, st));
            else NTTB200_CHECK(cudaMemcpy2DAsync(full, fp, sh, sp, w, blk[j].items, cudaMemcpyDeviceToDevice, st));
        }
    }
    return 0;
}
int nttb200_bfv_shard_from_full(nttb200_bfv *b, unsigned world, unsigned rank, nttb200_u64 *c_shard, const nttb200_u64 *c_full, unsigned batch, void *stream)
{
    return shard_copy(b, world, rank, c_shard, const_cast<nttb200_u64 *>(c_full), batch, true, (cudaStream_t)stream);
}
int nttb200_bfv_shard_to_full(nttb200_bfv *b, unsigned world, unsigned rank, nttb200_u64 *c_full, const nttb200_u64 *c_shard, unsigned batch, void *stream)
{
    return shard_copy(b, world, rank, const_cast<nttb200_u64 *>(c_shard), c_full, batch, false, (cudaStream_t)stream);
}

// ---- building blocks for callers that run their own collectives (and for single-GPU tests with virtual ranks) ------------------------
// limbs [first_limb, first_limb + limb_count) of `batch` ciphertexts, c_tile[batch][2][limb_count][n] (one tile of nttb200_shard_plan),
// loaded secret key, fused NTT (.) sk -> INTT kernel.  packed = 1: partial[batch][n + n/4] (gamma sums, then 16-bit t sums; needs
// (r-1) * (t-1) < 2^16), packed = 0: partial[batch][2][n] as nttb200_bfv_decrypt_partial writes it.  SUM-reduce over the limb windows.
int nttb200_bfv_decrypt_partial_tile(nttb200_bfv *b, nttb200_u64 *partial, int packed, nttb200_u64 *c_tile, unsigned first_limb, unsigned limb_count,
                                     unsigned batch, void *stream)
{
    if (!b || !partial || !c_tile || !batch || batch > 65535 || !limb_count || first_limb + limb_count > b->r - 1 || !b->sk_l) return NTTB200_EINVAL;
    if (packed && (u64)(b->r - 1) * (b->t - 1) >= 65536) return NTTB200_EINVAL;
    Pipe P = pipe_from_bfv(b, (cudaStream_t)stream);
    return dec_partial(b, P, partial, packed, c_tile, limb_count, first_limb, limb_count, batch);
}
// out16 = 1: m_out is unsigned short[batch][n] (t <= 2^16), else nttb200_u64[batch][n]
int nttb200_bfv_decrypt_finish_tile(nttb200_bfv *b, void *m_out, int out16, const nttb200_u64 *partial_sum, int packed, unsigned batch, void *stream)
{
    if (!b || !m_out || !partial_sum || !batch || batch > 65535 || (out16 && b->t > 65536)) return NTTB200_EINVAL;
    return dec_finish(b, m_out, out16, partial_sum, packed, batch, (cudaStream_t)stream);
}

// ---- limb-sharded encryption --------------------------------------------------------------------------------------------------------
// Every rank passes the same m[batch][n] and nonce0; c_shard receives the rank's (limb, block) tiles.  The loaded public key is used.
// Independent tiles alternate between the caller's stream and a second compute stream so that the launch tails of one tile's kernels
// overlap the next tile's (a tile is 512 items x 1-3 limbs at 8 GPUs: 9-28 waves per kernel).
int nttb200_bfv_encrypt_sharded(nttb200_bfv *b, nttb200_comm *comm, nttb200_u64 *c_shard, const nttb200_u64 *m, unsigned batch, nttb200_u64 nonce0,
                                void *stream)
{
    if (!b || !comm || !c_shard || !m || !batch || batch % (unsigned)comm->world || !b->pk_l) return NTTB200_EINVAL;
    if (!b->ctx->lazy_ok || !b->epi_ok) return NTTB200_EINVAL;            // every q_i < 2^57 with an exact reference Barrett (all reference sets but 4k_3q / 8k_3q)
    const unsigned G = (unsigned)comm->world, g = (unsigned)comm->rank, n = b->n, r = b->r, per = batch / G;
    if (per > 65535) return NTTB200_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    nttb200_shard_state *s;
    TRY(shard_state(b, &s, 10));
    std::vector<nttb200_shard_block> blk(G);
    plan_blocks(r - 1, n, batch, G, g, blk.data(), nullptr);
    unsigned char *ub; signed char *es; u64 *cl;
    TRY(shard_buf(s, kBufUb, (size_t)batch * n, (void **)&ub));
    // finished dropped limbs cl[batch][2][n] and signed-byte draws es[batch][2][n]: symmetric buffers when CUDA IPC is available (every
    // rank PUSHES its block into the peers' buffers with copy-engine copies: no collective kernel takes SMs from the transforms),
    // plain scratch + ncclAllGather otherwise
    const bool want_p2p = G > 1 && !comm->fake && s->mode >= 2;
    if (want_p2p) {
        TRY(sym_setup(s, comm, kSymCl, (size_t)batch * 2 * n * 8)); TRY(sym_setup(s, comm, kSymEs, (size_t)batch * 2 * n));
        TRY(sym_setup(s, comm, kSymUb, (size_t)batch * n));
    }
    const bool p2p = want_p2p && !s->p2p_failed;
    if (p2p) { cl = (u64 *)s->sym[kSymCl].local; es = (signed char *)s->sym[kSymEs].local; ub = s->sym[kSymUb].local; }
    else { TRY(shard_buf(s, kBufEs, (size_t)batch * 2 * n, (void **)&es)); TRY(shard_buf(s, kBufCl, (size_t)batch * 2 * n * 8, (void **)&cl)); }
    Pipe P = pipe_from_bfv(b, st), P2 = pipe_from_bfv(b, s->st2);
    const size_t own = (size_t)g * per;
    // the previous call's exchange (comm stream) must be done before this call overwrites its buffers; with one-sided pushes every
    // rank must also have finished READING the previous call's cl / es before anybody overwrites them: a barrier at entry
    NTTB200_CHECK(cudaEventRecord(s->ev[0], st));
    NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[0], 0));
    if (p2p) BARRIER();
    NTTB200_CHECK(cudaEventRecord(s->ev[1], s->cs));
    NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[1], 0));
    // 1. randomness.  Every rank transforms u on its limbs of EVERY item, but drawing all of it on every rank costs world x the
    //    sampling (0.25 ms per rank at 4096 ciphertexts): with peer-to-peer buffers each rank draws u and e of its own block only and
    //    pushes the u bytes (n per item) to the peers; otherwise it draws every item's u itself.
    if (p2p) {
        TRY(enc_sample(b, ub + own * n, es + own * 2 * n, per, nonce0 + own, 1, 1, st));
        NTTB200_CHECK(cudaEventRecord(s->ev[5], st));
        NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[5], 0));
        for (unsigned tau = 1; tau < G; tau++) {
            const unsigned k = (g + tau) % G;
            NTTB200_CHECK(cudaMemcpyAsync(s->sym[kSymUb].peer[k] + own * n, ub + own * n, (size_t)per * n, cudaMemcpyDeviceToDevice, s->cs));
        }
        BARRIER();
        NTTB200_CHECK(cudaEventRecord(s->ev[6], s->cs));           // every block's u bytes have arrived
    } else {
        TRY(enc_sample(b, ub, nullptr, batch, nonce0, 1, 0, st));
        TRY(enc_sample(b, nullptr, es + own * 2 * n, per, nonce0 + own, 0, 1, st));
    }
    // 2. dropped limb of the own block, finished (+ e, rounding offset)
    u64 *cl_own = cl + own * 2 * n;
    TRY(enc_front(b, P, cl_own, 1, r - 1, 1, per, ub + own * n));
    TRY(enc_finish_last(b, P, cl_own, (size_t)2 * n, (size_t)n, es + own * 2 * n, per));
    NTTB200_CHECK(cudaEventRecord(s->ev[2], st));
    // 3. exchange of the finished dropped limbs and the draws, on the comm stream, under the transforms of step 4
    if (G > 1) {
        NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[2], 0));
        if (p2p) {
            for (unsigned tau = 1; tau < G; tau++) {          // rotated order: one incoming copy per rank at a time
                const unsigned k = (g + tau) % G;
                NTTB200_CHECK(cudaMemcpyAsync((u64 *)s->sym[kSymCl].peer[k] + own * 2 * n, cl_own, (size_t)per * 2 * n * 8, cudaMemcpyDeviceToDevice, s->cs));
                NTTB200_CHECK(cudaMemcpyAsync(s->sym[kSymEs].peer[k] + own * 2 * n, es + own * 2 * n, (size_t)per * 2 * n, cudaMemcpyDeviceToDevice, s->cs));
            }
            BARRIER();                                         // every push has landed everywhere
        } else {
            COLL(nccl().GroupStart());
            COLL(nccl().AllGather(cl_own, cl, (size_t)per * 2 * n, ncclUint64, comm->comm, s->cs));
            COLL(nccl().AllGather(es + own * 2 * n, es, (size_t)per * 2 * n, ncclInt8, comm->comm, s->cs));
            COLL(nccl().GroupEnd());
        }
        NTTB200_CHECK(cudaEventRecord(s->ev[3], s->cs));
    }
    // 4. forward transform of u, (.) pk, contiguous inverse pass on the owned (limb, block) tiles; tiles alternate between two streams
    NTTB200_CHECK(cudaStreamWaitEvent(s->st2, s->ev[2], 0));   // the second stream joins here
    if (p2p) { NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[6], 0)); NTTB200_CHECK(cudaStreamWaitEvent(s->st2, s->ev[6], 0)); }   // the peers' u bytes
    const std::vector<ShardRun> runs = shard_runs(blk, per, n);
    for (const ShardRun &R : runs)
        TRY(enc_front(b, R.stream ? P2 : P, c_shard + R.offset, R.cnt, R.first_limb, R.cnt, R.items, ub + (size_t)R.first_item * n));
    if (G > 1) { NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[3], 0)); NTTB200_CHECK(cudaStreamWaitEvent(s->st2, s->ev[3], 0)); }
    // 5. last inverse kernel + mod-switch + Delta*m (same stream as the run's front: no cross-stream dependency)
    for (const ShardRun &R : runs) {
        const size_t it = R.first_item;
        TRY(enc_finish_limbs(b, R.stream ? P2 : P, c_shard + R.offset, R.cnt, R.first_limb, R.cnt, R.items, cl + it * 2 * n, (size_t)2 * n, (size_t)n,
                             es + it * 2 * n, m + it * n, (size_t)n));
    }
    NTTB200_CHECK(cudaEventRecord(s->ev[4], s->st2));          // the second stream rejoins the caller's
    NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[4], 0));
    return 0;
}

// ---- limb-sharded decryption --------------------------------------------------------------------------------------------------------
// c_shard is overwritten (as c1 is in the reference); m_out[batch][n] is complete on EVERY rank when the call's stream work is done.
int nttb200_bfv_decrypt_sharded(nttb200_bfv *b, nttb200_comm *comm, nttb200_u64 *m_out, nttb200_u64 *c_shard, unsigned batch, void *stream)
{
    if (!b || !comm || !c_shard || !m_out || !batch || batch % (unsigned)comm->world || !b->sk_l) return NTTB200_EINVAL;
    const unsigned G = (unsigned)comm->world, g = (unsigned)comm->rank, n = b->n, r = b->r, rp = r - 1, per = batch / G;
    if (per > 65535) return NTTB200_EINVAL;
    if (G > 1 && (~0ull / b->gamma) < G) return NTTB200_EINVAL;          // the 64-bit SUM of G partial sums below gamma must not wrap
    const bool packed = (u64)rp * (b->t - 1) < 65536, out16 = b->t <= 65536;
    const size_t pw = packed ? (size_t)n + n / 4 : (size_t)2 * n;        // words per item of partial sums
    cudaStream_t st = (cudaStream_t)stream;
    nttb200_shard_state *s;
    TRY(shard_state(b, &s, 1));
    std::vector<nttb200_shard_block> blk(G);
    plan_blocks(rp, n, batch, G, g, blk.data(), nullptr);
    // Default schedule (mode 4, chunks 0) by the size of the call: large batches push their sums with the copy engines in two rounds;
    // a small one (a block below 2^24 coefficient-limbs, e.g. 64 ciphertexts on 8 GPUs) is latency-bound, and one round of direct
    // stores by the kernel is the shortest chain (measured at batch 64, 8 GPUs: 0.30 ms against 0.40).
    int mode = s->mode;
    if (mode == 4 && !s->chunks && (size_t)per * n * rp < ((size_t)1 << 24)) mode = 3;
    unsigned chunks = G > 1 ? (s->chunks ? s->chunks : (mode == 4 ? 2u : 4u)) : 1u;
    while (chunks > 1 && per % chunks) chunks--;
    TRY(shard_state(b, &s, (size_t)3 * G * chunks + 8));
    const unsigned sub = per / chunks;                                   // items of one block in one chunk
    u64 *partial, *recv; unsigned short *plain;
    TRY(shard_buf(s, kBufPartial, (size_t)batch * pw * 8, (void **)&partial));
    TRY(shard_buf(s, kBufRecv, (size_t)per * pw * 8, (void **)&recv));
    TRY(shard_buf(s, kBufPlain, (size_t)batch * n * 2, (void **)&plain));
    Pipe P = pipe_from_bfv(b, st), P2 = pipe_from_bfv(b, s->st2);
    NTTB200_CHECK(cudaEventRecord(s->ev[0], s->cs));                     // scratch reuse across calls (see encrypt)
    NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[0], 0));
    NTTB200_CHECK(cudaEventRecord(s->ev[1], st));                        // the second compute stream joins (c_shard is ready at this point)
    NTTB200_CHECK(cudaStreamWaitEvent(s->st2, s->ev[1], 0));
    size_t evi = 2;
    const size_t own = (size_t)g * per;
    if (G == 1) {          // one rank, one block (two halves on two streams were measured 3 % SLOWER here: decryption's epilogue is multiplier-bound, not HBM-bound)
        TRY(dec_partial(b, P, partial, packed, c_shard, blk[0].limb_count, blk[0].first_limb, blk[0].limb_count, batch));
        TRY(dec_finish(b, m_out, 0, partial, packed, batch, st));
        NTTB200_CHECK(cudaEventRecord(s->ev[evi], s->st2));
        NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[evi], 0));
        return 0;
    }
    const bool p2p_mode = mode >= 2;          // 2 / 3: peer stores by the kernel; 4: local sums pushed by the copy engines
    if (p2p_mode && !s->p2p_failed && !comm->fake) {
        // Peer-to-peer: the partial-sum kernel of a tile writes straight into slot `rank` of the buffer of the items' OWNER, mapped
        // here through CUDA IPC -- NVLink stores issued by the kernel that produces the sums: compute and transfer are one kernel, no
        // collective kernel competes for SMs, no staging copy.  A 4-byte all-reduce is the "everybody has deposited" barrier; the
        // owner then sums the world slots while rounding (k_decrypt_finish) and the 16-bit plaintext words are all-gathered.
        const int rc = sym_setup(s, comm, kSymSlots, (size_t)batch * pw * 8);
        if (rc) return rc;
    }
    if (p2p_mode && (comm->fake || !s->p2p_failed)) {
        // Every rank visits the owners in ROTATED order (rank g starts with owner g + 1): at any moment each owner receives from one
        // sender -- a balanced all-to-all; in lock-step order all 7 peers would store into the same GPU at once (measured: 12.8 ms
        // instead of 9.8 for 4096 ciphertexts on 8 GPUs).
        // mode 2: `chunks` rounds, round c covering piece c of EVERY owner's block, so the owners round and gather piece c while the
        //         next round's transforms run; mode 3: one round (largest launches, everything after the last transform is exposed).
        const unsigned rounds = mode == 3 ? 1u : chunks;
        const unsigned piece = per / rounds;
        // (Running all transforms first -- same-window blocks merged into large launches -- and depositing afterwards was measured
        // SLOWER, 8.8 ms against 7.5: the peer stores then come in one burst with nothing to overlap them.  Tile by tile, the
        // deposits of one tile travel under the next tile's transforms.)
        for (unsigned c = 0; c < rounds; c++) {
            for (unsigned tau = 0; tau < G; tau++) {
                const unsigned j = (g + 1 + tau) % G;
                const unsigned cnt = blk[j].limb_count;
                // my slot at owner j, piece c (profiling with a fake communicator: the same volume into local scratch)
                u64 *dst = comm->fake ? partial + ((size_t)j * per + (size_t)c * piece) * pw
                                      : (u64 *)s->sym[kSymSlots].peer[j] + ((size_t)g * per + (size_t)c * piece) * pw;
                const bool alt = (tau & 1) != 0;
                const bool ce = mode == 4 && !comm->fake && j != g;          // sums into local scratch, pushed by a copy engine afterwards
                u64 *loc = partial + ((size_t)j * per + (size_t)c * piece) * pw;
                if (cnt) TRY(dec_partial(b, alt ? P2 : P, ce ? loc : dst, packed, c_shard + blk[j].offset + (size_t)c * piece * 2 * cnt * n, cnt, blk[j].first_limb, cnt, piece));
                else NTTB200_CHECK(cudaMemsetAsync(ce ? loc : dst, 0, (size_t)piece * pw * 8, alt ? s->st2 : st));
                if (ce) {
                    NTTB200_CHECK(cudaEventRecord(s->ev[evi], alt ? s->st2 : st));
                    NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[evi], 0));
                    evi++;
                    NTTB200_CHECK(cudaMemcpyAsync(dst, loc, (size_t)piece * pw * 8, cudaMemcpyDeviceToDevice, s->cs));
                }
            }
            for (cudaStream_t cst : {st, s->st2}) {              // both compute streams have deposited round c
                NTTB200_CHECK(cudaEventRecord(s->ev[evi], cst));
                NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[evi], 0));
                evi++;
            }
            BARRIER();                                                                            // round c is deposited everywhere
            const u64 *mine = (comm->fake ? partial : (const u64 *)s->sym[kSymSlots].local) + (size_t)c * piece * pw;     // piece c of every slot of my buffer
            if (out16) {
                // staging is round-major: plain[c][owner][piece][n]
                unsigned short *stg = plain + (size_t)c * G * piece * n;
                TRY(dec_finish(b, stg + (size_t)g * piece * n, 1, mine, packed, piece, s->cs, G, (size_t)per * pw));
                COLL(nccl().AllGather(stg + (size_t)g * piece * n, stg, (size_t)piece * n * 2, ncclInt8, comm->comm, s->cs));
                TRY(dec_expand16(stg, m_out + (size_t)c * piece * n, (size_t)piece * n, s->cs, G, (size_t)per * n));
            } else {
                TRY(dec_finish(b, m_out + (own + (size_t)c * piece) * n, 0, mine, packed, piece, s->cs, G, (size_t)per * pw));
                for (unsigned j = 0; j < G; j++)
                    COLL(nccl().Broadcast(m_out + ((size_t)j * per + (size_t)c * piece) * n, m_out + ((size_t)j * per + (size_t)c * piece) * n,
                                               (size_t)piece * n, ncclUint64, (int)j, comm->comm, s->cs));
            }
        }
    } else if (mode != 1) {
        // Block by block, each in `chunks` pieces: transforms + partial sums of a piece on the caller's stream; on the comm stream,
        // behind them, the piece's sums go to the block's owner (ncclReduce), the owner rounds it, and when a block is complete its
        // owner broadcasts the 16-bit plaintext words -- so the only exposed communication is the LAST piece's reduce + broadcast.
        for (unsigned j = 0; j < G; j++) {
            const unsigned cnt = blk[j].limb_count;
            for (unsigned c = 0; c < chunks; c++) {
                u64 *pj = partial + ((size_t)j * per + (size_t)c * sub) * pw;
                if (cnt) TRY(dec_partial(b, P, pj, packed, c_shard + blk[j].offset + (size_t)c * sub * 2 * cnt * n, cnt, blk[j].first_limb, cnt, sub));
                else NTTB200_CHECK(cudaMemsetAsync(pj, 0, (size_t)sub * pw * 8, st));
                NTTB200_CHECK(cudaEventRecord(s->ev[evi], st));
                NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[evi], 0));
                evi++;
                COLL(nccl().Reduce(pj, recv + (size_t)c * sub * pw, (size_t)sub * pw, ncclUint64, ncclSum, (int)j, comm->comm, s->cs));
                if (j == g) {
                    if (out16) TRY(dec_finish(b, plain + (own + (size_t)c * sub) * n, 1, recv + (size_t)c * sub * pw, packed, sub, s->cs));
                    else TRY(dec_finish(b, m_out + (own + (size_t)c * sub) * n, 0, recv + (size_t)c * sub * pw, packed, sub, s->cs));
                }
            }
            const size_t it = (size_t)j * per;
            if (out16) {
                COLL(nccl().Broadcast(plain + it * n, plain + it * n, (size_t)per * n * 2, ncclInt8, (int)j, comm->comm, s->cs));
                TRY(dec_expand16(plain + it * n, m_out + it * n, (size_t)per * n, s->cs));
            } else {
                COLL(nccl().Broadcast(m_out + it * n, m_out + it * n, (size_t)per * n, ncclUint64, (int)j, comm->comm, s->cs));
            }
        }
    } else {
        // mode 1: partial layout [chunk][block][sub items][pw], one ncclReduceScatter per chunk, rounding and ONE all-gather at the end
        for (unsigned c = 0; c < chunks; c++) {
            for (unsigned j = 0; j < G; j++) {
                const unsigned cnt = blk[j].limb_count;
                u64 *pj = partial + ((size_t)c * G + j) * sub * pw;
                if (cnt) TRY(dec_partial(b, P, pj, packed, c_shard + blk[j].offset + (size_t)c * sub * 2 * cnt * n, cnt, blk[j].first_limb, cnt, sub));
                else NTTB200_CHECK(cudaMemsetAsync(pj, 0, (size_t)sub * pw * 8, st));
            }
            NTTB200_CHECK(cudaEventRecord(s->ev[evi], st));
            NTTB200_CHECK(cudaStreamWaitEvent(s->cs, s->ev[evi], 0));
            evi++;
            COLL(nccl().ReduceScatter(partial + (size_t)c * G * sub * pw, recv + (size_t)c * sub * pw, (size_t)sub * pw, ncclUint64, ncclSum,
                                           comm->comm, s->cs));
        }
        if (out16) {
            TRY(dec_finish(b, plain + own * n, 1, recv, packed, per, s->cs));
            COLL(nccl().AllGather(plain + own * n, plain, (size_t)per * n * 2, ncclInt8, comm->comm, s->cs));
            TRY(dec_expand16(plain, m_out, (size_t)batch * n, s->cs));
        } else {
            TRY(dec_finish(b, m_out + own * n, 0, recv, packed, per, s->cs));
            COLL(nccl().AllGather(m_out + own * n, m_out, (size_t)per * n, ncclUint64, comm->comm, s->cs));
        }
    }
    NTTB200_CHECK(cudaEventRecord(s->ev[evi], s->cs));
    NTTB200_CHECK(cudaStreamWaitEvent(st, s->ev[evi], 0));
    return 0;
}

}  // extern "C"
