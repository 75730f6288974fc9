// One FibonacciSq proof sharded over several GPUs (BASELINE configs[4]), behind the C ABI:
//   zk_shard_comm_init (NCCL, one process per GPU) / zk_shard_comm_init_local (G virtual ranks on one device),
//   zk_prove_fibsq_sharded.
// The reference (prover.rs:9-293) has no notion of this; the transcript is byte-identical to the single-GPU one.
//
// Ownership.  Ranks first split the LDE by COSET -- f_eval[8 j + r] = NTT_n(c_k (5 h^r)^k)[j] are eight independent
// size-n transforms, so there is no communication inside a transform -- and the last pass of every transform stores
// straight into the send layout of ONE all-to-all that re-owns the evaluations blocked-cyclically:
//   global index i -> rank (i >> log_block) % G, local block (i >> log_block) / G.
// With that ownership
//   * every FRI fold pairs i with i + len/2 on the same rank: no exchange per layer;
//   * the bottom log_block levels of every Merkle tree are rank-local.  Only the block roots travel (all-gather of
//     len >> log_block digests); the small upper tree and the Fiat-Shamir step run replicated on every rank's device,
//     so all ranks draw identical challenges without a host round trip;
//   * the composition polynomial needs f at i, i+8, i+16: the 16-value head of the NEXT global block, which is a block
//     of rank + 1 -- one ring send/receive of 64 bytes per block.
// When a layer is small it is all-gathered and the rest of FRI runs replicated (a sharded layer costs more in launch
// and collective latency than hashing it whole).  Decommitment: every rank writes the values and sibling digests it
// owns into its transcript buffer; a byte-wise MAX all-reduce completes it everywhere.
//
// Everything is enqueued on the context stream -- kernels and NCCL collectives alike -- and the host synchronises ONCE,
// for the transcript.  All buffers are persistent per (communicator, size).
//
// NCCL is bound at run time (dlopen of libnccl.so.2, preferring an instance the process has already loaded, e.g.
// PyTorch's): single-GPU users of the library do not need it.
#include <dlfcn.h>
#include <string.h>

#include <string>
#include <vector>

#include "ctx.hpp"

using namespace zk;
using namespace zkh;

// ---------------------------------------------------------------- NCCL, bound at run time
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;   // 0 = ncclSuccess
enum { ncclUint8 = 1, ncclUint32 = 3 };   // nccl.h ncclDataType_t
enum { ncclMax = 2 };                     // nccl.h ncclRedOp_t

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string err;
};

NcclApi* nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    const char* env = getenv("ZK_NCCL_LIB");
    void* h = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the instance already in the process (PyTorch's)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return nullptr; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) { ok = false; api.err = std::string("libnccl lacks ") + n; } return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.Send = (decltype(api.Send))sym("ncclSend");
    api.Recv = (decltype(api.Recv))sym("ncclRecv");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return nullptr;
    api.lib = h;
    return &api;
}

// ---------------------------------------------------------------- small data-movement kernels
// The evaluations this rank owns, gathered from the G ranks that computed them: src.p[g] points at the [c][bl][tt] chunk
// source rank g produced for this rank (its c-th coset transform) -- in g's SEND buffer, read over NVLink through a peer
// mapping (fused exchange + unpack), or in the local receive buffer after an all-to-all -- and becomes the rank-local
// array in blocked-cyclic order: local[(bl jb + tt) 8 + (c G + g)].  One thread per point (bl, tt): 8 coalesced reads (a
// warp reads 128 contiguous bytes of each stream), 32 bytes out.
struct UnpackSrc { const uint32_t* p[8]; };
__global__ void __launch_bounds__(256) shard_unpack_kernel(UnpackSrc src, uint32_t* __restrict__ local,
                                                           uint32_t log_per_stream /* log2(nbl jb) */, uint32_t G)
{
    pdl_enter();
    const uint32_t q = blockIdx.x * 256 + threadIdx.x;
    if (q >> log_per_stream) return;
    uint32_t v[8];
#pragma unroll
    for (uint32_t s = 0; s < 8; s++) {
        const uint32_t c = s / G, g = s % G;
        v[s] = src.p[g][((size_t)c << log_per_stream) + q];
    }
    uint4* o = reinterpret_cast<uint4*>(local + (size_t)q * 8);
    o[0] = make_uint4(v[0], v[1], v[2], v[3]);
    o[1] = make_uint4(v[4], v[5], v[6], v[7]);
}

// heads[bl][t] = local[(bl << log_block) + t], t < 16: what the rank that owns the PRECEDING global blocks needs for
// f(g x), f(g^2 x) at the end of its blocks (prover.rs:134-135)
__global__ void __launch_bounds__(256) shard_heads_kernel(const uint32_t* __restrict__ local, uint32_t* __restrict__ heads,
                                                          uint32_t n_blocks, uint32_t log_block)
{
    pdl_enter();
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_blocks * 16) return;
    heads[i] = local[((size_t)(i >> 4) << log_block) + (i & 15)];
}

// [g][bl][unit] -> [bl][g][unit] in 16-byte words (per-rank pieces -> global block order b = bl G + g); `unit` 16-byte words
// per block: 2 for a digest, 2^log_block / 4 for a block of values.  src.p[g] is rank g's piece: a slice of an all-gather
// result, or rank g's own buffer read over NVLink through its peer mapping (fused gather + re-order).
struct PeerSrc { const uint4* p[8]; };
__global__ void __launch_bounds__(256) shard_interleave_kernel(PeerSrc src, uint4* __restrict__ dst, uint32_t G, uint32_t n_bl,
                                                               uint32_t unit)
{
    pdl_enter();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;   // index into dst
    const size_t total = (size_t)G * n_bl * unit;
    if (i >= total) return;
    const uint32_t u = (uint32_t)(i % unit);
    const size_t b = i / unit;
    const uint32_t g = (uint32_t)(b % G), bl = (uint32_t)(b / G);
    dst[i] = src.p[g][(size_t)bl * unit + u];
}

// Barrier between the ranks of a node through peer memory: rank `me` raises flag[me] = epoch in every rank's flag array
// (system-scope release stores over NVLink) and waits until its own array shows `epoch` from everybody.  The kernel sits
// in the stream behind the work whose results the peers are about to read (griddepcontrol.wait: that work is complete and
// its writes are performed) and in front of the kernel that reads the peers' results.  ~4 us instead of the ~15-20 us of a
// one-word NCCL collective.  A peer that never arrives (a failed process) trips a ~15-second time-out, reported to the host.
struct FlagPeers { uint32_t* p[8]; };
__global__ void shard_barrier_kernel(FlagPeers peers, uint32_t* mine, uint32_t me, uint32_t G, uint32_t epoch, uint32_t* err)
{
    pdl_wait();
    const uint32_t t = threadIdx.x;
    if (t < G) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peers.p[t] + me), "r"(epoch) : "memory");
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine + t) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (clock64() - t0 > 30000000000ll) { *err = 1u; break; }   // ~15 s of SM clock: a peer that died, not one that is late
        }
    }
    __syncthreads();
    pdl_trigger();
}

// virtual ranks: byte-wise maximum over G buffers, written back to all of them
struct MaxPtrs { uint32_t* p[8]; };
__global__ void __launch_bounds__(256) shard_max_kernel(MaxPtrs ptrs, uint32_t G, uint32_t n_words)
{
    pdl_enter();
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n_words) return;
    uint32_t m = 0;
    for (uint32_t g = 0; g < G; g++) m = __vmaxu4(m, ptrs.p[g][i]);
    for (uint32_t g = 0; g < G; g++) ptrs.p[g][i] = m;
}

// One (real or virtual) rank's persistent buffers for one proof size.
struct RankBufs {
    uint32_t rank = 0;
    DevBuf state, transcript, y, coeffs, work, send, recv, f_local, layers, heap_f, heap_layers, upper, stage, heads, halo,
        tail_vals, tail_heaps, cp_tab, slice;
    uint32_t cp_tab_log_n = 0, cp_tab_lb = 0;   // the geometry cp_tab was built for
};

enum Phase { PH_START = 0, PH_INTERP, PH_LDE, PH_A2A, PH_UNPACK, PH_FTREE, PH_HALO, PH_CPTREE, PH_FRI, PH_TAIL, PH_DECOMMIT, PH_END, PH_COUNT };

}  // namespace

struct zk_shard_comm {
    zk_ctx* ctx = nullptr;
    uint32_t G = 1;
    bool local = false;
    ncclComm_t nccl = nullptr;
    std::vector<RankBufs> rk;      // one entry (real rank) or G entries (virtual ranks)
    uint32_t log_block = 0;        // 0: chosen per proof -- 2^13 points, or half the rank's share when that is smaller
    uint64_t tail_len = (uint64_t)1 << 18;   // layers of at most this many points finish in the replicated tail
    uint64_t comm_bytes = 0;       // bytes this rank received over the interconnect for the last proof
    uint32_t first_tail = 0;
    cudaEvent_t ev[PH_COUNT] = {};
    float phase_ms[16] = {};
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    // Peer memory (cudaIpc over NVLink / NVSwitch): the buffers the other ranks read from are mapped into this process, so
    // the exchanges are fused into the kernels that consume them -- the unpack reads the evaluations a rank owns straight
    // out of the send buffers of the ranks that computed them, the block roots of every tree and the first tail layer are
    // gathered and re-ordered by one kernel -- with a flag barrier in peer memory in front of each instead of a collective.
    bool p2p_wanted = true;                  // ZK_SHARD_NO_P2P=1: NCCL collectives only
    bool p2p = false;                        // the peers' buffers are mapped (all ranks agree)
    enum { MAP_SEND = 0, MAP_HEAP_F, MAP_HEAP_L, MAP_LAYERS, MAP_Y, MAP_HEADS, MAP_FLAGS, MAP_COUNT };
    void* mapped_for[MAP_COUNT] = {};        // the local buffers the mappings were made for
    uint8_t* peer[MAP_COUNT][8] = {};        // [buffer][rank]: that rank's buffer in this process's address space
    DevBuf flags;                            // [0, 8) barrier flags written by the peers; word 16: time-out flag; word 32: scratch
    DevBuf xchg;                             // staging of the handle exchange
    uint32_t epoch = 0;
    bool broken = false;                     // a collective failed or a peer timed out: no further rendezvous is attempted
};

namespace {

int fail_nccl(zk_shard_comm* cm, ncclResult_t r, const char* what)
{
    NcclApi* A = nccl_api();
    cm->broken = true;
    return fail(cm->ctx, ZK_ERR_COMM, std::string(what) + ": " + (A && A->GetErrorString ? A->GetErrorString(r) : "NCCL error"));
}
#define NK(call)                                                   \
    do {                                                           \
        ncclResult_t r__ = (call);                                 \
        if (r__ != 0) return fail_nccl(cm, r__, #call);            \
    } while (0)

// ---- collectives over the local ranks of this process (one real rank: NCCL; G virtual ranks: device copies)
// every rank holds `chunk` bytes at buf + rank * chunk and ends with all G chunks
int all_gather_inplace(zk_shard_comm* cm, std::vector<uint8_t*>& bufs, size_t chunk)
{
    zk_ctx* ctx = cm->ctx;
    cudaStream_t s = ctx->stream;
    if (cm->G == 1) return ZK_OK;
    if (!cm->local) {
        NK(nccl_api()->AllGather(bufs[0] + cm->rk[0].rank * chunk, bufs[0], chunk, ncclUint8, cm->nccl, s));
    } else {
        for (uint32_t d = 0; d < cm->G; d++)
            for (uint32_t src = 0; src < cm->G; src++)
                if (src != d) CK(cudaMemcpyAsync(bufs[d] + src * chunk, bufs[src] + src * chunk, chunk, cudaMemcpyDeviceToDevice, s));
    }
    cm->comm_bytes += chunk * (cm->G - 1);
    return ZK_OK;
}
// send[r]: `chunk` bytes -> recv[r]: G chunks, rank-major
int all_gather(zk_shard_comm* cm, const std::vector<const uint8_t*>& send, std::vector<uint8_t*>& recv, size_t chunk)
{
    zk_ctx* ctx = cm->ctx;
    cudaStream_t s = ctx->stream;
    if (!cm->local && cm->G > 1) {
        NK(nccl_api()->AllGather(send[0], recv[0], chunk, ncclUint8, cm->nccl, s));
    } else {
        for (uint32_t d = 0; d < cm->G; d++)
            for (uint32_t src = 0; src < cm->G; src++)
                CK(cudaMemcpyAsync(recv[d] + src * chunk, send[src], chunk, cudaMemcpyDeviceToDevice, s));
    }
    cm->comm_bytes += chunk * (cm->G - 1);
    return ZK_OK;
}
// send[r] = G chunks by destination -> recv[r] = G chunks by source
int all_to_all(zk_shard_comm* cm, const std::vector<const uint8_t*>& send, std::vector<uint8_t*>& recv, size_t chunk)
{
    zk_ctx* ctx = cm->ctx;
    cudaStream_t s = ctx->stream;
    if (!cm->local && cm->G > 1) {
        NcclApi* A = nccl_api();
        const uint32_t me = cm->rk[0].rank;
        NK(A->GroupStart());
        for (uint32_t p = 0; p < cm->G; p++) {
            if (p == me) continue;
            NK(A->Send(send[0] + p * chunk, chunk, ncclUint8, (int)p, cm->nccl, s));
            NK(A->Recv(recv[0] + p * chunk, chunk, ncclUint8, (int)p, cm->nccl, s));
        }
        NK(A->GroupEnd());
        CK(cudaMemcpyAsync(recv[0] + me * chunk, send[0] + me * chunk, chunk, cudaMemcpyDeviceToDevice, s));
    } else {
        for (uint32_t d = 0; d < cm->G; d++)
            for (uint32_t src = 0; src < cm->G; src++)
                CK(cudaMemcpyAsync(recv[d] + src * chunk, send[src] + d * chunk, chunk, cudaMemcpyDeviceToDevice, s));
    }
    cm->comm_bytes += chunk * (cm->G - 1);
    return ZK_OK;
}
// every rank receives rank + 1's buffer (ring)
int ring_from_next(zk_shard_comm* cm, const std::vector<const uint8_t*>& send, std::vector<uint8_t*>& recv, size_t bytes)
{
    zk_ctx* ctx = cm->ctx;
    cudaStream_t s = ctx->stream;
    if (!cm->local && cm->G > 1) {
        NcclApi* A = nccl_api();
        const uint32_t me = cm->rk[0].rank, G = cm->G;
        NK(A->GroupStart());
        NK(A->Send(send[0], bytes, ncclUint8, (int)((me + G - 1) % G), cm->nccl, s));
        NK(A->Recv(recv[0], bytes, ncclUint8, (int)((me + 1) % G), cm->nccl, s));
        NK(A->GroupEnd());
        cm->comm_bytes += bytes;
    } else {
        for (uint32_t d = 0; d < cm->G; d++)
            CK(cudaMemcpyAsync(recv[d], send[(d + 1) % cm->G], bytes, cudaMemcpyDeviceToDevice, s));
        if (cm->G > 1) cm->comm_bytes += bytes;
    }
    return ZK_OK;
}
int all_reduce_max_bytes(zk_shard_comm* cm, std::vector<uint8_t*>& bufs, size_t bytes /* multiple of 4 */)
{
    zk_ctx* ctx = cm->ctx;
    cudaStream_t s = ctx->stream;
    if (cm->G == 1) return ZK_OK;
    if (!cm->local) {
        NK(nccl_api()->AllReduce(bufs[0], bufs[0], bytes, ncclUint8, ncclMax, cm->nccl, s));
    } else {
        MaxPtrs mp{};
        for (uint32_t g = 0; g < cm->G; g++) mp.p[g] = (uint32_t*)bufs[g];
        const uint32_t nw = (uint32_t)(bytes / 4);
        CK(zk_launch(shard_max_kernel, dim3((nw + 255) / 256), dim3(256), 0, s, mp, cm->G, nw));
        ctx->launches++;
    }
    cm->comm_bytes += bytes;
    return ZK_OK;
}

struct Geometry {
    uint32_t k, log_N, lb, G, log_G, C;
    size_t n, N, Nloc, nbl, nbt0, tl, trs;
};

size_t local_vals_off(const Geometry& g, uint32_t i) { return 2 * g.Nloc - 2 * (g.Nloc >> i); }      // sum_{j<i} Nloc >> j
size_t local_heap_off(const Geometry& g, uint32_t i) { return 64 * local_vals_off(g, i); }
// upper heaps: tree 0 = f, tree 1 + i = layer i; tree t covers nbt0 >> max(t - 1, 0) block roots
size_t upper_off(const Geometry& g, uint32_t t)
{
    size_t off = 0;
    for (uint32_t u = 0; u < t; u++) off += 64 * (g.nbt0 >> (u ? u - 1 : 0));
    return off;
}

void close_peers(zk_shard_comm* cm)
{
    for (int m = 0; m < zk_shard_comm::MAP_COUNT; m++) {
        for (uint32_t p = 0; p < cm->G; p++) {
            if (cm->peer[m][p] && !cm->local && p != cm->rk[0].rank) cudaIpcCloseMemHandle(cm->peer[m][p]);
            cm->peer[m][p] = nullptr;
        }
        cm->mapped_for[m] = nullptr;
    }
    cm->p2p = false;
}

// Exported memory must not be freed while a peer still has it mapped (cudaIpcOpenMemHandle: undefined behaviour).  Before
// a mapped buffer is re-allocated or released every rank therefore closes its mappings and all ranks meet (a one-byte
// all-reduce + stream synchronisation): only then may anybody free.  Collective; skipped on a communicator that has
// already seen a failure (its peers may be gone, and waiting for them would hang).
int quiesce_peers(zk_shard_comm* cm)
{
    if (cm->local || cm->G == 1 || !cm->p2p) return ZK_OK;
    zk_ctx* ctx = cm->ctx;
    CK(cudaStreamSynchronize(ctx->stream));
    close_peers(cm);
    if (cm->broken || !cm->nccl || !cm->xchg.p) return ZK_OK;
    NK(nccl_api()->AllReduce(cm->xchg.p, cm->xchg.p, 1, ncclUint8, ncclMax, cm->nccl, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// (Re)maps the peers' buffers after they were (re)allocated.  Collective: every rank calls it at the same point (buffer
// sizes depend on the geometry only).  Any rank that cannot map a peer makes ALL ranks fall back to NCCL collectives.
int setup_peers(zk_shard_comm* cm)
{
    zk_ctx* ctx = cm->ctx;
    if (cm->G == 1 || !cm->p2p_wanted) { cm->p2p = false; return ZK_OK; }
    if (cm->local) {   // virtual ranks: the "peers" are buffers of the same device, the single stream is the barrier
        cm->p2p = true;
        return ZK_OK;
    }
    constexpr int NM = zk_shard_comm::MAP_COUNT;
    RankBufs& b = cm->rk[0];
    CK(cm->flags.ensure(256));
    CK(cm->xchg.ensure(64 * NM * 8 + 64));
    void* bases[NM] = {b.send.p, b.heap_f.p, b.heap_layers.p, b.layers.p, b.slice.p, b.heads.p, cm->flags.p};
    bool same = cm->p2p;
    for (int m = 0; m < NM; m++) same = same && cm->mapped_for[m] == bases[m];
    if (same) return ZK_OK;
    close_peers(cm);
    cudaStream_t s = ctx->stream;
    const uint32_t me = b.rank, G = cm->G;
    // the flags start at zero and the epoch with them; the exchange below orders this against every peer's first signal
    CK(cudaMemsetAsync(cm->flags.p, 0, 256, s));
    cm->epoch = 0;
    cudaIpcMemHandle_t h[NM];
    unsigned char fail_flag = 0;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
    for (int m = 0; m < NM; m++)
        if (cudaIpcGetMemHandle(&h[m], bases[m]) != cudaSuccess) { cudaGetLastError(); fail_flag = 1; memset(&h[m], 0, 64); }
    uint8_t* st = (uint8_t*)cm->xchg.p;
    const size_t per = 64 * NM;
    CK(cudaMemcpyAsync(st + per * me, h, per, cudaMemcpyHostToDevice, s));
    NK(nccl_api()->AllGather(st + per * me, st, per, ncclUint8, cm->nccl, s));
    std::vector<cudaIpcMemHandle_t> all((size_t)NM * G);
    CK(cudaMemcpyAsync(all.data(), st, per * G, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint32_t p = 0; p < G && !fail_flag; p++) {
        for (int m = 0; m < NM && !fail_flag; m++) {
            if (p == me) { cm->peer[m][p] = (uint8_t*)bases[m]; continue; }
            void* ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, all[(size_t)p * NM + m], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                fail_flag = 1;
                break;
            }
            cm->peer[m][p] = (uint8_t*)ptr;
        }
    }
    // agree: one failure anywhere means nobody uses the mappings
    CK(cudaMemcpyAsync(st + per * G, &fail_flag, 1, cudaMemcpyHostToDevice, s));
    NK(nccl_api()->AllReduce(st + per * G, st + per * G, 1, ncclUint8, ncclMax, cm->nccl, s));
    unsigned char any = 0;
    CK(cudaMemcpyAsync(&any, st + per * G, 1, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (any) {   // agreed by all ranks: this communicator stays on NCCL collectives (no retry per proof)
        close_peers(cm);
        cm->p2p_wanted = false;
        return ZK_OK;
    }
    cm->p2p = true;
    for (int m = 0; m < NM; m++) cm->mapped_for[m] = bases[m];
    return ZK_OK;
}

// all ranks have finished what is in front of this point of their streams (peer mode; virtual ranks: one stream, nothing to do)
int peer_barrier(zk_shard_comm* cm)
{
    zk_ctx* ctx = cm->ctx;
    if (cm->local || cm->G == 1) return ZK_OK;
    FlagPeers fp{};
    for (uint32_t p = 0; p < cm->G; p++) fp.p[p] = (uint32_t*)cm->peer[zk_shard_comm::MAP_FLAGS][p];
    cm->epoch++;
    CK(zk_launch(shard_barrier_kernel, dim3(1), dim3(32), 0, ctx->stream, fp, (uint32_t*)cm->flags.p, cm->rk[0].rank, cm->G,
                 cm->epoch, (uint32_t*)cm->flags.p + 16));
    ctx->launches++;
    return ZK_OK;
}
// rank r's buffer `map` as local rank `i` sees it (peer mode)
const uint8_t* peer_ptr(zk_shard_comm* cm, int map, uint32_t r, DevBuf RankBufs::*member)
{
    return cm->local ? (const uint8_t*)(cm->rk[r].*member).p : cm->peer[map][r];
}

int ensure_bufs(zk_shard_comm* cm, const Geometry& g, size_t tail_first_len)
{
    zk_ctx* ctx = cm->ctx;
    {
        // the buffers the peers have mapped: if one of them is about to move, un-map everywhere first (every rank sees the
        // same sizes and the same call history, so all of them take this branch together)
        const RankBufs& b0 = cm->rk[0];
        const bool moves = b0.send.bytes < 4 * g.Nloc || b0.layers.bytes < 8 * g.Nloc || b0.heap_f.bytes < 64 * g.Nloc ||
                           b0.heap_layers.bytes < 128 * g.Nloc || b0.slice.bytes < 4 * (g.n / g.G) || b0.heads.bytes < 64 * g.nbl;
        if (moves) {
            int rc = quiesce_peers(cm);
            if (rc) return rc;
        }
    }
    for (RankBufs& b : cm->rk) {
        CK(b.state.ensure(sizeof(ProofState)));
        CK(b.transcript.ensure(g.trs + 64));
        CK(b.y.ensure(4 * g.n));
        CK(b.slice.ensure(4 * (g.n / g.G)));       // this rank's 1/G of the trace as uploaded: what the peers pull
        CK(b.coeffs.ensure(4 * g.n));
        CK(b.work.ensure(4 * g.n));
        CK(b.send.ensure(4 * g.Nloc));
        CK(b.recv.ensure(4 * g.Nloc));
        CK(b.f_local.ensure(4 * g.Nloc));
        CK(b.layers.ensure(8 * g.Nloc));
        CK(b.heap_f.ensure(64 * g.Nloc));
        CK(b.heap_layers.ensure(128 * g.Nloc));
        CK(b.upper.ensure(upper_off(g, g.k + 2) + 64));
        const size_t stage = 32 * g.nbt0 > 4 * tail_first_len ? 32 * g.nbt0 : 4 * tail_first_len;
        CK(b.stage.ensure(stage));
        CK(b.heads.ensure(64 * g.nbl));
        CK(b.halo.ensure(64 * g.nbl));
        CK(b.tail_vals.ensure(8 * tail_first_len));
        CK(b.tail_heaps.ensure(128 * tail_first_len));
    }
    const size_t pin = (g.trs + 64 + sizeof(ProofState)) * cm->rk.size();
    if (pin > cm->pinned_bytes) {
        if (cm->pinned) cudaFreeHost(cm->pinned);
        cm->pinned = nullptr;
        cm->pinned_bytes = 0;
        CK(cudaMallocHost(&cm->pinned, pin));
        cm->pinned_bytes = pin;
    }
    return setup_peers(cm);
}

// Distributed Merkle tree of one layer: the rank-local block subtrees are already in the heaps (levels down to the block
// roots; the layer's heap starts `heap_off` bytes into the buffer `member`); gathers the block roots, builds the replicated
// upper tree and runs the channel step `post` on every rank.
int commit_upper(zk_shard_comm* cm, const Geometry& g, uint32_t tree, size_t length, int map, DevBuf RankBufs::*member,
                 size_t heap_off, int post, uint32_t fri_layer)
{
    zk_ctx* ctx = cm->ctx;
    const size_t nbl = (length / g.G) >> g.lb, nbt = length >> g.lb;
    const uint32_t log_nbt = ilog2_u64(nbt);
    const size_t roots_off = heap_off + 32 * (nbl - 1);       // heap level log2(nbl): the block roots
    int rc;
    if (cm->p2p) {
        if ((rc = peer_barrier(cm))) return rc;
        cm->comm_bytes += 32 * nbl * (g.G - 1);
    } else {
        std::vector<const uint8_t*> send;
        std::vector<uint8_t*> recv;
        for (RankBufs& b : cm->rk) {
            send.push_back((const uint8_t*)(b.*member).p + roots_off);
            recv.push_back((uint8_t*)b.stage.p);
        }
        if ((rc = all_gather(cm, send, recv, 32 * nbl))) return rc;
    }
    for (RankBufs& b : cm->rk) {
        uint8_t* upper = (uint8_t*)b.upper.p + upper_off(g, tree);
        PeerSrc ps{};
        for (uint32_t r = 0; r < g.G; r++)
            ps.p[r] = cm->p2p ? (const uint4*)(peer_ptr(cm, map, r, member) + roots_off) : (const uint4*)((const uint8_t*)b.stage.p + 32 * nbl * r);
        const size_t total = nbt * 2;
        CK(zk_launch(shard_interleave_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, ctx->stream, ps,
                     (uint4*)(upper + 32 * (nbt - 1)), g.G, (uint32_t)nbl, 2u));
        ctx->launches++;
        if ((rc = zk_shard_upper_tree_dev(ctx, upper, log_nbt, post, fri_layer, g.k, b.state.p, (uint8_t*)b.transcript.p))) return rc;
    }
    return ZK_OK;
}

void mark(zk_shard_comm* cm, Phase p) { cudaEventRecord(cm->ev[p], cm->ctx->stream); }

int prove_sharded(zk_shard_comm* cm, uint32_t log_n, const uint32_t* trace, uint8_t* transcript_out, uint8_t state_out[32])
{
    zk_ctx* ctx = cm->ctx;
    cudaStream_t s = ctx->stream;
    Geometry g{};
    g.k = log_n; g.log_N = log_n + LOG_BLOWUP; g.G = cm->G; g.log_G = ilog2_u64(cm->G); g.C = 8 / cm->G;
    g.n = (size_t)1 << log_n; g.N = g.n << LOG_BLOWUP; g.Nloc = g.N / g.G;
    // blocks of 2^13 points keep the replicated upper trees three levels shorter than 2^10 (measured, 2^24 on 2 GPUs: -0.11 ms
    // over the trees; 2^14 and larger change nothing more); small proofs take half the rank's share, i.e. two blocks
    g.lb = cm->log_block;
    if (!g.lb) {
        const uint32_t half = ilog2_u64(g.Nloc) > 0 ? ilog2_u64(g.Nloc) - 1 : 0;
        g.lb = half < 13 ? half : 13;
    }
    if (g.lb < 5 || g.Nloc < ((size_t)2 << g.lb))
        return fail(ctx, ZK_ERR_INVALID, "proof too small to shard: every rank needs at least two blocks of 2^log_block >= 32 points");
    g.nbl = g.Nloc >> g.lb; g.nbt0 = g.N >> g.lb; g.tl = transcript_len(log_n); g.trs = (g.tl + 63) & ~(size_t)63;
    const uint32_t L = log_n, nr = (uint32_t)cm->rk.size();
    // which layers are committed distributed: layer i + 1 while layer i has >= 2 blocks per rank and is larger than tail_len
    uint32_t first_tail = 0;
    {
        size_t length = g.N;
        while ((length / g.G) >= ((size_t)2 << g.lb) && length > cm->tail_len && first_tail + 1 < L) { first_tail++; length >>= 1; }
    }
    const size_t tail_first_len = g.N >> first_tail;
    if (first_tail + 2 > SHARD_MAX_TREES) return fail(ctx, ZK_ERR_INVALID, "too many sharded layers");
    int rc = ensure_bufs(cm, g, tail_first_len);
    if (rc) return rc;
    SizeTables* T;
    if ((rc = get_tables(ctx, g.log_N, &T))) return rc;
    cm->comm_bytes = 0;
    cm->first_tail = first_tail;
    const uint32_t a_first = trace[0], a_last = trace[g.n - 2];

    mark(cm, PH_START);
    // ---- Part 1a: trace slices up (each rank 1/G of it over PCIe), all-gather over NVLink, interpolation replicated
    // (1/8 of the transform work).  prover.rs:60-61
    const size_t chunk = g.n / g.G;
    std::vector<uint8_t*> ys;
    for (RankBufs& b : cm->rk) {
        CK(cudaMemsetAsync(b.transcript.p, 0, g.trs + 64, s));
        if ((rc = zk_shard_state_init_dev(ctx, b.state.p, a_first, a_last))) return rc;
        const size_t lo = b.rank * chunk, hi = (lo + chunk < g.n - 1) ? lo + chunk : g.n - 1;
        // peer mode: into the slice buffer (y is transformed in place while a slower peer may still be pulling)
        uint32_t* up = (cm->p2p && g.G > 1) ? (uint32_t*)b.slice.p : (uint32_t*)b.y.p + lo;
        CK(cudaMemcpyAsync(up, trace + lo, 4 * (hi - lo), cudaMemcpyHostToDevice, s));
        if (hi - lo < chunk) CK(cudaMemsetAsync(up + (hi - lo), 0, 4 * (lo + chunk - hi), s));
        ys.push_back((uint8_t*)b.y.p);
    }
    if (cm->p2p && g.G > 1) {
        // every rank pulls the other slices straight out of the peers' buffers (NVLink reads at full width) behind a flag
        // barrier: ~0.07 ms for the 56 MB of a 2^24 trace on 8 GPUs against ~0.2 ms for the NCCL all-gather
        if ((rc = peer_barrier(cm))) return rc;
        for (RankBufs& b : cm->rk) {
            PeerSrc ps{};
            for (uint32_t r = 0; r < g.G; r++)
                ps.p[r] = (const uint4*)peer_ptr(cm, zk_shard_comm::MAP_Y, r, &RankBufs::slice);
            const size_t total = g.n / 4;
            CK(zk_launch(shard_interleave_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, ps, (uint4*)b.y.p, g.G, 1u,
                         (uint32_t)(chunk / 4)));
            ctx->launches++;
        }
        cm->comm_bytes += 4 * chunk * (g.G - 1);
    } else if ((rc = all_gather_inplace(cm, ys, 4 * chunk))) {
        return rc;
    }
    for (RankBufs& b : cm->rk)
        if ((rc = zk_interpolate_subgroup_dev(ctx, (uint32_t*)b.y.p, (uint32_t)(g.n - 1), log_n, (uint32_t*)b.coeffs.p))) return rc;
    mark(cm, PH_INTERP);

    // ---- Part 1b: LDE (prover.rs:69-70), cosets r, r + G, ... on rank r; the last pass of every transform stores into
    // the all-to-all send layout [destination][slot][local block][point]
    {
        NttPack pk{};
        pk.log_jb = g.lb - LOG_BLOWUP;
        pk.log_G = g.log_G;
        pk.log_nbl = ilog2_u64(g.nbl);
        pk.n_slots = g.C;
        for (RankBufs& b : cm->rk) {
            for (uint32_t c = 0; c < g.C; c++) {
                pk.slot = c;
                int nl = 0;
                CK(launch_lde_one_coset(s, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_n, b.rank + c * g.G,
                                        (const uint32_t*)b.coeffs.p, (uint32_t*)b.work.p, (uint32_t*)b.send.p, &nl, &pk));
                ctx->launches += nl;
            }
        }
    }
    mark(cm, PH_LDE);
    const size_t per_dest = g.Nloc / g.G;                     // words every rank produced for every other rank
    if (cm->p2p) {
        // every rank's transforms must be complete before anybody reads them
        if ((rc = peer_barrier(cm))) return rc;
        cm->comm_bytes += 4 * per_dest * (g.G - 1);
    } else {
        std::vector<const uint8_t*> send;
        std::vector<uint8_t*> recv;
        for (RankBufs& b : cm->rk) { send.push_back((const uint8_t*)b.send.p); recv.push_back((uint8_t*)b.recv.p); }
        if ((rc = all_to_all(cm, send, recv, 4 * per_dest))) return rc;
    }
    mark(cm, PH_A2A);
    {
        const uint32_t log_ps = ilog2_u64(g.Nloc >> LOG_BLOWUP);
        for (RankBufs& b : cm->rk) {
            UnpackSrc us{};
            for (uint32_t gs = 0; gs < g.G; gs++) {
                if (cm->p2p) {   // rank gs's send buffer, the chunk it produced for rank b.rank
                    const uint32_t* base = (const uint32_t*)peer_ptr(cm, zk_shard_comm::MAP_SEND, gs, &RankBufs::send);
                    us.p[gs] = base + (size_t)b.rank * per_dest;
                } else {         // after the all-to-all: chunk gs of the local receive buffer
                    us.p[gs] = (const uint32_t*)b.recv.p + (size_t)gs * per_dest;
                }
            }
            CK(zk_launch(shard_unpack_kernel, dim3((unsigned)(((g.Nloc >> LOG_BLOWUP) + 255) / 256)), dim3(256), 0, s, us,
                         (uint32_t*)b.f_local.p, log_ps, g.G));
            ctx->launches++;
        }
    }
    mark(cm, PH_UNPACK);

    // heads of the local blocks: what the rank before this one needs for f(g x), f(g^2 x) at the end of ITS blocks.  Produced
    // in front of the f tree so that, in peer mode, the tree's barrier covers them and the CP kernel reads them in place.
    for (RankBufs& b : cm->rk) {
        CK(zk_launch(shard_heads_kernel, dim3((unsigned)((16 * g.nbl + 255) / 256)), dim3(256), 0, s, (const uint32_t*)b.f_local.p,
                     (uint32_t*)b.heads.p, (uint32_t)g.nbl, g.lb));
        ctx->launches++;
    }
    // ---- f tree (prover.rs:81-85), then alpha_0..2 (prover.rs:163-165) on every rank
    for (RankBufs& b : cm->rk)
        if ((rc = zk_shard_merkle_blocks_dev(ctx, (const uint32_t*)b.f_local.p, g.Nloc, g.lb, (uint8_t*)b.heap_f.p))) return rc;
    if ((rc = commit_upper(cm, g, 0, g.N, zk_shard_comm::MAP_HEAP_F, &RankBufs::heap_f, 0, POST_F, 0))) return rc;
    mark(cm, PH_FTREE);

    // ---- Part 2: composition polynomial (prover.rs:101-180), fused into the leaf pass of the block subtrees
    const bool halo_in_place = cm->p2p && g.G > 1;   // the next rank's heads are read over NVLink by the CP kernel itself
    if (halo_in_place) {
        cm->comm_bytes += 64 * g.nbl;
    } else {
        std::vector<const uint8_t*> send;
        std::vector<uint8_t*> recv;
        for (RankBufs& b : cm->rk) {
            send.push_back((const uint8_t*)b.heads.p);
            recv.push_back((uint8_t*)b.halo.p);
        }
        if ((rc = ring_from_next(cm, send, recv, 64 * g.nbl))) return rc;
    }
    mark(cm, PH_HALO);
    for (RankBufs& b : cm->rk) {
        // the block after local block bl of the LAST rank is rank 0's block bl + 1 (wrapping around the domain)
        const uint32_t roll = (b.rank + 1 == g.G) ? 1u : 0u;
        if (b.cp_tab.p && (b.cp_tab_log_n != log_n || b.cp_tab_lb != g.lb)) b.cp_tab.release();   // another geometry: rebuild
        b.cp_tab_log_n = log_n;
        b.cp_tab_lb = g.lb;
        const uint32_t* halo = halo_in_place ? (const uint32_t*)peer_ptr(cm, zk_shard_comm::MAP_HEADS, (b.rank + 1) % g.G, &RankBufs::heads)
                                             : (const uint32_t*)b.halo.p;
        if ((rc = shard_cp_blocks(ctx, (const uint32_t*)b.f_local.p, halo, roll, g.Nloc, log_n, g.G, b.rank, g.lb,
                                  b.state.p, (uint32_t*)b.layers.p, (uint8_t*)b.heap_layers.p, &b.cp_tab)))
            return rc;
    }
    if ((rc = commit_upper(cm, g, 1, g.N, zk_shard_comm::MAP_HEAP_L, &RankBufs::heap_layers, 0, POST_CP, 0))) return rc;
    mark(cm, PH_CPTREE);

    // ---- Part 3: FRI (prover.rs:192-254).  Folds stay rank-local; fold + next layer's block subtrees are one pass.
    size_t length = g.N;
    for (uint32_t i = 0; i < first_tail; i++) {
        for (RankBufs& b : cm->rk) {
            uint32_t* cur = (uint32_t*)b.layers.p + local_vals_off(g, i);
            uint32_t* nxt = (uint32_t*)b.layers.p + local_vals_off(g, i + 1);
            uint8_t* heap = (uint8_t*)b.heap_layers.p + local_heap_off(g, i + 1);
            if ((rc = zk_shard_fold_blocks_dev(ctx, cur, length / g.G, i, g.log_N, g.G, b.rank, g.lb, b.state.p, nxt, heap))) return rc;
        }
        if ((rc = commit_upper(cm, g, 2 + i, length / 2, zk_shard_comm::MAP_HEAP_L, &RankBufs::heap_layers, local_heap_off(g, i + 1), POST_FRI, i + 1))) return rc;
        length >>= 1;
    }
    mark(cm, PH_FRI);
    // replicated tail: gather layer first_tail and finish on every rank
    {
        const size_t vals_off = 4 * local_vals_off(g, first_tail), piece = 4 * (length / g.G);   // bytes
        if (cm->p2p) {
            if ((rc = peer_barrier(cm))) return rc;
            cm->comm_bytes += piece * (g.G - 1);
        } else {
            std::vector<const uint8_t*> send;
            std::vector<uint8_t*> recv;
            for (RankBufs& b : cm->rk) {
                send.push_back((const uint8_t*)b.layers.p + vals_off);
                recv.push_back((uint8_t*)b.stage.p);
            }
            if ((rc = all_gather(cm, send, recv, piece))) return rc;
        }
        const size_t nblc = (length / g.G) >> g.lb;
        for (RankBufs& b : cm->rk) {
            PeerSrc ps{};
            for (uint32_t r = 0; r < g.G; r++)
                ps.p[r] = cm->p2p ? (const uint4*)(peer_ptr(cm, zk_shard_comm::MAP_LAYERS, r, &RankBufs::layers) + vals_off)
                                  : (const uint4*)((const uint8_t*)b.stage.p + piece * r);
            const size_t total = length / 4;
            CK(zk_launch(shard_interleave_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, s, ps, (uint4*)b.tail_vals.p, g.G,
                         (uint32_t)nblc, (uint32_t)(1u << (g.lb - 2))));
            ctx->launches++;
            if ((rc = zk_shard_fri_tail_dev(ctx, log_n, first_tail, (uint32_t*)b.tail_vals.p, (uint8_t*)b.tail_heaps.p, b.state.p,
                                            (uint8_t*)b.transcript.p)))
                return rc;
        }
    }
    mark(cm, PH_TAIL);

    // ---- Part 4: decommit (prover.rs:263-289)
    std::vector<uint8_t*> trs;
    for (RankBufs& b : cm->rk) {
        std::vector<const void*> ptrs;
        ptrs.push_back(b.f_local.p); ptrs.push_back(b.heap_f.p); ptrs.push_back((uint8_t*)b.upper.p + upper_off(g, 0));
        for (uint32_t i = 0; i <= first_tail; i++) {
            ptrs.push_back((uint32_t*)b.layers.p + local_vals_off(g, i));
            ptrs.push_back((uint8_t*)b.heap_layers.p + local_heap_off(g, i));
            ptrs.push_back((uint8_t*)b.upper.p + upper_off(g, 1 + i));
        }
        if ((rc = zk_shard_decommit_dev(ctx, log_n, g.G, b.rank, g.lb, first_tail, ptrs.data(), (const uint32_t*)b.tail_vals.p,
                                        (const uint8_t*)b.tail_heaps.p, b.state.p, (uint8_t*)b.transcript.p)))
            return rc;
        trs.push_back((uint8_t*)b.transcript.p);
    }
    if ((rc = all_reduce_max_bytes(cm, trs, g.trs + 64))) return rc;
    mark(cm, PH_DECOMMIT);
    uint8_t* pin = (uint8_t*)cm->pinned;
    const size_t per = g.trs + 64 + sizeof(ProofState);
    for (uint32_t i = 0; i < nr; i++) {
        CK(cudaMemcpyAsync(pin + per * i, cm->rk[i].transcript.p, g.trs + 64, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(pin + per * i + g.trs + 64, cm->rk[i].state.p, sizeof(ProofState), cudaMemcpyDeviceToHost, s));
    }
    uint32_t peer_timeout = 0;
    if (cm->p2p && !cm->local) CK(cudaMemcpyAsync(&peer_timeout, (uint32_t*)cm->flags.p + 16, 4, cudaMemcpyDeviceToHost, s));
    mark(cm, PH_END);
    CK(cudaStreamSynchronize(s));
    if (peer_timeout) {
        cudaMemsetAsync((uint32_t*)cm->flags.p + 16, 0, 4, s);
        cm->broken = true;
        return fail(ctx, ZK_ERR_COMM, "a rank did not reach a peer-memory barrier within ~15 s");
    }
    for (int p = 0; p + 1 < PH_COUNT; p++) cudaEventElapsedTime(&cm->phase_ms[p], cm->ev[p], cm->ev[p + 1]);
    cudaEventElapsedTime(&cm->phase_ms[PH_COUNT - 1], cm->ev[PH_START], cm->ev[PH_END]);

    const uint32_t flag_off = (uint32_t)((g.tl + 3) & ~(size_t)3);
    for (uint32_t i = 0; i < nr; i++) {
        const uint8_t* tr = pin + per * i;
        if (*reinterpret_cast<const uint32_t*>(tr + flag_off) != 0)
            return fail(ctx, ZK_ERR_ASSERT, "last FRI layer is not constant (prover.rs:238)");
        if (i && memcmp(tr, pin, g.tl) != 0)
            return fail(ctx, ZK_ERR_ASSERT, "virtual ranks assembled different transcripts");
    }
    ProofState ps;
    memcpy(&ps, pin + g.trs + 64, sizeof ps);
    memcpy(transcript_out, pin, g.tl);
    finish_channel_host(log_n, ps, pin, state_out);      // channel.rs:19-26 over the decommitment records
    return ZK_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int zk_shard_unique_id(uint8_t id_out[ZK_SHARD_ID_BYTES])
{
    if (!id_out) return ZK_ERR_INVALID;
    NcclApi* A = nccl_api();
    if (!A) return ZK_ERR_COMM;
    ncclUniqueId id;
    if (A->GetUniqueId(&id) != 0) return ZK_ERR_COMM;
    static_assert(sizeof(id) == ZK_SHARD_ID_BYTES, "ncclUniqueId size");
    memcpy(id_out, &id, sizeof id);
    return ZK_OK;
}

static int comm_common(zk_ctx* ctx, zk_shard_comm* cm)
{
    for (int i = 0; i < PH_COUNT; i++) CK(cudaEventCreate(&cm->ev[i]));
    if (const char* e = getenv("ZK_SHARD_LOG_BLOCK")) cm->log_block = (uint32_t)atoi(e);
    if (const char* e = getenv("ZK_SHARD_TAIL_LOG")) cm->tail_len = (uint64_t)1 << atoi(e);
    if (const char* e = getenv("ZK_SHARD_NO_P2P")) cm->p2p_wanted = !(e[0] == '1');
    return ZK_OK;
}

int zk_shard_comm_init(zk_ctx* ctx, const uint8_t id[ZK_SHARD_ID_BYTES], uint32_t n_ranks, uint32_t rank, zk_shard_comm** out)
{
    if (!ctx || !out) return ZK_ERR_INVALID;
    *out = nullptr;
    if (!id || rank >= n_ranks || (n_ranks != 1 && n_ranks != 2 && n_ranks != 4 && n_ranks != 8))
        return fail(ctx, ZK_ERR_INVALID, "n_ranks must be 1, 2, 4 or 8 (the LDE has 8 cosets) and rank < n_ranks");
    CK(cudaSetDevice(ctx->device));
    zk_shard_comm* cm = new zk_shard_comm();
    cm->ctx = ctx;
    cm->G = n_ranks;
    cm->rk.resize(1);
    cm->rk[0].rank = rank;
    if (n_ranks > 1) {
        NcclApi* A = nccl_api();
        if (!A) { delete cm; return fail(ctx, ZK_ERR_COMM, "NCCL is not available: libnccl.so.2 could not be loaded"); }
        ncclUniqueId nid;
        memcpy(&nid, id, sizeof nid);
        ncclResult_t r = A->CommInitRank(&cm->nccl, (int)n_ranks, nid, (int)rank);
        if (r != 0) { int rc = fail_nccl(cm, r, "ncclCommInitRank"); delete cm; return rc; }
    }
    int rc = comm_common(ctx, cm);
    if (rc) { zk_shard_comm_destroy(cm); return rc; }
    *out = cm;
    return ZK_OK;
}

int zk_shard_comm_init_local(zk_ctx* ctx, uint32_t n_virtual_ranks, zk_shard_comm** out)
{
    if (!ctx || !out) return ZK_ERR_INVALID;
    *out = nullptr;
    const uint32_t G = n_virtual_ranks;
    if (G != 1 && G != 2 && G != 4 && G != 8) return fail(ctx, ZK_ERR_INVALID, "n_virtual_ranks must be 1, 2, 4 or 8");
    CK(cudaSetDevice(ctx->device));
    zk_shard_comm* cm = new zk_shard_comm();
    cm->ctx = ctx;
    cm->G = G;
    cm->local = true;
    cm->rk.resize(G);
    for (uint32_t r = 0; r < G; r++) cm->rk[r].rank = r;
    int rc = comm_common(ctx, cm);
    if (rc) { zk_shard_comm_destroy(cm); return rc; }
    *out = cm;
    return ZK_OK;
}

void zk_shard_comm_destroy(zk_shard_comm* cm)
{
    if (!cm) return;
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    quiesce_peers(cm);      // real ranks with peer mappings: nobody frees what somebody still maps (collective, like ncclCommDestroy)
    close_peers(cm);
    cm->flags.release();
    cm->xchg.release();
    if (cm->nccl && nccl_api()) nccl_api()->CommDestroy(cm->nccl);
    for (RankBufs& b : cm->rk) {
        DevBuf* bufs[] = {&b.state, &b.transcript, &b.y, &b.coeffs, &b.work, &b.send, &b.recv, &b.f_local, &b.layers, &b.heap_f,
                          &b.heap_layers, &b.upper, &b.stage, &b.heads, &b.halo, &b.tail_vals, &b.tail_heaps, &b.cp_tab, &b.slice};
        for (DevBuf* d : bufs) d->release();
    }
    for (int i = 0; i < PH_COUNT; i++) if (cm->ev[i]) cudaEventDestroy(cm->ev[i]);
    if (cm->pinned) cudaFreeHost(cm->pinned);
    delete cm;
}

int zk_shard_comm_set(zk_shard_comm* cm, uint32_t log_block, uint32_t log_tail_len)
{
    if (!cm) return ZK_ERR_INVALID;
    if (log_block) {
        if (log_block < 5 || log_block > 20) return fail(cm->ctx, ZK_ERR_INVALID, "log_block must be in [5, 20]");
        cm->log_block = log_block;
    }
    if (log_tail_len != 0xffffffffu) cm->tail_len = log_tail_len >= 63 ? ~(uint64_t)0 : (log_tail_len ? (uint64_t)1 << log_tail_len : 0);
    return ZK_OK;
}

int zk_prove_fibsq_sharded(zk_shard_comm* cm, uint32_t log_n, const uint32_t* trace, uint8_t* transcript_out, size_t cap,
                           size_t* len_out, uint8_t state_out[32])
{
    if (!cm) return ZK_ERR_INVALID;
    zk_ctx* ctx = cm->ctx;
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!trace || !transcript_out || !state_out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (cap < transcript_len(log_n)) return fail(ctx, ZK_ERR_BUFFER, "transcript buffer too small");
    CK(cudaSetDevice(ctx->device));
    rc = prove_sharded(cm, log_n, trace, transcript_out, state_out);
    if (rc != ZK_OK) {
        const std::string keep = ctx->err;
        cudaStreamSynchronize(ctx->stream);
        cudaMemset(ctx->counters.p, 0, 12 * MAX_BATCH);
        cudaGetLastError();
        ctx->err = keep;
        return rc;
    }
    if (len_out) *len_out = transcript_len(log_n);
    return ZK_OK;
}

int zk_shard_last_phase_ms(const zk_shard_comm* cm, float out_ms[12])
{
    if (!cm || !out_ms) return ZK_ERR_INVALID;
    for (int i = 0; i < 12; i++) out_ms[i] = cm->phase_ms[i];
    return ZK_OK;
}

uint64_t zk_shard_last_comm_bytes(const zk_shard_comm* cm) { return cm ? cm->comm_bytes : 0; }
uint32_t zk_shard_last_first_tail(const zk_shard_comm* cm) { return cm ? cm->first_tail : 0; }

}  // extern "C"
