// C ABI (include/zkstark_b200.h): context, constant tables, work space, the device-resident
// prove() pipeline (prover.rs:9-293) and the per-function entry points.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cstdlib>
#include <thread>
#include <map>
#include <string>
#include <vector>

#include "ctx.hpp"
#include "sha256.cuh"


using namespace zk;
using namespace zkh;

namespace zkh {

int fail(zk_ctx* c, int code, const std::string& msg)
{
    if (c) c->err = msg;
    return code;
}
int fail_cuda(zk_ctx* c, cudaError_t e, const char* what)
{
    cudaGetLastError();
    return fail(c, e == cudaErrorMemoryAllocation ? ZK_ERR_NOMEM : ZK_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
void free_powtab(PowTabOwned& t)
{
    if (t.lo) cudaFree(t.lo);
    if (t.hi) cudaFree(t.hi);
    t = PowTabOwned{};
}
cudaError_t build_powtab(PowTabOwned& t, uint32_t base, uint32_t log_E /* exponents < 2^log_E */, uint32_t log_N_tag)
{
    uint32_t lo_bits = (log_E + 1) / 2, hi_bits = log_E - lo_bits;
    std::vector<uint32_t> lo((size_t)1 << lo_bits), hi((size_t)1 << hi_bits);
    uint32_t c = 1;
    for (size_t i = 0; i < lo.size(); i++) { lo[i] = h_to_mont(c); c = h_mul(c, base); }
    uint32_t step = c;  // base^(2^lo_bits)
    c = 1;
    for (size_t i = 0; i < hi.size(); i++) { hi[i] = h_to_mont(c); c = h_mul(c, step); }
    cudaError_t e;
    if ((e = cudaMalloc(&t.lo, lo.size() * 4)) != cudaSuccess) { t.lo = nullptr; return e; }
    if ((e = cudaMalloc(&t.hi, hi.size() * 4)) != cudaSuccess) { t.hi = nullptr; free_powtab(t); return e; }
    if ((e = cudaMemcpy(t.lo, lo.data(), lo.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaMemcpy(t.hi, hi.data(), hi.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) {
        free_powtab(t);
        return e;
    }
    t.tab.lo = t.lo;
    t.tab.hi = t.hi;
    t.tab.lo_bits = lo_bits;
    t.tab.log_N = log_N_tag;
    return cudaSuccess;
}

unsigned long long* phantom_acc(zk_ctx* ctx)
{
    return (unsigned long long*)((char*)ctx->counters.p + 4 * MAX_BATCH);
}

int get_tables(zk_ctx* ctx, uint32_t log_N, SizeTables** out)
{
    auto it = ctx->tabs.find(log_N);
    if (it == ctx->tabs.end()) {
        SizeTables st;
        CK(build_powtab(st.h, h_root(log_N), log_N, log_N));
        cudaError_t ew = build_powtab(st.w, GEN, log_N, log_N);
        if (ew != cudaSuccess) { free_powtab(st.h); return fail_cuda(ctx, ew, "build_powtab"); }
        it = ctx->tabs.emplace(log_N, st).first;
    }
    *out = &it->second;
    return ZK_OK;
}

int init_tw(zk_ctx* ctx)
{
    // radix tables: for R = 2^m (m = 1..10) the R/2 powers of w_R at offset 2^(m-1)-1
    const size_t words = ntt_table_words();   // 1024 plain + the narrow-tile kernels' chunked copies (built on the device)
    std::vector<uint32_t> f(1023), inv(1023);
    for (uint32_t m = 1; m <= 10; m++) {
        uint32_t w = h_root(m), wi = h_inv(w), c = 1, ci = 1;
        size_t off = ((size_t)1 << (m - 1)) - 1;
        for (uint32_t j = 0; j < (1u << (m - 1)); j++) {
            f[off + j] = h_to_mont(c);
            inv[off + j] = h_to_mont(ci);
            c = h_mul(c, w);
            ci = h_mul(ci, wi);
        }
    }
    CK(cudaMalloc(&ctx->tw_fwd, words * 4));
    CK(cudaMalloc(&ctx->tw_inv, words * 4));
    CK(cudaMemcpy(ctx->tw_fwd, f.data(), 1023 * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->tw_inv, inv.data(), 1023 * 4, cudaMemcpyHostToDevice));
    CK(ntt2_build_tables(ctx->stream, ctx->tw_fwd, ctx->tw_inv));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

bool is_pow2(uint64_t v) { return v && !(v & (v - 1)); }
uint32_t ilog2_u64(uint64_t v)
{
    uint32_t l = 0;
    while (((uint64_t)1 << (l + 1)) <= v) l++;
    return l;
}

FoldArgs make_fold_args(uint32_t layer_k)
{
    FoldArgs f{};
    f.layer_k = layer_k;
    f.winv2k_m = h_to_mont(h_inv(h_pow(GEN, (uint64_t)1 << layer_k)));
    f.inv2_m = h_to_mont(h_inv(2));
    return f;
}

CpArgs make_cp_args(const zk_ctx* ctx, uint32_t log_n)
{
    CpArgs c{};
    c.air = ctx->air_dev;
    const uint32_t log_N = log_n + LOG_BLOWUP;
    const uint64_t n = (uint64_t)1 << log_n;
    const uint32_t g = h_root(log_n), h = h_root(log_N);
    c.log_n = log_n;
    c.log_len = log_N;
    c.g1 = h_pow(g, n - 1);
    c.g2 = h_pow(g, n - 2);
    c.g3 = h_pow(g, n - 3);
    c.halo_mask = 0xffffffffu;   // halo rows in local block order unless a caller rolls them (csrc/sharded.cu)
    const uint32_t wn = h_pow(GEN, n), hn = h_pow(h, n);
    uint32_t xn = wn;
    for (int i = 0; i < 8; i++) {
        uint32_t zi = h_inv((uint32_t)(((uint64_t)xn + P - 1) % P));
        for (int e = 0; e < 6; e++) zi = h_to_mont(zi);   // times R^6
        c.zinv_r6[i] = zi;
        xn = h_mul(xn, hn);
    }
    return c;
}

// Points c.tab at the per-size constant tables of the composition polynomial, building them on first use (once per
// context and size: input-independent domain constants, 12 N bytes).  Leaves c.tab null when they are switched off or
// do not fit in memory; the kernels then compute the same factors per point.
cudaError_t attach_cp_tab(zk_ctx* ctx, SizeTables* T, CpArgs& c, int* nl)
{
    c.tab = nullptr;
    if (!ctx->use_cp_tab) return cudaSuccess;
    const size_t N = (size_t)1 << T->h.tab.log_N;
    if (!T->cp_tab.p) {
        if (T->cp_tab.ensure(12 * N) != cudaSuccess) {
            cudaGetLastError();
            return cudaSuccess;
        }
        cudaError_t e = launch_cp_tables(ctx->stream, (uint32_t*)T->cp_tab.p, c, T->h.tab);
        if (e != cudaSuccess) return e;
        if (nl) (*nl)++;
    }
    c.tab = (const uint32_t*)T->cp_tab.p;
    return cudaSuccess;
}

struct StageTimer {
    zk_ctx* ctx;
    int idx = 0;
    explicit StageTimer(zk_ctx* c) : ctx(c) {}
    void mark()
    {
        if (ctx->stage_timing && idx < 8) cudaEventRecord(ctx->ev[idx], ctx->stream);
        idx++;
    }
};

size_t per_proof_bytes(uint32_t log_n)
{
    const size_t n = (size_t)1 << log_n, N = n << LOG_BLOWUP;
    return 4 * (2 * n + 4 * N) + 64 * N + 128 * N + sizeof(ProofState) + ((transcript_len(log_n) + 63) & ~(size_t)63);
}

int ensure_workspace(zk_ctx* ctx, uint32_t log_n, uint32_t batch)
{
    const size_t n = (size_t)1 << log_n, N = n << LOG_BLOWUP;
    const size_t trs = (transcript_len(log_n) + 63) & ~(size_t)63;
    ctx->last_nb = 0;   // buffers may move: the previous proof is no longer addressable by zk_open_queries
    CK(ctx->traces.ensure(4 * n * batch));
    CK(ctx->coeffs.ensure(4 * n * batch));
    CK(ctx->work.ensure(4 * N * batch));
    CK(ctx->f_eval.ensure(4 * N * batch));
    CK(ctx->layers.ensure(4 * 2 * N * batch));
    CK(ctx->f_heap.ensure(64 * N * batch));
    CK(ctx->layer_heaps.ensure(128 * N * batch));
    CK(ctx->states.ensure(sizeof(ProofState) * (size_t)batch));
    CK(ctx->transcripts.ensure(trs * batch));
    size_t pin = (trs + sizeof(ProofState)) * batch;
    if (pin > ctx->pinned_bytes) {
        if (ctx->pinned) cudaFreeHost(ctx->pinned);
        ctx->pinned = nullptr;
        ctx->pinned_bytes = 0;
        CK(cudaMallocHost(&ctx->pinned, pin));
        ctx->pinned_bytes = pin;
    }
    ctx->w_log_n = log_n;
    ctx->w_batch = batch;
    return ZK_OK;
}

// channel.rs:19-26 over the decommitment records (prover.rs:274-288), on the host.
void finish_channel_host(uint32_t log_n, const ProofState& ps, const uint8_t* tr, uint8_t state_out[32])
{
    const uint32_t L = log_n, log_N = log_n + LOG_BLOWUP;
    uint8_t buf[32 + 24 + 64 * 32];
    for (int i = 0; i < 8; i++) {
        uint32_t w = ps.chan_at_query[i];
        buf[4 * i] = (uint8_t)(w >> 24); buf[4 * i + 1] = (uint8_t)(w >> 16);
        buf[4 * i + 2] = (uint8_t)(w >> 8); buf[4 * i + 3] = (uint8_t)w;
    }
    size_t off = tr_off_decommit(L);
    for (uint32_t r = 0; r < 4 + L; r++) {
        size_t len = r < 4 ? 12 + 32 * (size_t)log_N : 24 + 64 * (size_t)(log_N - (r - 4));
        memcpy(buf + 32, tr + off, len);
        sha::host_sha256(buf, 32 + len, buf);
        off += len;
    }
    memcpy(state_out, buf, 32);
}

// Enqueues the whole pipeline for proofs [p0, p0 + nb) of the work space on stream s (no synchronisation).
int enqueue_prove(zk_ctx* ctx, cudaStream_t s, uint32_t log_n, uint32_t p0, uint32_t nb, SizeTables* T, const uint32_t* cp_tab,
                  bool device_tail, StageTimer* tm, int* nl_out)
{
    const uint32_t log_N = log_n + LOG_BLOWUP, L = log_n;
    const size_t n = (size_t)1 << log_n, N = n << LOG_BLOWUP;
    const size_t trs = (transcript_len(log_n) + 63) & ~(size_t)63;
    int nl = 0;
    uint32_t* traces = (uint32_t*)ctx->traces.p + (size_t)p0 * n;
    uint32_t* coeffs = (uint32_t*)ctx->coeffs.p + (size_t)p0 * n;
    uint32_t* work = (uint32_t*)ctx->work.p + (size_t)p0 * N;
    uint32_t* f_eval = (uint32_t*)ctx->f_eval.p + (size_t)p0 * N;
    uint32_t* layers = (uint32_t*)ctx->layers.p + (size_t)p0 * 2 * N;
    uint8_t* f_heap = (uint8_t*)ctx->f_heap.p + (size_t)p0 * 64 * N;
    uint8_t* layer_heaps = (uint8_t*)ctx->layer_heaps.p + (size_t)p0 * 128 * N;
    ProofState* states = (ProofState*)ctx->states.p + p0;
    uint8_t* trd = (uint8_t*)ctx->transcripts.p + (size_t)p0 * trs;

    if (tm) tm->mark();
    CK(launch_state_init(s, states, traces, n, log_n, nb, ctx->chan_seed));
    nl++;
    // Part 1 (prover.rs:60-85)
    CK(launch_interpolate(s, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_n, true, traces, n, coeffs, n, nb, &nl,
                          phantom_acc(ctx) + p0));
    if (tm) tm->mark();
    CK(launch_lde(s, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_n, coeffs, n, work, N, f_eval, N, nb, &nl));
    if (tm) tm->mark();
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p + p0;
    a.state = states;
    a.transcript = trd;
    a.transcript_stride = trs;
    a.n_layers = L;
    a.pow = T->h.tab;
    a.vals = f_eval;
    a.vals_stride = N;
    a.heap = f_heap;
    a.heap_stride = 64 * N;
    a.log_size = log_N;
    a.post = POST_F;
    CK(launch_merkle_tree(s, SRC_VALUES, a, nb, &nl));
    if (tm) tm->mark();
    // Part 2 (prover.rs:101-180)
    a.vals = layers;
    a.vals_stride = 2 * N;
    a.heap = layer_heaps;
    a.heap_stride = 128 * N;
    a.post = POST_CP;
    a.cp = make_cp_args(ctx, log_n);
    a.cp.f = f_eval;
    a.cp.f_stride = N;
    a.cp.alpha_from_state = 1;
    a.cp.tab = cp_tab;
    CK(launch_merkle_tree(s, SRC_CP, a, nb, &nl));
    if (tm) tm->mark();
    // Part 3 (prover.rs:192-254)
    for (uint32_t i = 1; i <= L; i++) {
        a.fold = make_fold_args(i - 1);
        a.fold.cur = layers + layer_vals_off(log_N, i - 1);
        a.fold.cur_stride = 2 * N;
        a.fold.log_half = log_N - i;
        a.fold.beta_from_state = 1;
        a.vals = layers + layer_vals_off(log_N, i);
        a.heap = layer_heaps + layer_heap_off(log_N, i);
        a.log_size = log_N - i;
        a.post = POST_FRI;
        a.fri_layer = i;
        if (a.log_size <= TILE_MIN_LOG && ctx->fri_tail) {   // this and every following layer: one launch
            CK(launch_fri_tail(s, a, nb, L));
            nl++;
            break;
        }
        CK(launch_merkle_tree(s, SRC_FOLD, a, nb, &nl));
    }
    if (tm) tm->mark();
    // Part 4 (prover.rs:263-289)
    CK(launch_decommit(s, log_n, f_eval, N, f_heap, 64 * N, layers, 2 * N, layer_heaps, 128 * N, states, trd, trs, nb));
    nl++;
    if (device_tail) {
        CK(launch_chan_tail(s, states, trd, trs, log_n, nb));
        nl++;
    }
    if (tm) tm->mark();
    *nl_out += nl;
    return ZK_OK;
}

// One device-resident sub-batch: traces already in ctx->traces (rows of n words).
// Batches of >= 4 proofs run as TWO lanes on two streams (the ctx stream and a side stream, event fork / join): while
// one lane is in the latency-bound upper levels and Fiat-Shamir steps of a tree, the other lane's bulk hashing fills the
// machine (measured with two contexts: +2.4 % proofs/s at 16 proofs per step).
// host_rows (optional): nb rows of n-1 trace values in (pinned) host memory, uploaded here -- the second lane's rows on
// its own stream, so that their copy overlaps the first lane's kernels.
static int prove_subbatch_body(zk_ctx* ctx, uint32_t log_n, uint32_t nb, uint8_t* transcripts_out, uint8_t* states_out,
                               const uint32_t* host_rows)
{
    const uint32_t log_N = log_n + LOG_BLOWUP;
    const size_t n_row = (size_t)1 << log_n;
    // rows of n-1 host values -> rows of n device words (one plain copy per row: a pitched 2-D copy of rows this wide
    // falls off the DMA fast path)
    auto upload = [&](cudaStream_t st, uint32_t b0, uint32_t b1) -> cudaError_t {
        for (uint32_t b = b0; host_rows && b < b1; b++) {
            cudaError_t e = cudaMemcpyAsync((uint32_t*)ctx->traces.p + n_row * b, host_rows + (n_row - 1) * (size_t)b, 4 * (n_row - 1),
                                            cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
    const size_t trs = (transcript_len(log_n) + 63) & ~(size_t)63, trl = transcript_len(log_n);
    SizeTables* T;
    int rc = get_tables(ctx, log_N, &T);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    int nl = 0;
    StageTimer tm(ctx);
    ProofState* states = (ProofState*)ctx->states.p;
    uint8_t* trd = (uint8_t*)ctx->transcripts.p;
    const bool device_tail = nb >= 8;  // the sequential channel tail: one CTA per proof on the device, or the host
    // per-size constants are built on the ctx stream before the lanes fork
    CpArgs cp0 = make_cp_args(ctx, log_n);
    CK(attach_cp_tab(ctx, T, cp0, &nl));
    const bool two_lanes = ctx->lanes && nb >= 4 && !ctx->stage_timing && ctx->lane_stream;
    if (!two_lanes) {
        CK(upload(s, 0, nb));
        if ((rc = enqueue_prove(ctx, s, log_n, 0, nb, T, cp0.tab, device_tail, &tm, &nl))) return rc;
    } else {
        const uint32_t nb0 = (nb + 1) / 2;
        CK(cudaEventRecord(ctx->ev_fork, s));
        CK(cudaStreamWaitEvent(ctx->lane_stream, ctx->ev_fork, 0));
        CK(upload(s, 0, nb0));
        if ((rc = enqueue_prove(ctx, s, log_n, 0, nb0, T, cp0.tab, device_tail, nullptr, &nl))) return rc;
        CK(upload(ctx->lane_stream, nb0, nb));
        if ((rc = enqueue_prove(ctx, ctx->lane_stream, log_n, nb0, nb - nb0, T, cp0.tab, device_tail, nullptr, &nl))) return rc;
        CK(cudaEventRecord(ctx->ev_join, ctx->lane_stream));
        CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    }
    uint8_t* pin_tr = (uint8_t*)ctx->pinned;
    ProofState* pin_st = (ProofState*)(pin_tr + trs * nb);
    CK(cudaMemcpyAsync(pin_tr, trd, trs * nb, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(pin_st, states, sizeof(ProofState) * (size_t)nb, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    ctx->launches += nl;
    ctx->last_log_n = log_n;
    ctx->last_nb = nb;
    if (ctx->stage_timing) {
        for (int i = 0; i < 6; i++) cudaEventElapsedTime(&ctx->stage_ms[i], ctx->ev[i], ctx->ev[i + 1]);
        cudaEventElapsedTime(&ctx->stage_ms[6], ctx->ev[0], ctx->ev[6]);
    }
    for (uint32_t b = 0; b < nb; b++) {
        if (pin_st[b].pad[0] != 0)
            return fail(ctx, ZK_ERR_ASSERT, "last FRI layer is not constant (prover.rs:238)");
        memcpy(transcripts_out + trl * b, pin_tr + trs * b, trl);
        if (device_tail) {
            uint8_t* so = states_out + 32 * (size_t)b;
            for (int i = 0; i < 8; i++) {
                uint32_t wv = pin_st[b].chan[i];
                so[4 * i] = (uint8_t)(wv >> 24); so[4 * i + 1] = (uint8_t)(wv >> 16);
                so[4 * i + 2] = (uint8_t)(wv >> 8); so[4 * i + 3] = (uint8_t)wv;
            }
        } else {
            finish_channel_host(log_n, pin_st[b], pin_tr + trs * b, states_out + 32 * (size_t)b);
        }
    }
    return ZK_OK;
}

int prove_subbatch(zk_ctx* ctx, uint32_t log_n, uint32_t nb, uint8_t* transcripts_out, uint8_t* states_out,
                   const uint32_t* host_rows = nullptr)
{
    ctx->last_nb = 0;   // nothing is resident for zk_open_queries until this sub-batch has completed
    const int rc = prove_subbatch_body(ctx, log_n, nb, transcripts_out, states_out, host_rows);
    if (rc != ZK_OK) {
        // A failed launch may leave the second lane forked and the per-proof tile-arrival counters / phantom accumulators
        // non-zero (only a kernel that completes re-arms them): drain both streams and zero them, so that the next
        // prove on this context starts clean.  Best effort: the original error is what the caller sees.
        const std::string keep = ctx->err;
        cudaStreamSynchronize(ctx->stream);
        if (ctx->lane_stream) cudaStreamSynchronize(ctx->lane_stream);
        cudaMemset(ctx->counters.p, 0, 12 * MAX_BATCH);
        cudaGetLastError();
        ctx->err = keep;
        ctx->last_nb = 0;
    }
    return rc;
}

uint32_t pick_subbatch(zk_ctx* ctx, uint32_t log_n, uint32_t n_proofs)
{
    // fast path: the work space already fits this request (cudaMemGetInfo costs ~1 ms)
    if (!ctx->mem_budget && ctx->w_log_n == log_n && ctx->w_batch >= n_proofs) return n_proofs;
    size_t budget = ctx->mem_budget;
    if (!budget) {
        size_t fr = 0, tot = 0;
        cudaMemGetInfo(&fr, &tot);
        size_t held = ctx->traces.bytes + ctx->coeffs.bytes + ctx->work.bytes + ctx->f_eval.bytes + ctx->layers.bytes +
                      ctx->f_heap.bytes + ctx->layer_heaps.bytes + ctx->states.bytes + ctx->transcripts.bytes;
        budget = (size_t)((fr + held) * 0.7);
    }
    // the per-size constant tables of the composition polynomial (12 N bytes, once per size) come out of the same budget
    // when they are not built yet
    {
        auto it = ctx->tabs.find(log_n + LOG_BLOWUP);
        const bool have = it != ctx->tabs.end() && it->second.cp_tab.p;
        const size_t tab = ctx->use_cp_tab && !have ? (size_t)12 << (log_n + LOG_BLOWUP) : 0;
        budget = budget > tab ? budget - tab : 0;
    }
    size_t per = per_proof_bytes(log_n);
    size_t nb = budget / per;
    if (nb < 1) nb = 1;
    if (nb > MAX_BATCH) nb = MAX_BATCH;
    if (nb > n_proofs) nb = n_proofs;
    return (uint32_t)nb;
}

int check_log_n(zk_ctx* ctx, uint32_t log_n)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (log_n < ZK_MIN_LOG_N || log_n > ZK_MAX_LOG_N) return fail(ctx, ZK_ERR_INVALID, "log_n out of range [3, 27]");
    return ZK_OK;
}

}  // namespace zkh

// =============================================================================== C ABI
extern "C" {

const char* zk_version(void) { return "zkstark_b200 0.1 (sm_100a)"; }

int zk_ctx_create_on_stream(int device, void* cuda_stream, zk_ctx** out)
{
    if (!out) return ZK_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        cudaGetLastError();
        return ZK_ERR_CUDA;  // no CPU fallback
    }
    if (cudaSetDevice(device) != cudaSuccess) return ZK_ERR_CUDA;
    zk_ctx* ctx = new zk_ctx();
    ctx->device = device;
    if (cuda_stream) {
        ctx->stream = (cudaStream_t)cuda_stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return ZK_ERR_CUDA; }
        ctx->own_stream = true;
    }
    for (int i = 0; i < 8; i++)
        if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) { delete ctx; return ZK_ERR_CUDA; }
    { const char* e = getenv("ZK_NO_LANES"); if (e && e[0] == '1') ctx->lanes = false; }
    { const char* e = getenv("ZK_NO_FRI_TAIL"); if (e && e[0] == '1') ctx->fri_tail = false; }
    { const char* e = getenv("ZK_TRACE_DEVICE"); if (e && e[0] == '1') ctx->trace_on_device = true; }
    { const char* e = getenv("ZK_HOST_THREADS"); if (e) ctx->host_threads = (uint32_t)atoi(e); }
    if (cudaStreamCreateWithFlags(&ctx->lane_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) { delete ctx; return ZK_ERR_CUDA; }
    int rc = init_tw(ctx);
    if (rc == ZK_OK && sha_init_constants() != cudaSuccess) rc = ZK_ERR_CUDA;
    // [0, 4 MAX_BATCH): tile-arrival counters; [4 MAX_BATCH, 12 MAX_BATCH): 64-bit phantom-point accumulators
    if (rc == ZK_OK && (ctx->counters.ensure(12 * MAX_BATCH) != cudaSuccess ||
                        cudaMemset(ctx->counters.p, 0, 12 * MAX_BATCH) != cudaSuccess)) rc = ZK_ERR_CUDA;
    if (rc) { delete ctx; return rc; }
    *out = ctx;
    return ZK_OK;
}

int zk_ctx_create(int device, zk_ctx** out) { return zk_ctx_create_on_stream(device, nullptr, out); }

void zk_ctx_destroy(zk_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->tabs) { free_powtab(kv.second.h); free_powtab(kv.second.w); kv.second.cp_tab.release(); }
    for (auto& kv : ctx->shift_tabs) free_powtab(kv.second);
    DevBuf* bufs[] = {&ctx->counters, &ctx->flushbuf, &ctx->s0, &ctx->s1, &ctx->s2, &ctx->traces, &ctx->coeffs, &ctx->work, &ctx->f_eval,
                      &ctx->layers, &ctx->f_heap, &ctx->layer_heaps, &ctx->states, &ctx->transcripts};
    for (DevBuf* b : bufs) b->release();
    if (ctx->tw_fwd) cudaFree(ctx->tw_fwd);
    if (ctx->tw_inv) cudaFree(ctx->tw_inv);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    for (int i = 0; i < 8; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->lane_stream) { cudaStreamSynchronize(ctx->lane_stream); cudaStreamDestroy(ctx->lane_stream); }
    if (ctx->trace_stream) { cudaStreamSynchronize(ctx->trace_stream); cudaStreamDestroy(ctx->trace_stream); }
    for (auto& sl : ctx->slot) {
        if (sl.worker.joinable()) sl.worker.join();
        sl.traces.release();
        sl.secrets.release();
        if (sl.pinned) cudaFreeHost(sl.pinned);
        if (sl.ready) cudaEventDestroy(sl.ready);
    }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* zk_last_error(const zk_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }
uint64_t zk_ctx_launch_count(const zk_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* zk_ctx_stream(const zk_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int zk_ctx_synchronize(zk_ctx* ctx)
{
    if (!ctx) return ZK_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
int zk_ctx_set_stage_timing(zk_ctx* ctx, int on)
{
    if (!ctx) return ZK_ERR_INVALID;
    ctx->stage_timing = on != 0;
    return ZK_OK;
}
int zk_ctx_last_stage_ms(zk_ctx* ctx, float out_ms[7])
{
    if (!ctx || !out_ms) return ZK_ERR_INVALID;
    for (int i = 0; i < 7; i++) out_ms[i] = ctx->stage_ms[i];
    return ZK_OK;
}
int zk_ctx_trim(zk_ctx* ctx, int tables_too)
{
    if (!ctx) return ZK_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->lane_stream) CK(cudaStreamSynchronize(ctx->lane_stream));
    if (ctx->trace_stream) CK(cudaStreamSynchronize(ctx->trace_stream));
    if (ctx->slot_count) return fail(ctx, ZK_ERR_INVALID, "batches are still in flight (zk_prove_stream_collect them first)");
    DevBuf* bufs[] = {&ctx->s0, &ctx->s1, &ctx->s2, &ctx->flushbuf, &ctx->traces, &ctx->coeffs, &ctx->work, &ctx->f_eval, &ctx->layers,
                      &ctx->f_heap, &ctx->layer_heaps, &ctx->states, &ctx->transcripts, &ctx->slot[0].traces, &ctx->slot[1].traces};
    for (DevBuf* b : bufs) b->release();
    for (auto& sl : ctx->slot) {
        if (sl.pinned) cudaFreeHost(sl.pinned);
        sl.pinned = nullptr;
        sl.pinned_bytes = 0;
    }
    ctx->w_log_n = ctx->w_batch = 0;
    ctx->last_nb = 0;
    if (tables_too) {
        for (auto& kv : ctx->tabs) { free_powtab(kv.second.h); free_powtab(kv.second.w); kv.second.cp_tab.release(); }
        ctx->tabs.clear();
        for (auto& kv : ctx->shift_tabs) free_powtab(kv.second);
        ctx->shift_tabs.clear();
    } else {
        for (auto& kv : ctx->tabs) kv.second.cp_tab.release();   // 12 N bytes each: rebuilt on the next prove of that size
    }
    return ZK_OK;
}

int zk_ctx_set_memory_budget(zk_ctx* ctx, size_t bytes)
{
    if (!ctx) return ZK_ERR_INVALID;
    ctx->mem_budget = bytes;
    return ZK_OK;
}

// ------------------------------------------------------------------ device-pointer stages
int zk_interpolate_subgroup_dev(zk_ctx* ctx, uint32_t* y, uint32_t n_points, uint32_t log_n, uint32_t* coeffs_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    const uint32_t n = 1u << log_n;
    if (!y || !coeffs_out || (n_points != n && n_points != n - 1)) return fail(ctx, ZK_ERR_INVALID, "n_points must be n or n-1");
    if (y == coeffs_out && log_n > 10) return fail(ctx, ZK_ERR_INVALID, "in-place interpolation needs log_n <= 10");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    if ((rc = get_tables(ctx, log_n + LOG_BLOWUP, &T))) return rc;
    int nl = 0;
    CK(launch_interpolate(ctx->stream, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_n, n_points == n - 1, y, n,
                          coeffs_out, n, 1, &nl, phantom_acc(ctx)));
    ctx->launches += nl;
    return ZK_OK;
}

int zk_lde_coset_dev(zk_ctx* ctx, const uint32_t* coeffs, uint32_t log_n, uint32_t log_blowup, uint32_t shift,
                     uint32_t* evals_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!coeffs || !evals_out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (log_blowup != LOG_BLOWUP) return fail(ctx, ZK_ERR_INVALID, "log_blowup must be 3 (prover.rs:56-57)");
    CK(cudaSetDevice(ctx->device));
    const size_t N = (size_t)1 << (log_n + LOG_BLOWUP);
    SizeTables* T;
    if ((rc = get_tables(ctx, log_n + LOG_BLOWUP, &T))) return rc;
    CK(ctx->s2.ensure(4 * N));
    PowTab wtab = T->w.tab;
    shift %= P;
    if (shift == 0) return fail(ctx, ZK_ERR_INVALID, "shift must be non-zero mod p");
    if (shift != GEN) {
        // powers of another coset shift: built once per (shift, size) and kept with the context (no stream synchronisation)
        const uint64_t key = ((uint64_t)shift << 8) | (log_n + LOG_BLOWUP);
        auto it = ctx->shift_tabs.find(key);
        if (it == ctx->shift_tabs.end()) {
            PowTabOwned tmp;
            CK(build_powtab(tmp, shift, log_n, log_n + LOG_BLOWUP));
            it = ctx->shift_tabs.emplace(key, tmp).first;
        }
        wtab = it->second.tab;
    }
    int nl = 0;
    cudaError_t e = launch_lde(ctx->stream, ctx->tw_fwd, ctx->tw_inv, T->h.tab, wtab, log_n, coeffs, 0,
                               (uint32_t*)ctx->s2.p, N, evals_out, N, 1, &nl);
    if (e != cudaSuccess) return fail_cuda(ctx, e, "launch_lde");
    ctx->launches += nl;
    return ZK_OK;
}

int zk_fri_fold_dev(zk_ctx* ctx, const uint32_t* cur, uint64_t len, uint32_t layer_k, uint32_t beta_raw, uint32_t* next_out)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!cur || !next_out || !is_pow2(len) || len < 2) return fail(ctx, ZK_ERR_INVALID, "len must be a power of two >= 2");
    const uint32_t log_len = ilog2_u64(len);
    if (log_len + layer_k > 30) return fail(ctx, ZK_ERR_INVALID, "len << layer_k exceeds 2^30");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    int rc = get_tables(ctx, log_len + layer_k, &T);
    if (rc) return rc;
    FoldArgs f = make_fold_args(layer_k);
    f.cur = cur;
    f.cur_stride = 0;
    f.beta_raw = beta_raw;
    CK(launch_fold(ctx->stream, cur, next_out, log_len - 1, f, T->h.tab, 1, 0));
    ctx->launches += 1;
    return ZK_OK;
}

int zk_cp_eval_fibsq_dev(zk_ctx* ctx, const uint32_t* f_eval, uint32_t log_n, uint32_t log_blowup,
                         const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last, uint32_t* cp_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!f_eval || !cp_out || !alpha_raw) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (log_blowup != LOG_BLOWUP) return fail(ctx, ZK_ERR_INVALID, "log_blowup must be 3");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    if ((rc = get_tables(ctx, log_n + LOG_BLOWUP, &T))) return rc;
    CpArgs c = make_cp_args(ctx, log_n);
    c.f = f_eval;
    c.f_stride = 0;
    for (int j = 0; j < 3; j++) c.alpha_raw[j] = alpha_raw[j];
    c.a_first = a_first % P;
    c.a_last = a_last % P;
    int nl = 1;
    CK(attach_cp_tab(ctx, T, c, &nl));
    CK(launch_cp(ctx->stream, cp_out, c, T->h.tab, 1, 0));
    ctx->launches += nl;
    return ZK_OK;
}

int zk_merkle_build_dev(zk_ctx* ctx, const uint32_t* leaves, uint64_t size, uint8_t* heap_out)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!leaves || !heap_out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (!is_pow2(size) || size > ((uint64_t)1 << 30)) return fail(ctx, ZK_ERR_INVALID, "size must be a power of two (merkle.rs:17-21)");
    CK(cudaSetDevice(ctx->device));
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p;
    a.vals = const_cast<uint32_t*>(leaves);
    a.heap = heap_out;
    a.log_size = ilog2_u64(size);
    a.post = POST_NONE;
    int nl = 0;
    CK(launch_merkle_tree(ctx->stream, SRC_VALUES, a, 1, &nl));
    ctx->launches += nl;
    return ZK_OK;
}

// ------------------------------------------------------------------ host-buffer stages
int zk_interpolate_subgroup(zk_ctx* ctx, const uint32_t* y, uint32_t n_points, uint32_t log_n, uint32_t* coeffs_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    const size_t n = (size_t)1 << log_n;
    if (!y || !coeffs_out || (n_points != n && n_points != n - 1)) return fail(ctx, ZK_ERR_INVALID, "n_points must be n or n-1");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->s0.ensure(4 * n));
    CK(ctx->s1.ensure(4 * n));
    CK(cudaMemsetAsync(ctx->s0.p, 0, 4 * n, ctx->stream));
    CK(cudaMemcpyAsync(ctx->s0.p, y, 4 * (size_t)n_points, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = zk_interpolate_subgroup_dev(ctx, (uint32_t*)ctx->s0.p, n_points, log_n, (uint32_t*)ctx->s1.p))) return rc;
    CK(cudaMemcpyAsync(coeffs_out, ctx->s1.p, 4 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_lde_coset(zk_ctx* ctx, const uint32_t* coeffs, uint32_t log_n, uint32_t log_blowup, uint32_t shift,
                 uint32_t* evals_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!coeffs || !evals_out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (log_blowup != LOG_BLOWUP) return fail(ctx, ZK_ERR_INVALID, "log_blowup must be 3 (prover.rs:56-57)");
    const size_t n = (size_t)1 << log_n, N = n << LOG_BLOWUP;
    CK(cudaSetDevice(ctx->device));
    CK(ctx->s0.ensure(4 * n));
    CK(ctx->s1.ensure(4 * N));
    CK(cudaMemcpyAsync(ctx->s0.p, coeffs, 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = zk_lde_coset_dev(ctx, (const uint32_t*)ctx->s0.p, log_n, log_blowup, shift, (uint32_t*)ctx->s1.p))) return rc;
    CK(cudaMemcpyAsync(evals_out, ctx->s1.p, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_fri_fold(zk_ctx* ctx, const uint32_t* cur, uint64_t len, uint32_t layer_k, uint32_t beta_raw, uint32_t* next_out)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!cur || !next_out || !is_pow2(len) || len < 2) return fail(ctx, ZK_ERR_INVALID, "len must be a power of two >= 2");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->s0.ensure(4 * len));
    CK(ctx->s1.ensure(2 * len));
    CK(cudaMemcpyAsync(ctx->s0.p, cur, 4 * len, cudaMemcpyHostToDevice, ctx->stream));
    int rc = zk_fri_fold_dev(ctx, (const uint32_t*)ctx->s0.p, len, layer_k, beta_raw, (uint32_t*)ctx->s1.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(next_out, ctx->s1.p, 2 * len, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

int zk_cp_eval_fibsq(zk_ctx* ctx, const uint32_t* f_eval, uint32_t log_n, uint32_t log_blowup,
                     const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last, uint32_t* cp_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!f_eval || !cp_out || !alpha_raw) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    const size_t N = (size_t)1 << (log_n + LOG_BLOWUP);
    CK(cudaSetDevice(ctx->device));
    CK(ctx->s0.ensure(4 * N));
    CK(ctx->s1.ensure(4 * N));
    CK(cudaMemcpyAsync(ctx->s0.p, f_eval, 4 * N, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = zk_cp_eval_fibsq_dev(ctx, (const uint32_t*)ctx->s0.p, log_n, log_blowup, alpha_raw, a_first, a_last,
                                   (uint32_t*)ctx->s1.p)))
        return rc;
    CK(cudaMemcpyAsync(cp_out, ctx->s1.p, 4 * N, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

static int merkle_host(zk_ctx* ctx, const uint32_t* leaves, uint64_t size, uint8_t* out, bool root_only)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!leaves || !out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (!is_pow2(size) || size > ((uint64_t)1 << 30)) return fail(ctx, ZK_ERR_INVALID, "size must be a power of two (merkle.rs:17-21)");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->s0.ensure(4 * size));
    CK(ctx->s1.ensure(64 * size));
    CK(cudaMemcpyAsync(ctx->s0.p, leaves, 4 * size, cudaMemcpyHostToDevice, ctx->stream));
    int rc = zk_merkle_build_dev(ctx, (const uint32_t*)ctx->s0.p, size, (uint8_t*)ctx->s1.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(out, ctx->s1.p, root_only ? 32 : 32 * (2 * size - 1), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}
int zk_merkle_build(zk_ctx* ctx, const uint32_t* leaves, uint64_t size, uint8_t* heap_out)
{
    return merkle_host(ctx, leaves, size, heap_out, false);
}
int zk_merkle_root(zk_ctx* ctx, const uint32_t* leaves, uint64_t size, uint8_t root_out[32])
{
    return merkle_host(ctx, leaves, size, root_out, true);
}

int zk_merkle_trace(const uint8_t* heap, uint64_t size, uint64_t index, uint8_t* path_out)
{
    if (!heap || !path_out || !is_pow2(size) || index >= size) return ZK_ERR_INVALID;
    uint64_t i = index + size - 1;  // merkle.rs:55
    int n = 0;
    while (i != 0) {
        uint64_t sib = (i % 2 == 0) ? i - 1 : i + 1;  // merkle.rs:58-66
        memcpy(path_out + 32 * (size_t)n, heap + 32 * sib, 32);
        i = (i - 1) >> 1;
        n++;
    }
    return n;
}

// ------------------------------------------------------------------ prove
size_t zk_transcript_len(uint32_t log_n) { return transcript_len(log_n); }

int zk_trace_fibsq(uint32_t log_n, uint32_t secret, uint32_t* out)
{
    if (!out || log_n < 1 || log_n > 30) return ZK_ERR_INVALID;
    const uint64_t n1 = ((uint64_t)1 << log_n) - 1;
    // Montgomery domain for the sequential recurrence a_i = a_{i-2}^2 + a_{i-1}^2 (prover.rs:32-39)
    // (one dependent squaring per step: a_{i-1}^2 of step i is a_{i-2}^2 of step i+1)
    uint32_t x = to_mont(1), y = to_mont(secret % P);
    uint32_t xx = mont_mul(x, x);
    out[0] = 1;
    if (n1 > 1) out[1] = secret % P;
    for (uint64_t i = 2; i < n1; i++) {
        const uint32_t yy = mont_mul(y, y);
        const uint32_t z = fadd(xx, yy);
        out[i] = from_mont(z);
        xx = yy;
        y = z;
    }
    return ZK_OK;
}

// prover.rs:32-39 for n_proofs secrets at once on the host: groups of 8 traces advance as SIMD lanes (AVX2), the groups are
// spread over `threads` host threads (0: all hardware threads).  out: n_proofs rows of 2^log_n - 1 values.
int zk_trace_fibsq_batch(uint32_t log_n, const uint32_t* secrets, uint32_t n_proofs, uint32_t threads, uint32_t* out)
{
    return zk_trace_air_batch(nullptr, log_n, secrets, n_proofs, threads, out);
}

// the public description of a program (include/zkstark_b200.h zk_air) -> host and device forms.  Null: the
// reference's FibonacciSq.  Coefficients are reduced like every u32 that enters the field (field.rs:20-24).
static void air_forms(const zk_air* a, HostAir& h, AirCoef& d)
{
    h = HostAir{};
    d = AirCoef{};
    if (!a) return;
    h.a0 = a->a0 % P; h.c20 = a->c20 % P; h.c02 = a->c02 % P; h.c11 = a->c11 % P; h.c10 = a->c10 % P; h.c01 = a->c01 % P;
    h.c00 = a->c00 % P;
    const bool fibsq_shape = h.c20 == 1 && h.c02 == 1 && !h.c11 && !h.c10 && !h.c01 && !h.c00;
    const bool keep_generic = (a->flags & ZK_AIR_NO_SPECIALISE) != 0;
    h.generic = keep_generic || !(fibsq_shape && h.a0 == 1);        // the trace recurrence
    d.generic = (keep_generic || !fibsq_shape) ? 1u : 0u;           // the transition constraint
    d.c20m = h_to_mont(h.c20); d.c02m = h_to_mont(h.c02); d.c11m = h_to_mont(h.c11);
    d.c10 = h.c10; d.c01 = h.c01;
    d.c00r = from_mont(h.c00);
}

int zk_ctx_set_air(zk_ctx* ctx, const zk_air* air)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (air && (air->flags & ~ZK_AIR_NO_SPECIALISE)) return fail(ctx, ZK_ERR_INVALID, "unknown zk_air flag");
    if (ctx->slot_count) return fail(ctx, ZK_ERR_INVALID, "a submitted batch is still uncollected");
    air_forms(air, ctx->air, ctx->air_dev);
    ctx->air_public = air ? *air : zk_air{1, 1, 1, 0, 0, 0, 0, 0};
    return ZK_OK;
}

int zk_ctx_get_air(const zk_ctx* ctx, zk_air* out)
{
    if (!ctx || !out) return ZK_ERR_INVALID;
    *out = ctx->air_public;
    return ZK_OK;
}

int zk_trace_air(const zk_air* air, uint32_t log_n, uint32_t secret, uint32_t* out)
{
    return zk_trace_air_batch(air, log_n, &secret, 1, 1, out);
}

int zk_trace_air_batch(const zk_air* air_in, uint32_t log_n, const uint32_t* secrets, uint32_t n_proofs, uint32_t threads,
                       uint32_t* out)
{
    if (!secrets || !out || n_proofs == 0 || log_n < 1 || log_n > 30) return ZK_ERR_INVALID;
    if (air_in && (air_in->flags & ~ZK_AIR_NO_SPECIALISE)) return ZK_ERR_INVALID;
    HostAir air;
    AirCoef unused;
    air_forms(air_in, air, unused);
    const size_t n1 = ((size_t)1 << log_n) - 1;
    uint32_t T = threads ? threads : std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    const uint32_t gsz = host_trace_group(n_proofs, T, air.generic);
    const uint32_t n_groups = (n_proofs + gsz - 1) / gsz;
    if (T > n_groups) T = n_groups;
    auto body = [&](uint32_t t) {
        for (uint32_t gi = t; gi < n_groups; gi += T) {
            const uint32_t b0 = gsz * gi, cnt = n_proofs - b0 < gsz ? n_proofs - b0 : gsz;
            uint32_t* rows[16];
            for (uint32_t j = 0; j < cnt; j++) rows[j] = out + n1 * (b0 + j);
            host_traces(air, log_n, secrets + b0, rows, cnt);
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < T; t++) th.emplace_back(body, t);
    body(0);
    for (auto& x : th) x.join();
    return ZK_OK;
}

int zk_prove_fibsq_resident(zk_ctx* ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t* traces_dev,
                            uint8_t* transcripts_out, uint8_t* states_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!traces_dev || !transcripts_out || !states_out || n_proofs == 0) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << log_n, trl = transcript_len(log_n);
    const uint32_t nbmax = pick_subbatch(ctx, log_n, n_proofs);
    if ((rc = ensure_workspace(ctx, log_n, nbmax))) return rc;
    for (uint32_t done = 0; done < n_proofs; done += nbmax) {
        uint32_t nb = n_proofs - done < nbmax ? n_proofs - done : nbmax;
        CK(cudaMemcpyAsync(ctx->traces.p, traces_dev + n * done, 4 * n * nb, cudaMemcpyDeviceToDevice, ctx->stream));
        if ((rc = prove_subbatch(ctx, log_n, nb, transcripts_out + trl * done, states_out + 32 * (size_t)done))) return rc;
    }
    return ZK_OK;
}

int zk_prove_fibsq_batch(zk_ctx* ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t* traces,
                         uint8_t* transcripts_out, uint8_t* states_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!traces || !transcripts_out || !states_out || n_proofs == 0) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    CK(cudaSetDevice(ctx->device));
    const size_t n = (size_t)1 << log_n, trl = transcript_len(log_n);
    const uint32_t nbmax = pick_subbatch(ctx, log_n, n_proofs);
    if ((rc = ensure_workspace(ctx, log_n, nbmax))) return rc;
    for (uint32_t done = 0; done < n_proofs; done += nbmax) {
        uint32_t nb = n_proofs - done < nbmax ? n_proofs - done : nbmax;
        if ((rc = prove_subbatch(ctx, log_n, nb, transcripts_out + trl * done, states_out + 32 * (size_t)done,
                                 traces + (n - 1) * (size_t)done))) return rc;
    }
    return ZK_OK;
}

int zk_prove_stream_submit(zk_ctx* ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t* secrets);
int zk_prove_stream_collect(zk_ctx* ctx, uint8_t* transcripts_out, uint8_t* states_out);
int zk_prove_fibsq_batch_secrets(zk_ctx* ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t* secrets,
                                 uint8_t* transcripts_out, uint8_t* states_out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!secrets || !transcripts_out || !states_out || n_proofs == 0) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (ctx->slot_count) return fail(ctx, ZK_ERR_INVALID, "batches of zk_prove_stream_submit are in flight: collect them first");
    CK(cudaSetDevice(ctx->device));
    const size_t trl = transcript_len(log_n);
    // sub-batches go through the submit / collect pipeline: the traces of sub-batch j + 1 are produced while j is proved.
    // A call that would fit one sub-batch is still cut in four (of at least 16 proofs) so that only the first quarter's
    // trace generation is exposed.
    uint32_t nbmax = pick_subbatch(ctx, log_n, n_proofs);
    if (n_proofs >= 64 && nbmax >= n_proofs) nbmax = (n_proofs + 3) / 4;
    auto count = [&](uint32_t done) { return n_proofs - done < nbmax ? n_proofs - done : nbmax; };
    if ((rc = zk_prove_stream_submit(ctx, log_n, count(0), secrets))) return rc;
    for (uint32_t done = 0; done < n_proofs; done += nbmax) {
        const uint32_t next = done + nbmax;
        if (next < n_proofs && (rc = zk_prove_stream_submit(ctx, log_n, count(next), secrets + next))) break;
        if ((rc = zk_prove_stream_collect(ctx, transcripts_out + trl * done, states_out + 32 * (size_t)done))) break;
    }
    if (rc) {   // drain what is still in flight (results are discarded; the caller sees the first error)
        const std::string keep = ctx->err;
        while (ctx->slot_count) {
            zk_ctx::StreamSlot& sl = ctx->slot[ctx->slot_head];
            if (sl.worker.joinable()) sl.worker.join();
            ctx->slot_head ^= 1;
            ctx->slot_count--;
        }
        if (ctx->trace_stream) cudaStreamSynchronize(ctx->trace_stream);
        ctx->err = keep;
    }
    return rc;
}

// ---- pipelined secrets -> proofs.  The FibonacciSq recurrence (prover.rs:32-39) is one dependent squaring per step: a 2^20
// trace costs ~7.5 ms on a host core and ~95 ms as a lone GPU thread, whatever the batch size -- against ~4 ms of proving
// per proof.  It is therefore produced on the HOST, one trace per thread over all cores, one batch AHEAD of the proving:
// submit() starts the worker threads of a batch (each writes its rows into pinned memory and enqueues their upload on a
// side stream), collect() proves the oldest submitted batch.  With one batch submitted ahead, trace generation and upload
// disappear behind the previous batch's hashing.  (ZK_TRACE_DEVICE=1: the traces are generated by trace_batch_kernel on
// the side stream instead -- one GPU thread per proof; pays only for very many short traces.)
static void stream_worker(zk_ctx* ctx, zk_ctx::StreamSlot* sl)
{
    cudaSetDevice(ctx->device);
    const size_t n = (size_t)1 << sl->log_n;
    uint32_t T = ctx->host_threads ? ctx->host_threads : std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if (T > sl->n) T = sl->n;
    std::vector<cudaError_t> errs(T, cudaSuccess);
    // a task is a group of 4 or 8 proofs (host_trace_group: what the CPU's vector width and the thread count make
    // fastest): they advance together as SIMD lanes (host_traces), then their rows are uploaded
    const uint32_t gsz = host_trace_group(sl->n, T, ctx->air.generic);
    const uint32_t n_groups = (sl->n + gsz - 1) / gsz;
    if (T > n_groups) T = n_groups;
    errs.assign(T, cudaSuccess);
    auto body = [&](uint32_t t) {
        cudaSetDevice(ctx->device);
        for (uint32_t gi = t; gi < n_groups; gi += T) {
            const uint32_t b0 = gsz * gi, cnt = sl->n - b0 < gsz ? sl->n - b0 : gsz;
            uint32_t* rows[16];
            for (uint32_t j = 0; j < cnt; j++) rows[j] = (uint32_t*)sl->pinned + n * (b0 + j);
            host_traces(ctx->air, sl->log_n, sl->host_secrets.data() + b0, rows, cnt);
            for (uint32_t j = 0; j < cnt; j++) {
                cudaError_t e = cudaMemcpyAsync((uint32_t*)sl->traces.p + n * (b0 + j), rows[j], 4 * (n - 1), cudaMemcpyHostToDevice,
                                                ctx->trace_stream);
                if (e != cudaSuccess) errs[t] = e;
            }
        }
    };
    std::vector<std::thread> th;
    for (uint32_t t = 1; t < T; t++) th.emplace_back(body, t);
    body(0);
    for (auto& x : th) x.join();
    sl->err = cudaSuccess;
    for (cudaError_t e : errs) if (e != cudaSuccess) sl->err = e;
    cudaError_t e = cudaEventRecord(sl->ready, ctx->trace_stream);
    if (e != cudaSuccess) sl->err = e;
}

int zk_prove_stream_submit(zk_ctx* ctx, uint32_t log_n, uint32_t n_proofs, const uint32_t* secrets)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!secrets || n_proofs == 0) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (ctx->slot_count >= 2) return fail(ctx, ZK_ERR_INVALID, "two batches are already in flight: collect one first");
    CK(cudaSetDevice(ctx->device));
    if (!ctx->trace_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CK(cudaStreamCreateWithPriority(&ctx->trace_stream, cudaStreamNonBlocking, hi));
    }
    zk_ctx::StreamSlot& sl = ctx->slot[(ctx->slot_head + ctx->slot_count) & 1];
    const size_t n = (size_t)1 << log_n;
    if (!sl.ready) CK(cudaEventCreateWithFlags(&sl.ready, cudaEventDisableTiming));
    CK(sl.traces.ensure(4 * n * n_proofs));
    sl.host_secrets.assign(secrets, secrets + n_proofs);   // the caller's array may go away before the workers read it
    sl.log_n = log_n;
    sl.n = n_proofs;
    sl.err = cudaSuccess;
    if (ctx->trace_on_device) {
        const HostAir& h = ctx->air;
        const TraceAir ta{h.generic ? 1u : 0u, h.a0, h_to_mont(h.c20), h_to_mont(h.c02), h_to_mont(h.c11), h_to_mont(h.c10),
                          h_to_mont(h.c01), h_to_mont(h.c00)};
        CK(sl.secrets.ensure(4 * (size_t)n_proofs));
        CK(cudaMemcpyAsync(sl.secrets.p, sl.host_secrets.data(), 4 * (size_t)n_proofs, cudaMemcpyHostToDevice, ctx->trace_stream));
        CK(launch_trace_batch(ctx->trace_stream, (const uint32_t*)sl.secrets.p, (uint32_t*)sl.traces.p, n, log_n, n_proofs, ta));
        CK(cudaEventRecord(sl.ready, ctx->trace_stream));
        ctx->launches += 1;
    } else {
        if (sl.pinned_bytes < 4 * n * n_proofs) {
            if (sl.pinned) cudaFreeHost(sl.pinned);
            sl.pinned = nullptr;
            sl.pinned_bytes = 0;
            CK(cudaMallocHost(&sl.pinned, 4 * n * n_proofs));
            sl.pinned_bytes = 4 * n * n_proofs;
        }
        sl.worker = std::thread(stream_worker, ctx, &sl);
    }
    ctx->slot_count++;
    return ZK_OK;
}

int zk_prove_stream_collect(zk_ctx* ctx, uint8_t* transcripts_out, uint8_t* states_out)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!transcripts_out || !states_out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (ctx->slot_count == 0) return fail(ctx, ZK_ERR_INVALID, "nothing submitted");
    CK(cudaSetDevice(ctx->device));
    zk_ctx::StreamSlot& sl = ctx->slot[ctx->slot_head];
    ctx->slot_head ^= 1;
    ctx->slot_count--;
    if (sl.worker.joinable()) sl.worker.join();        // usually long finished: its batch was submitted one collect ago
    if (sl.err != cudaSuccess) return fail_cuda(ctx, sl.err, "trace upload");
    CK(cudaStreamWaitEvent(ctx->stream, sl.ready, 0));
    return zk_prove_fibsq_resident(ctx, sl.log_n, sl.n, (const uint32_t*)sl.traces.p, transcripts_out, states_out);
}

int zk_prove_fibsq_from_trace(zk_ctx* ctx, uint32_t log_n, const uint32_t* trace, uint8_t* transcript_out, size_t cap,
                              size_t* len_out, uint8_t state_out[32])
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!trace || !transcript_out || !state_out) return fail(ctx, ZK_ERR_INVALID, "null buffer");
    if (cap < transcript_len(log_n)) return fail(ctx, ZK_ERR_BUFFER, "transcript buffer too small");
    rc = zk_prove_fibsq_batch(ctx, log_n, 1, trace, transcript_out, state_out);
    if (rc == ZK_OK && len_out) *len_out = transcript_len(log_n);
    return rc;
}

int zk_prove_fibsq(zk_ctx* ctx, uint32_t log_n, uint32_t secret, uint8_t* transcript_out, size_t cap, size_t* len_out,
                   uint8_t state_out[32])
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    std::vector<uint32_t> a(((size_t)1 << log_n) - 1);
    if (ctx->air.generic) {
        uint32_t* row = a.data();
        host_traces(ctx->air, log_n, &secret, &row, 1);
    } else {
        zk_trace_fibsq(log_n, secret, a.data());
    }
    if (!ctx->air.generic && log_n == 10 && secret == 3141592u && a[1022] != 2338775057u)  // prover.rs:42
        return fail(ctx, ZK_ERR_ASSERT, "a[1022] != 2338775057 (prover.rs:42)");
    return zk_prove_fibsq_from_trace(ctx, log_n, a.data(), transcript_out, cap, len_out, state_out);
}

// the reference's protocol with a seeded ("random", readme.md:1) initial channel state and with more
// test points than its single one, every further one drawn from the channel like the first (prover.rs:263): after the
// reference's transcript come, n_extra_queries times, the raw query word and the same 4 + log_n records for
// x = word % (8n - 16).  The evaluation arrays and heaps of the proof are still resident, so a further test point is one
// gather (zk_open_queries) and 4 + log_n host hashes.  With init_state null and n_extra_queries 0: zk_prove_fibsq.
int zk_prove_fibsq_ex(zk_ctx* ctx, uint32_t log_n, uint32_t secret, const uint8_t* init_state, uint32_t n_extra_queries,
                      uint8_t* transcript_out, size_t cap, size_t* len_out, uint8_t state_out[32])
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!transcript_out || !state_out || n_extra_queries > 65535) return fail(ctx, ZK_ERR_INVALID, "bad argument");
    const size_t base = transcript_len(log_n), dl = zk_decommit_len(log_n), total = base + (size_t)n_extra_queries * (4 + dl);
    if (cap < total) return fail(ctx, ZK_ERR_BUFFER, "transcript buffer too small (zk_transcript_len_ex)");
    if (init_state)
        for (int i = 0; i < 8; i++)
            ctx->chan_seed.w[i] = ((uint32_t)init_state[4 * i] << 24) | ((uint32_t)init_state[4 * i + 1] << 16) |
                                  ((uint32_t)init_state[4 * i + 2] << 8) | init_state[4 * i + 3];
    rc = zk_prove_fibsq(ctx, log_n, secret, transcript_out, cap, nullptr, state_out);
    ctx->chan_seed = ChanSeed{};
    if (rc) return rc;
    const uint32_t L = log_n, log_N = log_n + LOG_BLOWUP;
    const uint64_t N = (uint64_t)1 << log_N;
    std::vector<uint8_t> buf(32 + 24 + 64 * (size_t)log_N);
    uint8_t* st = state_out;
    for (uint32_t q = 0; q < n_extra_queries; q++) {
        uint8_t* out = transcript_out + base + (size_t)q * (4 + dl);
        const uint32_t f = ((uint32_t)st[0] << 24) | ((uint32_t)st[1] << 16) | ((uint32_t)st[2] << 8) | st[3];   // channel.rs:28-32
        out[0] = (uint8_t)f; out[1] = (uint8_t)(f >> 8); out[2] = (uint8_t)(f >> 16); out[3] = (uint8_t)(f >> 24);
        memcpy(buf.data(), st, 32);
        memcpy(buf.data() + 32, out, 4);
        sha::host_sha256(buf.data(), 36, st);
        const uint32_t x = (uint32_t)((uint64_t)f % (N - 16));                                                    // prover.rs:263
        if ((rc = zk_open_queries(ctx, log_n, 0, &x, 1, out + 4, dl))) return rc;
        size_t off = 4;
        for (uint32_t r = 0; r < 4 + L; r++) {                                                                    // prover.rs:274-288
            const size_t len = r < 4 ? 12 + 32 * (size_t)log_N : 24 + 64 * (size_t)(log_N - (r - 4));
            memcpy(buf.data(), st, 32);
            memcpy(buf.data() + 32, out + off, len);
            sha::host_sha256(buf.data(), 32 + len, st);
            off += len;
        }
    }
    if (len_out) *len_out = total;
    return ZK_OK;
}

// more query positions than the single one the reference draws, at positions the caller chooses.  The evaluation arrays and Merkle heaps
// of the last prove call are still resident, so further openings are one more gather.
size_t zk_decommit_len(uint32_t log_n) { return transcript_len(log_n) - tr_off_decommit(log_n); }

int zk_open_queries(zk_ctx* ctx, uint32_t log_n, uint32_t proof_index, const uint32_t* xs, uint32_t n_queries, uint8_t* out,
                    size_t out_cap)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!xs || !out || n_queries == 0 || n_queries > 65535) return fail(ctx, ZK_ERR_INVALID, "bad argument");
    if (ctx->last_nb == 0 || proof_index >= ctx->last_nb) return fail(ctx, ZK_ERR_INVALID, "no resident proof with that index");
    if (log_n != ctx->last_log_n) return fail(ctx, ZK_ERR_INVALID, "log_n differs from the resident proof's");
    const size_t n = (size_t)1 << log_n, N = n << LOG_BLOWUP, dl = zk_decommit_len(log_n);
    if (out_cap < dl * (size_t)n_queries) return fail(ctx, ZK_ERR_BUFFER, "out buffer smaller than n_queries * zk_decommit_len(log_n)");
    for (uint32_t q = 0; q < n_queries; q++)
        if (xs[q] >= N - 16) return fail(ctx, ZK_ERR_INVALID, "query position must be < 8n - 16 (prover.rs:263)");
    CK(cudaSetDevice(ctx->device));
    const size_t stride = (dl + 63) & ~(size_t)63;
    CK(ctx->s0.ensure(4 * (size_t)n_queries));
    CK(ctx->s1.ensure(stride * n_queries));
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->s0.p, xs, 4 * (size_t)n_queries, cudaMemcpyHostToDevice, s));
    const size_t p = proof_index;
    CK(launch_open_queries(s, log_n, (const uint32_t*)ctx->f_eval.p + p * N, (const uint8_t*)ctx->f_heap.p + p * 64 * N,
                           (const uint32_t*)ctx->layers.p + p * 2 * N, (const uint8_t*)ctx->layer_heaps.p + p * 128 * N,
                           (ProofState*)ctx->states.p + p, (const uint32_t*)ctx->s0.p, n_queries, (uint8_t*)ctx->s1.p, stride));
    ctx->launches += 1;
    std::vector<uint8_t> host(stride * n_queries);
    CK(cudaMemcpyAsync(host.data(), ctx->s1.p, stride * n_queries, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint32_t q = 0; q < n_queries; q++) memcpy(out + dl * q, host.data() + stride * q, dl);
    return ZK_OK;
}

// ------------------------------------------------------------------ one proof sharded over several GPUs
// Building blocks for zkstark_b200/sharded.py (one process per GPU; torch.distributed moves the data).
int zk_shard_state_init_dev(zk_ctx* ctx, void* state_dev, uint32_t a_first, uint32_t a_last)
{
    if (!ctx || !state_dev) return ZK_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(launch_state_set(ctx->stream, (ProofState*)state_dev, a_first % P, a_last % P));
    ctx->launches += 1;
    return ZK_OK;
}

int zk_shard_lde_coset_dev(zk_ctx* ctx, const uint32_t* coeffs, uint32_t log_n, uint32_t coset, uint32_t* work, uint32_t* out)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!coeffs || !work || !out || coset >= (1u << LOG_BLOWUP)) return fail(ctx, ZK_ERR_INVALID, "bad coset/buffer");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    if ((rc = get_tables(ctx, log_n + LOG_BLOWUP, &T))) return rc;
    int nl = 0;
    CK(launch_lde_one_coset(ctx->stream, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_n, coset, coeffs, work, out, &nl, nullptr));
    ctx->launches += nl;
    return ZK_OK;
}

int zk_shard_merkle_blocks_dev(zk_ctx* ctx, const uint32_t* vals_local, uint64_t local_len, uint32_t log_block, uint8_t* heap_local)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!vals_local || !heap_local || !is_pow2(local_len)) return fail(ctx, ZK_ERR_INVALID, "local_len must be a power of two");
    const uint32_t lg = ilog2_u64(local_len);
    if (lg < log_block || log_block < 1) return fail(ctx, ZK_ERR_INVALID, "local_len smaller than a block");
    CK(cudaSetDevice(ctx->device));
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p;
    a.vals = const_cast<uint32_t*>(vals_local);
    a.heap = heap_local;
    a.log_size = lg;
    int nl = 0;
    CK(launch_merkle_partial(ctx->stream, SRC_VALUES, a, 1, lg - log_block, &nl));
    ctx->launches += nl;
    return ZK_OK;
}

int zk_shard_upper_tree_dev(zk_ctx* ctx, uint8_t* upper_heap, uint32_t log_count, int post_op, uint32_t fri_layer,
                            uint32_t n_layers, void* state_dev, uint8_t* transcript_dev)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!upper_heap || log_count > 30 || post_op < POST_NONE || post_op > POST_FRI) return fail(ctx, ZK_ERR_INVALID, "bad argument");
    if (post_op != POST_NONE && (!state_dev || !transcript_dev)) return fail(ctx, ZK_ERR_INVALID, "post op needs state and transcript");
    if (post_op == POST_FRI && fri_layer >= n_layers) return fail(ctx, ZK_ERR_INVALID, "the last layer belongs to the replicated tail");
    CK(cudaSetDevice(ctx->device));
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p;
    a.heap = upper_heap;
    a.log_size = log_count;
    a.in_level = log_count;
    a.post = post_op;
    a.fri_layer = fri_layer;
    a.n_layers = n_layers;
    a.state = (ProofState*)state_dev;
    a.transcript = transcript_dev;
    int nl = 0;
    CK(launch_merkle_upper(ctx->stream, a, 1, &nl));
    ctx->launches += nl;
    return ZK_OK;
}

int zk_shard_cp_dev(zk_ctx* ctx, const uint32_t* f_local, const uint32_t* halo, uint64_t local_len, uint32_t log_n,
                    uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void* state_dev, uint32_t* cp_local)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!f_local || !halo || !state_dev || !cp_local || !is_pow2(local_len) || n_ranks == 0 || rank >= n_ranks)
        return fail(ctx, ZK_ERR_INVALID, "bad argument");
    if ((local_len * n_ranks) != ((uint64_t)1 << (log_n + LOG_BLOWUP)) || local_len < ((uint64_t)1 << log_block) || log_block < 5)
        return fail(ctx, ZK_ERR_INVALID, "local_len * n_ranks must be 8n and hold at least one block of >= 32");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    if ((rc = get_tables(ctx, log_n + LOG_BLOWUP, &T))) return rc;
    CpArgs c = make_cp_args(ctx, log_n);
    c.f = f_local;
    c.halo = halo;
    c.log_len = ilog2_u64(local_len);
    c.shard = ShardMap{n_ranks, rank, log_block, 0};
    c.alpha_from_state = 1;
    CK(launch_cp(ctx->stream, cp_local, c, T->h.tab, 1, 0, (const ProofState*)state_dev));
    ctx->launches += 1;
    return ZK_OK;
}

int zk_shard_fold_dev(zk_ctx* ctx, const uint32_t* cur_local, uint64_t local_len, uint32_t layer_k, uint32_t log_N0,
                      uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void* state_dev, uint32_t* next_local)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!cur_local || !next_local || !state_dev || !is_pow2(local_len) || local_len < 2 || n_ranks == 0 || rank >= n_ranks || log_N0 > 30)
        return fail(ctx, ZK_ERR_INVALID, "bad argument");
    if (n_ranks > 1 && (local_len >> 1) < ((uint64_t)1 << log_block))
        return fail(ctx, ZK_ERR_INVALID, "a sharded fold needs at least two blocks per rank");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    int rc = get_tables(ctx, log_N0, &T);
    if (rc) return rc;
    FoldArgs f = make_fold_args(layer_k);
    f.cur = cur_local;
    f.beta_from_state = 1;
    f.shard = ShardMap{n_ranks, rank, log_block, 0};
    CK(launch_fold(ctx->stream, cur_local, next_local, ilog2_u64(local_len) - 1, f, T->h.tab, 1, 0, (const ProofState*)state_dev));
    ctx->launches += 1;
    return ZK_OK;
}

// zk_shard_cp_dev + zk_shard_merkle_blocks_dev in one pass (north_star item 3 fused into the CP tree's leaf pass).
int zk_shard_cp_blocks_dev(zk_ctx* ctx, const uint32_t* f_local, const uint32_t* halo, uint64_t local_len, uint32_t log_n,
                           uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void* state_dev, uint32_t* cp_local,
                           uint8_t* heap_local)
{
    return zkh::shard_cp_blocks(ctx, f_local, halo, 0, local_len, log_n, n_ranks, rank, log_block, state_dev, cp_local, heap_local, nullptr);
}
}  // extern "C"

// halo_roll != 0: the halo rows are the heads of rank + 1's local blocks as they are (no re-ordering): on the last rank the
// block that follows local block bl is rank 0's block bl + 1 (wrapping around the whole domain)
int zkh::shard_cp_blocks(zk_ctx* ctx, const uint32_t* f_local, const uint32_t* halo, uint32_t halo_roll, uint64_t local_len,
                         uint32_t log_n, uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void* state_dev,
                         uint32_t* cp_local, uint8_t* heap_local, DevBuf* tab_cache)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!f_local || !halo || !state_dev || !cp_local || !heap_local || !is_pow2(local_len) || n_ranks == 0 || rank >= n_ranks)
        return fail(ctx, ZK_ERR_INVALID, "bad argument");
    if ((local_len * n_ranks) != ((uint64_t)1 << (log_n + LOG_BLOWUP)) || local_len < ((uint64_t)1 << log_block) || log_block < 5)
        return fail(ctx, ZK_ERR_INVALID, "local_len * n_ranks must be 8n and hold at least one block of >= 32");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    if ((rc = get_tables(ctx, log_n + LOG_BLOWUP, &T))) return rc;
    const uint32_t lg = ilog2_u64(local_len);
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p;
    a.state = (ProofState*)const_cast<void*>(state_dev);
    a.pow = T->h.tab;
    a.cp = make_cp_args(ctx, log_n);
    a.cp.f = f_local;
    a.cp.halo = halo;
    if (halo_roll) {
        a.cp.halo_roll = halo_roll;
        a.cp.halo_mask = (uint32_t)(local_len >> log_block) - 1u;
    }
    a.cp.log_len = lg;
    a.cp.shard = ShardMap{n_ranks, rank, log_block, 0};
    a.cp.alpha_from_state = 1;
    a.vals = cp_local;
    a.heap = heap_local;
    a.log_size = lg;
    int nl = 0;
    if (tab_cache && ctx->use_cp_tab) {
        // rank-local slice of the per-size constant tables (CpArgs::tab), in local order: 12 bytes per local point, built
        // once per (communicator, size) -- the caller keeps tab_cache for exactly one geometry
        if (!tab_cache->p && tab_cache->ensure(12 * local_len) == cudaSuccess) {
            CK(launch_cp_tables(ctx->stream, (uint32_t*)tab_cache->p, a.cp, T->h.tab));
            nl++;
        }
        cudaGetLastError();
        a.cp.tab = (const uint32_t*)tab_cache->p;   // null when the allocation failed: the per-point form is used
    }
    CK(launch_merkle_partial(ctx->stream, SRC_CP, a, 1, lg - log_block, &nl));
    ctx->launches += nl;
    return ZK_OK;
}
extern "C" {

// zk_shard_fold_dev + zk_shard_merkle_blocks_dev in one pass: the leaf pass of the block subtrees folds the previous
// layer first (north_star item 4: fold, beta-mix and next-layer leaf hashing fused), so the new layer is written once
// and hashed from L1 instead of being re-read by a second kernel.
int zk_shard_fold_blocks_dev(zk_ctx* ctx, const uint32_t* cur_local, uint64_t local_len, uint32_t layer_k, uint32_t log_N0,
                             uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void* state_dev,
                             uint32_t* next_local, uint8_t* heap_local)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!cur_local || !next_local || !heap_local || !state_dev || !is_pow2(local_len) || local_len < 2 || n_ranks == 0 ||
        rank >= n_ranks || log_N0 > 30)
        return fail(ctx, ZK_ERR_INVALID, "bad argument");
    const uint32_t lg_next = ilog2_u64(local_len) - 1;
    if (log_block < 5 || lg_next < log_block)
        return fail(ctx, ZK_ERR_INVALID, "the folded layer must hold at least one block of >= 32 values per rank");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    int rc = get_tables(ctx, log_N0, &T);
    if (rc) return rc;
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p;
    a.state = (ProofState*)const_cast<void*>(state_dev);
    a.pow = T->h.tab;
    a.fold = make_fold_args(layer_k);
    a.fold.cur = cur_local;
    a.fold.log_half = lg_next;
    a.fold.beta_from_state = 1;
    a.fold.shard = ShardMap{n_ranks, rank, log_block, 0};
    a.vals = next_local;
    a.heap = heap_local;
    a.log_size = lg_next;
    int nl = 0;
    CK(launch_merkle_partial(ctx->stream, SRC_FOLD, a, 1, lg_next - log_block, &nl));
    ctx->launches += nl;
    return ZK_OK;
}

// Replicated tail: layer `first_layer` (its tree already committed) is complete in layers_arena[0 .. N>>first_layer);
// folds + trees + channel steps of layers first_layer+1 .. L follow, appended in the arenas.
int zk_shard_fri_tail_dev(zk_ctx* ctx, uint32_t log_n, uint32_t first_layer, uint32_t* layers_arena, uint8_t* heaps_arena,
                          void* state_dev, uint8_t* transcript_dev)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    const uint32_t L = log_n, log_N = log_n + LOG_BLOWUP;
    if (!layers_arena || !heaps_arena || !state_dev || !transcript_dev || first_layer >= L) return fail(ctx, ZK_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(ctx->device));
    SizeTables* T;
    if ((rc = get_tables(ctx, log_N, &T))) return rc;
    TreeArgs a{};
    a.tune = &ctx->tune;
    a.counters = (uint32_t*)ctx->counters.p;
    a.state = (ProofState*)state_dev;
    a.transcript = transcript_dev;
    a.n_layers = L;
    a.pow = T->h.tab;
    int nl = 0;
    size_t voff = 0, hoff = 0;
    for (uint32_t i = first_layer + 1; i <= L; i++) {
        const size_t prev_len = (size_t)1 << (log_N - (i - 1));
        a.fold = make_fold_args(i - 1);
        a.fold.cur = layers_arena + voff;
        a.fold.log_half = log_N - i;
        a.fold.beta_from_state = 1;
        voff += prev_len;
        hoff += 64 * prev_len;
        a.vals = layers_arena + voff;
        a.heap = heaps_arena + hoff;
        a.log_size = log_N - i;
        a.post = POST_FRI;
        a.fri_layer = i;
        if (a.log_size <= TILE_MIN_LOG && ctx->fri_tail) {
            CK(launch_fri_tail(ctx->stream, a, 1, L));
            nl++;
            break;
        }
        CK(launch_merkle_tree(ctx->stream, SRC_FOLD, a, 1, &nl));
    }
    ctx->launches += nl;
    return ZK_OK;
}

// prover.rs:263-289 for the sharded proof: see shard_decommit_kernel (tree.cu).
int zk_shard_decommit_dev(zk_ctx* ctx, uint32_t log_n, uint32_t n_ranks, uint32_t rank, uint32_t log_block, uint32_t first_tail,
                          const void* const* tree_ptrs, const uint32_t* tail_vals, const uint8_t* tail_heaps,
                          const void* state_dev, uint8_t* transcript_dev)
{
    int rc = check_log_n(ctx, log_n);
    if (rc) return rc;
    if (!tree_ptrs || !state_dev || !transcript_dev || n_ranks == 0 || (n_ranks & (n_ranks - 1)) || rank >= n_ranks ||
        first_tail >= log_n || first_tail + 2 > SHARD_MAX_TREES || log_block < 1)
        return fail(ctx, ZK_ERR_INVALID, "bad argument");
    if (first_tail + 1 < log_n && (!tail_vals || !tail_heaps)) return fail(ctx, ZK_ERR_INVALID, "tail arenas missing");
    CK(cudaSetDevice(ctx->device));
    ShardDecommitArgs a{};
    a.log_n = log_n; a.n_ranks = n_ranks; a.rank = rank; a.log_block = log_block; a.first_tail = first_tail;
    for (uint32_t t = 0; t < first_tail + 2; t++) {
        a.vals[t] = (const uint32_t*)tree_ptrs[3 * t];
        a.heap[t] = (const uint8_t*)tree_ptrs[3 * t + 1];
        a.upper[t] = (const uint8_t*)tree_ptrs[3 * t + 2];
        if (!a.vals[t] || !a.heap[t] || !a.upper[t]) return fail(ctx, ZK_ERR_INVALID, "null tree pointer");
    }
    a.tail_vals = tail_vals; a.tail_heaps = tail_heaps;
    a.state = (const ProofState*)state_dev; a.out = transcript_dev;
    a.flag_off = (uint32_t)((transcript_len(log_n) + 3) & ~(size_t)3);
    CK(launch_shard_decommit(ctx->stream, a));
    ctx->launches += 1;
    return ZK_OK;
}

// ------------------------------------------------------------------ K1 on its own (tests)
int zk_field_ops(zk_ctx* ctx, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (!a || !b || !out || n == 0 || n > ((size_t)1 << 28)) return fail(ctx, ZK_ERR_INVALID, "bad argument");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->s0.ensure(8 * n));
    CK(ctx->s1.ensure(24 * n));
    uint32_t* da = (uint32_t*)ctx->s0.p;
    CK(cudaMemcpyAsync(da, a, 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(da + n, b, 4 * n, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_field_ops(ctx->stream, da, da + n, n, (uint32_t*)ctx->s1.p));
    ctx->launches += 1;
    CK(cudaMemcpyAsync(out, ctx->s1.p, 24 * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ZK_OK;
}

// ------------------------------------------------------------------ measurement support
int zk_bench_set(zk_ctx* ctx, int sha_variant, uint32_t leaf_pass_k)
{
    if (!ctx) return ZK_ERR_INVALID;
    if (sha_variant == -100 || sha_variant == -101) {   // composition polynomial: -100 per-point form, -101 constant tables
        ctx->use_cp_tab = (sha_variant == -101);
        return ZK_OK;
    }
    if (sha_variant == -110 || sha_variant == -111) {   // batches: -110 one lane, -111 two lanes on two streams
        ctx->lanes = (sha_variant == -111);
        return ZK_OK;
    }
    ctx->bench_variant = sha_variant;
    ctx->bench_k = leaf_pass_k & 0xffu;
    if (leaf_pass_k >> 8) ctx->tune.coop_cnt = (leaf_pass_k >> 8) - 1;   // bits 8..: 1 + widest cooperative tile level (0: keep)
    return ZK_OK;
}

int zk_bench_set_tile_direct_log(zk_ctx* ctx, uint32_t log_leaves)
{
    if (!ctx || (log_leaves & 0xffu) > 2 * TILE_LOG || ((log_leaves >> 8) & 0xffu) > TILE_LOG) return ZK_ERR_INVALID;
    set_tile_direct_log(ctx->tune, log_leaves);
    return ZK_OK;
}

int zk_bench_kernel(zk_ctx* ctx, int which, uint32_t log_size, uint32_t iters, int flush_l2, float ms_out[3])
{
    if (!ctx || !ms_out || iters == 0 || iters > 1000) return ZK_ERR_INVALID;
    if (which == 7) {  // single-CTA node-hash latency: log_size = warps, iters = chain length; ms_out[2] = cycles per node
        CK(cudaSetDevice(ctx->device));
        CK(ctx->s0.ensure(4096));
        CK(cudaMemsetAsync(ctx->s0.p, 1, 4096, ctx->stream));
        long long* cyc = (long long*)((char*)ctx->s0.p + 2048);
        CK(launch_node_latency(ctx->stream, (uint32_t*)ctx->s0.p, iters, log_size ? log_size : 1, ctx->bench_variant, cyc));
        long long h = 0;
        CK(cudaMemcpyAsync(&h, cyc, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ms_out[0] = ms_out[1] = 0.f;
        ms_out[2] = (float)((double)h / iters);
        ctx->launches += 1;
        return ZK_OK;
    }
    if (which < 0 || (which > 6 && which != 8)) return fail(ctx, ZK_ERR_INVALID, "unknown kernel id");
    if (which != 6 && which != 8 && (log_size < 3 || log_size > 30)) return fail(ctx, ZK_ERR_INVALID, "log_size out of range");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int rc = ZK_OK;
    const size_t flush_bytes = (size_t)512 << 20;
    if (flush_l2) CK(ctx->flushbuf.ensure(flush_bytes));
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, ctx->device);
    SizeTables* T = nullptr;
    size_t n_in = (size_t)1 << log_size, N8 = n_in << LOG_BLOWUP;
    TreeArgs ta{};
    ta.tune = &ctx->tune;
    ta.counters = (uint32_t*)ctx->counters.p;
    FoldArgs fa{};
    CpArgs ca{};
    switch (which) {
        case 0: case 1:
            CK(ctx->s0.ensure(4 * n_in)); CK(ctx->s1.ensure(64 * n_in));
            CK(bench_fill(s, (uint32_t*)ctx->s0.p, n_in, 1));
            ta.vals = (uint32_t*)ctx->s0.p; ta.heap = (uint8_t*)ctx->s1.p; ta.log_size = log_size; ta.post = POST_NONE;
            break;
        case 2:
            if ((rc = check_log_n(ctx, log_size))) return rc;
            if ((rc = get_tables(ctx, log_size + LOG_BLOWUP, &T))) return rc;
            CK(ctx->s0.ensure(4 * n_in)); CK(ctx->s1.ensure(4 * N8)); CK(ctx->s2.ensure(4 * N8));
            CK(bench_fill(s, (uint32_t*)ctx->s0.p, n_in, 2));
            break;
        case 3:
            if ((rc = check_log_n(ctx, log_size))) return rc;
            if ((rc = get_tables(ctx, log_size + LOG_BLOWUP, &T))) return rc;
            CK(ctx->s0.ensure(4 * n_in)); CK(ctx->s1.ensure(4 * n_in));
            break;
        case 4:
            if ((rc = get_tables(ctx, log_size, &T))) return rc;
            CK(ctx->s0.ensure(4 * n_in)); CK(ctx->s1.ensure(2 * n_in));
            CK(bench_fill(s, (uint32_t*)ctx->s0.p, n_in, 4));
            fa = make_fold_args(0); fa.cur = (const uint32_t*)ctx->s0.p; fa.beta_raw = 0x9abcdef1u;
            break;
        case 5:
            if ((rc = check_log_n(ctx, log_size))) return rc;
            if ((rc = get_tables(ctx, log_size + LOG_BLOWUP, &T))) return rc;
            CK(ctx->s0.ensure(4 * N8)); CK(ctx->s1.ensure(4 * N8));
            CK(bench_fill(s, (uint32_t*)ctx->s0.p, N8, 5));
            ca = make_cp_args(ctx, log_size); ca.f = (const uint32_t*)ctx->s0.p;
            ca.alpha_raw[0] = 361545003u; ca.alpha_raw[1] = 3235878091u; ca.alpha_raw[2] = 2708123352u;
            ca.a_first = 1; ca.a_last = 2338775057u;
            CK(attach_cp_tab(ctx, T, ca, nullptr));
            break;
        default:
            CK(ctx->s0.ensure(4096));
            break;
    }
    std::vector<cudaEvent_t> ev(2 * (size_t)iters);
    for (auto& e : ev) CK(cudaEventCreate(&e));
    double ops = 0;
    int nl = 0;
    for (uint32_t it = 0; it < iters + 1 && rc == ZK_OK; it++) {  // iteration 0 is an untimed warm-up
        if (which == 3) {  // the inverse transform works in place: refill its input every time
            CK(bench_fill(s, (uint32_t*)ctx->s0.p, n_in, 3 + it));
        }
        if (flush_l2) CK(bench_flush(s, ctx->flushbuf.p, flush_bytes));
        if (it) CK(cudaEventRecord(ev[2 * (it - 1)], s));
        cudaError_t e = cudaSuccess;
        switch (which) {
            case 0: {
                TreeArgs a = ta; a.in_level = log_size;
                a.k = ctx->bench_k ? ctx->bench_k : (log_size > 15 ? 5 : 1);
                // the production launch shape of a single proof: one CTA epilogue level (tree.cu pick_ep)
                a.ep_levels = (ctx->tune.pass_epilogue && log_size >= a.k + 8) ? 1 : 0;
                e = launch_tree_leaf_pass(s, a, 1, ctx->bench_variant); nl++;
                break;
            }
            case 1: e = launch_merkle_tree(s, SRC_VALUES, ta, 1, &nl); break;
            case 2: e = launch_lde(s, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_size, (const uint32_t*)ctx->s0.p, 0,
                                   (uint32_t*)ctx->s2.p, N8, (uint32_t*)ctx->s1.p, N8, 1, &nl); break;
            case 3: e = launch_interpolate(s, ctx->tw_fwd, ctx->tw_inv, T->h.tab, T->w.tab, log_size, true,
                                           (uint32_t*)ctx->s0.p, n_in, (uint32_t*)ctx->s1.p, n_in, 1, &nl,
                                           ctx->bench_variant == -200 ? nullptr : phantom_acc(ctx)); break;
            case 4: e = launch_fold(s, (const uint32_t*)ctx->s0.p, (uint32_t*)ctx->s1.p, log_size - 1, fa, T->h.tab, 1, 0); nl++; break;
            case 5: e = launch_cp(s, (uint32_t*)ctx->s1.p, ca, T->h.tab, 1, 0); nl++; break;
            case 8: ops = bench_bfly(s, (uint32_t*)ctx->s0.p, n_sm, (int)log_size); e = cudaGetLastError(); nl++; break;
            default: ops = bench_intpipe(s, (uint32_t*)ctx->s0.p, n_sm, (int)log_size); e = cudaGetLastError(); nl++; break;
        }
        if (e != cudaSuccess) { rc = fail_cuda(ctx, e, "bench launch"); break; }
        if (it) CK(cudaEventRecord(ev[2 * (it - 1) + 1], s));
    }
    if (rc == ZK_OK) {
        CK(cudaStreamSynchronize(s));
        if (which == 2 || which == 3) {   // ZK_NTT_TIMING=1: per-pass times of the last transform on stderr
            float pm[4];
            const int np = ntt_pass_times(pm);
            if (np > 0) {
                fprintf(stderr, "[zk] %s 2^%u passes:", which == 2 ? "lde" : "interpolate", log_size);
                for (int i = 0; i < np; i++) fprintf(stderr, " %.1f us", pm[i] * 1e3);
                fprintf(stderr, "\n");
            }
        }
        double sum = 0, mn = 1e30;
        for (uint32_t it = 0; it < iters; it++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[2 * it], ev[2 * it + 1]);
            sum += ms;
            if (ms < mn) mn = ms;
        }
        ms_out[0] = (float)(sum / iters);
        ms_out[1] = (float)mn;
        ms_out[2] = (which == 6 || which == 8) ? (float)(ops / (mn * 1e-3) / 1e12) : 0.f;   // Tiop/s, or tera-butterflies/s
    }
    for (auto& e : ev) cudaEventDestroy(e);
    ctx->launches += nl;
    return rc;
}

}  // extern "C"
