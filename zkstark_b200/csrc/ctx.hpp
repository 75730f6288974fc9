// Private host-side definitions shared by the C-ABI translation units (capi.cu, sharded.cu): the context
// (device buffers, tables, streams), error helpers and the per-size constant builders.  Not installed.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <thread>
#include <vector>

#include "../../include/zkstark_b200.h"
#include "common.cuh"
#include "hostair.hpp"

// launchers of ntt.cu / trace.cu / benchsupport.cu (tree.cu's are in common.cuh)
namespace zk {
cudaError_t launch_interpolate(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow,
                               PowTab wpow, uint32_t log_n, bool phantom, uint32_t* y, uint64_t y_stride,
                               uint32_t* coeffs, uint64_t coeffs_stride, uint32_t batch, int* n_launches,
                               unsigned long long* acc);
cudaError_t launch_lde(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow, PowTab wpow,
                       uint32_t log_n, const uint32_t* coeffs, uint64_t coeffs_stride, uint32_t* work,
                       uint64_t work_stride, uint32_t* evals, uint64_t evals_stride, uint32_t batch,
                       int* n_launches);
cudaError_t launch_lde_one_coset(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow, PowTab wpow,
                                 uint32_t log_n, uint32_t r, const uint32_t* coeffs, uint32_t* work, uint32_t* out,
                                 int* n_launches, const NttPack* pack);
cudaError_t launch_trace_batch(cudaStream_t s, const uint32_t* secrets_dev, uint32_t* traces, uint64_t stride,
                               uint32_t log_n, uint32_t batch, const TraceAir& air);
cudaError_t bench_fill(cudaStream_t s, uint32_t* p, size_t n, uint32_t seed);
cudaError_t bench_flush(cudaStream_t s, void* p, size_t bytes);
double bench_intpipe(cudaStream_t s, uint32_t* out, int n_sm, int mode);
void host_traces(const HostAir& air, uint32_t log_n, const uint32_t* secrets, uint32_t* const* rows, uint32_t count);
uint32_t host_trace_group(uint32_t n, uint32_t threads, bool generic);   // traces per group: 4 or 8
int host_trace_isa();                                                    // 0 scalar, 1 AVX2, 2 AVX-512 IFMA
int ntt_pass_times(float ms[4]);
size_t ntt_table_words();
cudaError_t ntt2_build_tables(cudaStream_t s, uint32_t* tw_fwd, uint32_t* tw_inv);
double bench_bfly(cudaStream_t s, uint32_t* out, int n_sm, int mode);
cudaError_t launch_field_ops(cudaStream_t s, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out);
}  // namespace zk

namespace zkh {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    // grows only; the old allocation is released only once the new one exists, so a failed call leaves the
    // buffer as it was
    cudaError_t ensure(size_t need)
    {
        if (need <= bytes) return cudaSuccess;
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, need);
        if (e != cudaSuccess) {
            // second chance: give the old block back first (the device may simply be full)
            cudaGetLastError();
            if (!p) return e;
            cudaFree(p);
            p = nullptr;
            bytes = 0;
            e = cudaMalloc(&q, need);
            if (e != cudaSuccess) return e;
        } else if (p) {
            cudaFree(p);
        }
        p = q;
        bytes = need;
        return cudaSuccess;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
};

struct PowTabOwned {
    uint32_t* lo = nullptr;
    uint32_t* hi = nullptr;
    zk::PowTab tab{};
};

struct SizeTables {
    PowTabOwned h;  // powers of h = 5^((p-1)/N)
    PowTabOwned w;  // powers of 5
    DevBuf cp_tab;  // CpArgs::tab of the size log_N - 3 (built on first use by the prove pipeline)
};

constexpr uint32_t MAX_BATCH = 256;  // proofs per sub-batch (one tile counter each)

}  // namespace zkh

struct zk_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;
    uint32_t* tw_fwd = nullptr;
    uint32_t* tw_inv = nullptr;
    std::map<uint32_t, zkh::SizeTables> tabs;  // by log_N
    std::map<uint64_t, zkh::PowTabOwned> shift_tabs;  // powers of a caller-chosen coset shift, by (shift << 8 | log_N)
    zkh::DevBuf s0, s1, s2;                    // scratch for the host-buffer stage calls
    zkh::DevBuf flushbuf;
    zkh::DevBuf counters;                      // tile-arrival counters of the tree kernels (one per proof of a batch)
    int bench_variant = -1;
    uint32_t bench_k = 0;
    bool use_cp_tab = true;               // composition polynomial from per-size constant tables (bench knob)
    zk::ChanSeed chan_seed{};             // initial channel state of the next proofs (zk_prove_fibsq_ex sets and clears it)
    bool fri_tail = true;                 // the single-tile FRI layers of a proof run in one launch (ZK_NO_FRI_TAIL=1: one each)
    bool lanes = true;                    // batches run as two lanes on two streams (bench knob)
    zk::TreeTune tune;                    // launch-shape knobs of the tree kernels (per context; zk_bench_set*)
    cudaStream_t lane_stream = nullptr;   // side stream of the second lane
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // prove work space
    uint32_t w_log_n = 0, w_batch = 0;
    uint32_t last_log_n = 0, last_nb = 0;  // the sub-batch whose arrays and heaps are still resident (zk_open_queries)
    zkh::DevBuf traces, coeffs, work, f_eval, layers, f_heap, layer_heaps, states, transcripts;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;
    size_t mem_budget = 0;
    bool stage_timing = false;
    float stage_ms[7] = {0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    // pipelined secrets -> proofs stream (zk_prove_stream_*)
    struct StreamSlot {
        zkh::DevBuf traces, secrets;
        std::vector<uint32_t> host_secrets;
        void* pinned = nullptr;            // host traces of the batch (rows of n words), written by the worker threads
        size_t pinned_bytes = 0;
        std::thread worker;                // produces + uploads the traces; joined by collect
        cudaError_t err = cudaSuccess;
        cudaEvent_t ready = nullptr;       // recorded on trace_stream behind the last upload
        uint32_t log_n = 0, n = 0;
    };
    StreamSlot slot[2];
    uint32_t slot_head = 0, slot_count = 0;
    cudaStream_t trace_stream = nullptr;
    zk::HostAir air;                       // the program being proven (zk_ctx_set_air; default: the reference's FibonacciSq)
    zk::AirCoef air_dev{};                 //   its transition constraint as the CP kernels take it
    zk_air air_public{1, 1, 1, 0, 0, 0, 0, 0};
    bool trace_on_device = false;          // ZK_TRACE_DEVICE=1
    uint32_t host_threads = 0;             // ZK_HOST_THREADS (0: all hardware threads)
};

namespace zkh {

int fail(zk_ctx* c, int code, const std::string& msg);
int fail_cuda(zk_ctx* c, cudaError_t e, const char* what);
#define CK(call)                                                          \
    do {                                                                  \
        cudaError_t e__ = (call);                                         \
        if (e__ != cudaSuccess) return zkh::fail_cuda(ctx, e__, #call);   \
    } while (0)

cudaError_t build_powtab(PowTabOwned& t, uint32_t base, uint32_t log_E, uint32_t log_N_tag);
void free_powtab(PowTabOwned& t);
unsigned long long* phantom_acc(zk_ctx* ctx);
int get_tables(zk_ctx* ctx, uint32_t log_N, SizeTables** out);
bool is_pow2(uint64_t v);
uint32_t ilog2_u64(uint64_t v);
zk::FoldArgs make_fold_args(uint32_t layer_k);
zk::CpArgs make_cp_args(const zk_ctx* ctx, uint32_t log_n);
cudaError_t attach_cp_tab(zk_ctx* ctx, SizeTables* T, zk::CpArgs& c, int* nl);
void finish_channel_host(uint32_t log_n, const zk::ProofState& ps, const uint8_t* tr, uint8_t state_out[32]);
int check_log_n(zk_ctx* ctx, uint32_t log_n);
int shard_cp_blocks(zk_ctx* ctx, const uint32_t* f_local, const uint32_t* halo, uint32_t halo_roll, uint64_t local_len,
                    uint32_t log_n, uint32_t n_ranks, uint32_t rank, uint32_t log_block, const void* state_dev,
                    uint32_t* cp_local, uint8_t* heap_local, DevBuf* tab_cache);

}  // namespace zkh
