// K2 -- interpolation (iNTT over the trace subgroup) and coset low-degree extension.
//
// Replaces the reference's O(n^3) `lagrange` (polynomial.rs:337-383, called at prover.rs:60-61)
// and the O(N n) per-point `solve` loop (polynomial.rs:49-56, prover.rs:69-70).  Same values,
// natural order in and out (leaf i <-> domain point 5 h^i, prover.rs:69-70,81).
//
// A length-M = 2^mu transform is done in ceil(mu/10) passes of CTA-local radix-R (R = 2^m <= 1024)
// decimation-in-frequency transforms (multi-pass "four-step" without explicit transposes):
//   strided pass : M_sub = R*S.  X_sub[c + R d] = sum_b [ (sum_a x[a S + b] w_R^(a c)) w_Msub^(b c) ] w_S^(b d)
//                  -> CTA owns 32 consecutive b (coalesced 128 B rows), all a; in-place: (a,b) -> (c,b).
//   final pass   : contiguous length-R blocks; the CTA takes 32 rows whose OUTPUT indices are
//                  consecutive, transposes through shared memory and writes natural order.
// In shared memory the tile is [R][33]: lanes always run along the independent dimension, so
// every butterfly is conflict-free and its twiddle is warp-uniform (broadcast from smem).
// LDE = 8 size-n coset transforms  f_eval[8 j + r] = NTT_n( c_k (5 h^r)^k )[j]  (h^8 = g),
// written interleaved so stores stay contiguous.
#include <cstdlib>

#include "common.cuh"

namespace zk {

struct NttPassArgs {
    const uint32_t* src; uint64_t src_stride;   // per proof (elements)
    uint32_t* dst; uint64_t dst_stride;
    uint32_t log_M;      // transform length
    uint32_t log_S;      // stride of this pass (strided kernel)
    uint32_t log_Msub;   // log_R + log_S
    int inverse;
    int use_scale;
    uint32_t scale_m;    // Montgomery constant multiplied into the final outputs (n^-1)
    const uint32_t* tw;  // w_R^j (or inverse), j < R/2, Montgomery form
    const uint32_t* tw2; // the same radix's twiddles in the narrow-tile kernels' per-thread order (ntt2_build_tables)
    PowTab hpow;         // powers of h, order N = 2^hpow.log_N  (M divides N)
    PowTab wpow;         // powers of 5, exponents < 2^wpow.log_N
    uint32_t log_R1;     // final pass: first-pass radix (0 when single pass)
    uint32_t log_Q2;     // final pass: second-pass radix (0 unless three passes)
    uint32_t coset_base; // index of the first coset of this launch (sharded LDE: one coset per launch)
    int use_coset;       // single-pass narrow-tile kernel: scale the input by sigma_r^k
    NttPack pack;        // final pass of a one-coset launch: store straight into the all-to-all send layout
};

constexpr int TP = 33;  // padded tile pitch

// Everything that shapes the inner loops is a compile-time constant of the radix: the block size, the rows a
// warp moves in the load/store phases, the register-blocked rounds of the CTA-local transform and all their
// shared-memory offsets (immediates in the LDS/STS).  The first version of these kernels took the radix at run
// time and spent ~60 % of its 457 instructions per point (ncu, profiles/r1c) on index arithmetic and on
// recomputing twiddles from the power tables; this one issues ~190.
template <int LOG_R>
struct NttCfg {
    static constexpr int R = 1 << LOG_R;
    static constexpr int NW = (R / 16) < 2 ? 2 : ((R / 16) > 32 ? 32 : (R / 16));  // warps per CTA
    static constexpr int THREADS = 32 * NW;
    static constexpr int RW = R / NW;  // rows per warp in the load / store phases (4 .. 32)
};

__device__ __forceinline__ uint32_t pw(const PowTab& t, uint32_t e)
{
    return mont_mul(__ldg(t.lo + (e & ((1u << t.lo_bits) - 1u))), __ldg(t.hi + (e >> t.lo_bits)));
}

// One register-blocked round of the in-place DIF: stages S0 .. S0+LR-1 of the size-R transform along
// the row index of sm[R][TP].  A thread owns one column (its lane) of one radix-2^LR group: it pulls
// the 2^LR rows into registers, runs LR butterfly stages there, and writes them back -- one shared
// memory round trip per LR stages.  Twiddles are warp-uniform (they depend on the row group only):
// stage S0+t, in-group position j uses w_R^((j*stride + lo) << (S0+t)), a broadcast LDS at an immediate
// offset from one per-stage base.  In the last round (stride 1) the j = 0 twiddles are 1 and are skipped.
template <int LOG_R, int S0, int LR>
__device__ __forceinline__ void dif_round(uint32_t* sm, const uint32_t* tws, uint32_t warp, uint32_t lane)
{
    constexpr int RAD = 1 << LR, LOG_STRIDE = LOG_R - S0 - LR, N_MINI = 1 << (LOG_R - LR);
    constexpr int NW = NttCfg<LOG_R>::NW, TRIPS = (N_MINI + NW - 1) / NW;
#pragma unroll
    for (int i = 0; i < TRIPS; i++) {
        const uint32_t q = warp + i * NW;
        if ((N_MINI % NW) == 0 || q < (uint32_t)N_MINI) {
            const uint32_t lo = q & ((1u << LOG_STRIDE) - 1u), hi = q >> LOG_STRIDE;
            uint32_t* p = sm + ((hi << (LOG_R - S0)) + lo) * TP + lane;
            uint32_t x[RAD];
#pragma unroll
            for (int d = 0; d < RAD; d++) x[d] = p[(d << LOG_STRIDE) * TP];
#pragma unroll
            for (int t = 0; t < LR; t++) {
                const int half = RAD >> (t + 1);
                const uint32_t* tp = tws + (lo << (S0 + t));
#pragma unroll
                for (int j = 0; j < half; j++) {
                    const bool trivial = (LOG_STRIDE == 0 && j == 0);
                    const uint32_t tw = trivial ? 0u : tp[j << (LOG_STRIDE + S0 + t)];
#pragma unroll
                    for (int blk = 0; blk < RAD / (2 * half); blk++) {
                        const int i0 = blk * 2 * half + j, i1 = i0 + half;
                        const uint32_t u = x[i0], v = x[i1];
                        x[i0] = fadd(u, v);
                        x[i1] = trivial ? fsub(u, v) : mont_mul(fsub(u, v), tw);
                    }
                }
            }
#pragma unroll
            for (int d = 0; d < RAD; d++) p[(d << LOG_STRIDE) * TP] = x[d];
        }
    }
    __syncthreads();
}

// In-place DIF over the row index of sm[R][TP], all 32 columns; leaves X[c] at row brev(c).
template <int LOG_R, int S0 = 0>
__device__ __forceinline__ void cta_dif(uint32_t* sm, const uint32_t* tws, uint32_t warp, uint32_t lane)
{
    if constexpr (S0 < LOG_R) {
        constexpr int rem = LOG_R - S0;
        constexpr int LR = rem > 4 ? (rem == 5 ? 3 : 4) : rem;
        dif_round<LOG_R, S0, LR>(sm, tws, warp, lane);
        cta_dif<LOG_R, S0 + LR>(sm, tws, warp, lane);
    }
}

template <int BITS>
__device__ __forceinline__ uint32_t brev_c(uint32_t x) { return BITS ? (__brev(x) >> (32 - BITS)) : 0u; }

// Montgomery form of sigma^e, sigma = 5 h^rr: the coset shift of coset rr (prover.rs:69)
__device__ __forceinline__ uint32_t sigma_pow_m(const NttPassArgs& a, uint32_t rr, uint32_t e)
{
    uint32_t v = pw(a.wpow, e);
    if (rr) v = mont_mul(v, pw(a.hpow, (uint32_t)(((uint64_t)rr * e) & ((1u << a.hpow.log_N) - 1u))));
    return v;
}

// Strided pass: M_sub = R*S.  The CTA owns 32 consecutive columns b and all R rows a of x[a S + b]:
// load (with the coset scaling c_k sigma^k fused into the first LDE pass), CTA-local DIF, multiply by
// w_Msub^(b c), store (c, b) in place.  Each warp moves a contiguous block of RW rows; the per-row
// factors sigma^(k0 + u S) and w^(b (c0 + u)) advance along four independent multiplicative chains, so a
// point costs two multiplications instead of two table look-ups (4 loads + 3 multiplications).
template <int LOG_R, bool COSET>
__global__ void __launch_bounds__(NttCfg<LOG_R>::THREADS) ntt_strided_kernel(const NttPassArgs a)
{
    pdl_enter();
    using C = NttCfg<LOG_R>;
    constexpr int R = C::R, RW = C::RW;
    extern __shared__ uint32_t smem[];
    uint32_t* sm = smem;
    uint32_t* tws = smem + R * TP;
    uint32_t* su = tws + R / 2;  // sigma^(m S), m = 1..4 (COSET)
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t log_tpb = a.log_S - 5;                        // tiles of 32 columns per block of M_sub
    const uint32_t log_per_array = a.log_M - LOG_R - 5;
    uint32_t tile = blockIdx.x;
    const uint32_t r = tile >> log_per_array;                    // coset
    tile &= (1u << log_per_array) - 1u;
    const uint32_t blk = tile >> log_tpb, b = ((tile & ((1u << log_tpb) - 1u)) << 5) + lane;
    const size_t M = (size_t)1 << a.log_M;
    const uint32_t* src = a.src + (size_t)blockIdx.y * a.src_stride + (COSET ? 0 : r * M);
    uint32_t* dst = a.dst + (size_t)blockIdx.y * a.dst_stride + r * M;
    const size_t base = ((size_t)blk << a.log_Msub) + b;
    const uint32_t rr = r + a.coset_base;

    for (uint32_t i = threadIdx.x; i < R / 2; i += C::THREADS) tws[i] = a.tw[i];
    if (COSET && threadIdx.x < 4) su[threadIdx.x] = sigma_pow_m(a, rr, (threadIdx.x + 1u) << a.log_S);
    const uint32_t row0 = warp * RW;
    uint32_t v[RW];
    {
        const uint32_t* p = src + base + ((size_t)row0 << a.log_S);
#pragma unroll
        for (int u = 0; u < RW; u++) v[u] = p[(size_t)u << a.log_S];
    }
    if (COSET) {
        __syncthreads();
        uint32_t t[4];
        t[0] = sigma_pow_m(a, rr, (uint32_t)(base + ((size_t)row0 << a.log_S)));
#pragma unroll
        for (int u = 1; u < 4; u++) t[u] = mont_mul(t[0], su[u - 1]);
        const uint32_t step = su[3];
#pragma unroll
        for (int u = 0; u < RW; u++) {
            v[u] = mont_mul(v[u], t[u & 3]);
            if (u + 4 < RW) t[u & 3] = mont_mul(t[u & 3], step);
        }
    }
#pragma unroll
    for (int u = 0; u < RW; u++) sm[(row0 + u) * TP + lane] = v[u];
    __syncthreads();
    cta_dif<LOG_R>(sm, tws, warp, lane);
    // twiddle w_Msub^(b c) and store (c, b)
    {
        const uint32_t N1 = (1u << a.hpow.log_N) - 1u, sh = a.hpow.log_N - a.log_Msub;
        uint32_t e1 = b << sh, e0 = (b * row0) << sh;
        if (a.inverse) { e1 = ((N1 + 1u) - e1) & N1; e0 = ((N1 + 1u) - e0) & N1; }
        const uint32_t w1 = pw(a.hpow, e1), w2 = mont_mul(w1, w1), w3 = mont_mul(w2, w1), w4 = mont_mul(w2, w2);
        uint32_t t[4];
        t[0] = pw(a.hpow, e0);
        t[1] = mont_mul(t[0], w1); t[2] = mont_mul(t[0], w2); t[3] = mont_mul(t[0], w3);
        uint32_t* q = dst + ((size_t)blk << a.log_Msub) + ((size_t)row0 << a.log_S) + b;
#pragma unroll
        for (int u = 0; u < RW; u++) {
            const uint32_t x = sm[brev_c<LOG_R>(row0 + u) * TP + lane];
            q[(size_t)u << a.log_S] = mont_mul(x, t[u & 3]);
            if (u + 4 < RW) t[u & 3] = mont_mul(t[u & 3], w4);
        }
    }
}

// Final pass.  Lane l of the tile owns one contiguous row of R elements; lanes are arranged so
// that for every output digit c the 32 lanes write 32 consecutive words:
//   lane = u * CS + r   (r = coset, u = index along the first-pass digit c1)
//   out[(rho_nat + (M/R) c) * CS + r],  rho_nat = c1 + R1 c2
template <int LOG_R, int LOG_CS, bool COSET>
__global__ void __launch_bounds__(NttCfg<LOG_R>::THREADS) ntt_final_kernel(const NttPassArgs a)
{
    pdl_enter();
    using C = NttCfg<LOG_R>;
    constexpr int R = C::R, RW = C::RW, CS = 1 << LOG_CS;
    extern __shared__ uint32_t smem[];
    uint32_t* sm = smem;
    uint32_t* tws = smem + R * TP;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t log_U = (5 - LOG_CS) < a.log_R1 ? (5 - LOG_CS) : a.log_R1;   // lanes along c1
    const uint32_t log_tiles_c1 = a.log_R1 - log_U;
    const uint32_t c2 = blockIdx.x >> log_tiles_c1, c1_0 = (blockIdx.x & ((1u << log_tiles_c1) - 1u)) << log_U;
    const size_t M = (size_t)1 << a.log_M;
    const uint32_t* srcb = a.src + (size_t)blockIdx.y * a.src_stride;
    uint32_t* dstb = a.dst + (size_t)blockIdx.y * a.dst_stride;

    for (uint32_t i = threadIdx.x; i < R / 2; i += C::THREADS) tws[i] = a.tw[i];
    // load: the tile is n_rows contiguous rows of R words; flat index f = l * R + a0, a warp reads 32
    // consecutive words of one row per access (R >= 32), all RW accesses of a thread in flight
    const uint32_t n_rows = CS << log_U, total = n_rows << LOG_R;
    uint32_t v[RW];
#pragma unroll
    for (int u = 0; u < RW; u++) {
        const uint32_t f = threadIdx.x + u * C::THREADS;
        if (f < total) {
            const uint32_t l = f >> LOG_R, a0 = f & (R - 1);
            const uint32_t rr = l & (CS - 1), uu = l >> LOG_CS;
            const size_t q = ((size_t)(c1_0 + uu) << a.log_Q2) + c2;
            v[u] = srcb[(COSET ? 0 : rr * M) + (q << LOG_R) + a0];
        }
    }
#pragma unroll
    for (int u = 0; u < RW; u++) {
        const uint32_t f = threadIdx.x + u * C::THREADS;
        if (f < total) {
            const uint32_t l = f >> LOG_R, a0 = f & (R - 1);
            if (COSET) {  // single-pass transforms only (M = R <= 1024): the plain table form is fine here
                const uint32_t rr = (l & (CS - 1)) + a.coset_base, uu = l >> LOG_CS;
                const size_t q = ((size_t)(c1_0 + uu) << a.log_Q2) + c2;
                v[u] = mont_mul(v[u], sigma_pow_m(a, rr, (uint32_t)((q << LOG_R) + a0)));
            }
            sm[a0 * TP + l] = v[u];
        }
    }
    __syncthreads();
    cta_dif<LOG_R>(sm, tws, warp, lane);
    if (lane < n_rows) {
        const uint32_t r = lane & (CS - 1), u = lane >> LOG_CS;
        const size_t rho = (size_t)(c1_0 + u) + ((size_t)c2 << a.log_R1);
        const uint32_t log_rows = a.log_M - LOG_R;
        uint32_t* q = dstb + (rho << LOG_CS) + r;
        const bool packed = LOG_CS == 0 && a.pack.log_jb != 0;
#pragma unroll
        for (int i = 0; i < RW; i++) {
            const uint32_t c = warp + i * C::NW;
            uint32_t x = sm[brev_c<LOG_R>(c) * TP + lane];
            if (a.use_scale) x = mont_mul(x, a.scale_m);
            if (packed) dstb[pack_index(a.pack, (uint32_t)(rho + ((size_t)c << log_rows)))] = x;   // 32 lanes: 32 consecutive j, one block
            else q[((size_t)c << log_rows) << LOG_CS] = x;
        }
    }
}

// ================================================================ v2: narrow tiles (8 columns), two passes up to 2^20
// The kernels above move tiles of 32 columns, so a radix-2^10 pass needs 135 KB of shared memory and one 1024-thread CTA
// per SM, and a 2^20 transform is cheaper as THREE passes of 2^7.  Here a tile is R rows x 8 columns (one 32-byte sector
// per row): 40 KB at R = 2^10, so a length-2^20 transform is TWO passes -- one global round trip and one twiddle
// multiplication per point less -- and the shared-memory traffic, which the instruction queue of the memory pipe bounds
// before the banks do, is cut to ~3 instructions per point and pass:
//   * a thread owns 16 points of one column in every round; the first round takes them straight from global memory and
//     the last one stores straight to global memory (a warp always touches 4 rows x 8 consecutive columns: full sectors);
//   * between rounds the tile lives in shared memory as THREAD-CONTIGUOUS chunks: the 16 points a thread needs next are
//     16 consecutive words (chunk pitch 20 words), read with four 128-bit loads; the producing round scatters its results
//     with 32-bit stores.  Every row index is (run-time bits of the thread) | (compile-time bits of the point), so a
//     scatter address is one run-time base plus an immediate.  The thread -> rows map of a round is a bit permutation
//     chosen so that the 4 row slots of a warp fall on 4 different words mod 4 of their target chunks: conflict-free;
//   * twiddles (and the per-row coset factors) are kept in shared memory in the same per-thread order: 128-bit broadcasts.
constexpr int N2_E = 16;        // points per thread
constexpr int N2_PITCH = 20;    // words per thread chunk (16 + 4: eight consecutive chunks start in eight different bank groups)

constexpr int n2_round_lr(int rem) { return rem > 4 ? (rem == 5 ? 3 : 4) : rem; }
struct N2Round { int S0, LR, LS, last; };
constexpr N2Round n2_round(int LOG_R, int k)
{
    int s0 = 0;
    for (int i = 0; i < k; i++) s0 += n2_round_lr(LOG_R - s0);
    const int lr = n2_round_lr(LOG_R - s0);
    return N2Round{s0, lr, LOG_R - s0 - lr, s0 + lr == LOG_R};
}
constexpr int n2_rounds(int LOG_R)
{
    int s0 = 0, k = 0;
    while (s0 < LOG_R) { s0 += n2_round_lr(LOG_R - s0); k++; }
    return k;
}
// Which row bit every bit of (row slot, trip) of a round stands for.  Digit bits [LS, LS+LR) belong to the butterflies.
//   last round : rows of a thread are consecutive: trip bits [LR, 4), slot bits [4, LOG_R)
//   otherwise  : slot bits 0,1 -> the two lowest bits of the NEXT round's position-in-chunk (its digit's low bits), so the
//                4 row slots of a warp scatter to 4 different words mod 4; trip bits -> the highest free bits; the other
//                slot bits -> the remaining free bits, ascending
struct N2Map { int rs[8]; int trip[4]; int n_rs, n_trip; };
constexpr N2Map n2_map(int LOG_R, int k)
{
    const N2Round r = n2_round(LOG_R, k);
    N2Map m{};
    m.n_rs = LOG_R - 4;
    m.n_trip = 4 - r.LR;
    bool used[12] = {};
    for (int b = r.LS; b < r.LS + r.LR; b++) used[b] = true;
    if (r.last) {
        for (int b = 0; b < m.n_trip; b++) { m.trip[b] = r.LR + b; used[r.LR + b] = true; }
        for (int b = 0; b < m.n_rs; b++) m.rs[b] = 4 + b;
        return m;
    }
    const N2Round nx = n2_round(LOG_R, k + 1);
    const int lo2 = nx.last ? 0 : nx.LS;          // low bits of the next round's position-in-chunk
    int nrs = 0;
    for (int b = 0; b < 2 && nrs < m.n_rs; b++) { m.rs[nrs++] = lo2 + b; used[lo2 + b] = true; }
    for (int b = 0, top = LOG_R - 1; b < m.n_trip; top--) {
        if (!used[top]) { m.trip[m.n_trip - 1 - b] = top; used[top] = true; b++; }
    }
    for (int b = 0; b < LOG_R && nrs < m.n_rs; b++)
        if (!used[b]) { m.rs[nrs++] = b; used[b] = true; }
    return m;
}
// rows of (slot, trip) without the digit
template <int LOG_R, int K>
__host__ __device__ constexpr uint32_t n2_rows(uint32_t rslot, uint32_t trip)
{
    constexpr N2Map m = n2_map(LOG_R, K);
    uint32_t row = 0;
    for (int b = 0; b < m.n_rs; b++) row |= ((rslot >> b) & 1u) << m.rs[b];
    for (int b = 0; b < m.n_trip; b++) row |= ((trip >> b) & 1u) << m.trip[b];
    return row;
}
// inverse: which thread chunk (slot) and which position in it a row has in round K
template <int LOG_R, int K>
__host__ __device__ constexpr uint32_t n2_slot_of(uint32_t row)
{
    constexpr N2Map m = n2_map(LOG_R, K);
    uint32_t s = 0;
    for (int b = 0; b < m.n_rs; b++) s |= ((row >> m.rs[b]) & 1u) << b;
    return s;
}
template <int LOG_R, int K>
__host__ __device__ constexpr uint32_t n2_pos_of(uint32_t row)
{
    constexpr N2Map m = n2_map(LOG_R, K);
    constexpr N2Round r = n2_round(LOG_R, K);
    uint32_t t = 0;
    for (int b = 0; b < m.n_trip; b++) t |= ((row >> m.trip[b]) & 1u) << b;
    return (t << r.LR) | ((row >> r.LS) & ((1u << r.LR) - 1u));
}
__host__ __device__ constexpr uint32_t brevc(uint32_t x, int bits)
{
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

template <int LOG_R>
struct Ntt2Cfg {
    static constexpr int R = 1 << LOG_R;
    static constexpr int THREADS = R * 8 / N2_E;       // 512 at R = 2^10 .. 32 at R = 2^6
    static constexpr int NSLOT = THREADS / 8;          // row slots: R / 16
    static constexpr int NR = n2_rounds(LOG_R);
    static constexpr int TILE_WORDS = NSLOT * 8 * N2_PITCH;
    // twiddle chunks: round k has 2^LS distinct `lo`, each with RAD - 1 twiddles stored in a chunk of RAD words
    static constexpr int tw_off(int k)
    {
        int o = 0;
        for (int i = 0; i < k; i++) { const N2Round r = n2_round(LOG_R, i); o += (1 << r.LS) << r.LR; }
        return o;
    }
    static constexpr int TW_WORDS = tw_off(NR);
    static constexpr int RF_WORDS = NSLOT * N2_PITCH;  // per-row coset factors in round-0 chunk order (strided pass with COSET)
};

// twiddle tables of all rounds in per-thread order: T_k[lo * RAD + slot(t, j)] = w_R^(((j << LS) + lo) << (S0 + t)).
// Built once per context for every radix and direction (ntt2_build_tables) behind the plain tables; a CTA copies its
// radix's table into shared memory.
template <int LOG_R>
__device__ __forceinline__ void n2_fill_twiddles(uint32_t* out, const uint32_t* __restrict__ tw_g, uint32_t tid, uint32_t nthreads)
{
    using C = Ntt2Cfg<LOG_R>;
#pragma unroll
    for (int k = 0; k < C::NR; k++) {
        const N2Round r = n2_round(LOG_R, k);
        const int rad = 1 << r.LR, total = (1 << r.LS) << r.LR;
        uint32_t* T = out + C::tw_off(k);
        for (int e = tid; e < total; e += nthreads) {
            const int lo = e >> r.LR, slot = e & (rad - 1);
            int t = 0, off = 0, h = rad >> 1;
            while (h && slot >= off + h) { off += h; h >>= 1; t++; }
            const int j = slot - off;
            T[e] = slot == rad - 1 ? 0u : tw_g[((j << r.LS) + lo) << (r.S0 + t)];
        }
    }
}
// offset (words) of radix 2^LOG_R's chunked table behind the 1024 words of plain tables
constexpr int NTT2_TAB_BASE = 1024;
__host__ __device__ constexpr int ntt2_tab_off(int log_r)
{
    int o = NTT2_TAB_BASE;
    if (log_r > 6) o += Ntt2Cfg<6>::TW_WORDS;
    if (log_r > 7) o += Ntt2Cfg<7>::TW_WORDS;
    if (log_r > 8) o += Ntt2Cfg<8>::TW_WORDS;
    if (log_r > 9) o += Ntt2Cfg<9>::TW_WORDS;
    if (log_r > 10) o += Ntt2Cfg<10>::TW_WORDS;
    return o;
}
__global__ void ntt2_tables_kernel(uint32_t* tw)   // tw: [plain 1023 | pad | chunked tables], built in place
{
    const uint32_t t = threadIdx.x, n = blockDim.x;
    n2_fill_twiddles<6>(tw + ntt2_tab_off(6), tw + ((1 << 5) - 1), t, n);
    n2_fill_twiddles<7>(tw + ntt2_tab_off(7), tw + ((1 << 6) - 1), t, n);
    n2_fill_twiddles<8>(tw + ntt2_tab_off(8), tw + ((1 << 7) - 1), t, n);
    n2_fill_twiddles<9>(tw + ntt2_tab_off(9), tw + ((1 << 8) - 1), t, n);
    n2_fill_twiddles<10>(tw + ntt2_tab_off(10), tw + ((1 << 9) - 1), t, n);
}
size_t ntt_table_words() { return (size_t)ntt2_tab_off(11); }
cudaError_t ntt2_build_tables(cudaStream_t s, uint32_t* tw_fwd, uint32_t* tw_inv)
{
    ntt2_tables_kernel<<<1, 256, 0, s>>>(tw_fwd);
    ntt2_tables_kernel<<<1, 256, 0, s>>>(tw_inv);
    return cudaGetLastError();
}
template <int LOG_R>
__device__ __forceinline__ void n2_load_twiddles(uint32_t* tw_s, const uint32_t* __restrict__ tab)
{
    using C = Ntt2Cfg<LOG_R>;
    const uint4* src = reinterpret_cast<const uint4*>(tab);
    uint4* dst = reinterpret_cast<uint4*>(tw_s);
    for (int e = threadIdx.x; e < C::TW_WORDS / 4; e += C::THREADS) dst[e] = src[e];
}

// the butterflies of one round on the 16 points of a thread (TRIPS mini-transforms of radix 2^LR); x[trip * RAD + d]
template <int LOG_R, int K>
__device__ __forceinline__ void n2_butterflies(uint32_t (&x)[N2_E], const uint32_t* tw_s, uint32_t row_rt)
{
    using C = Ntt2Cfg<LOG_R>;
    constexpr N2Round r = n2_round(LOG_R, K);
    constexpr int RAD = 1 << r.LR, TRIPS = N2_E / RAD;
    constexpr N2Map m = n2_map(LOG_R, K);
    // the trip bits never fall below the digit (they are the highest free bits, or bits [LR, 4) of the last round where
    // LS = 0), so every trip of a thread has the same `lo` and the same twiddles
    static_assert(m.n_trip == 0 || r.last || m.trip[0] >= r.LS + r.LR, "trip bits must lie above the digit");
    uint32_t tw[RAD];
    if (RAD >= 4) {
        const uint32_t lo = row_rt & ((1u << r.LS) - 1u);
        const uint4* tp = reinterpret_cast<const uint4*>(tw_s + C::tw_off(K) + (lo << r.LR));
#pragma unroll
        for (int q = 0; q < RAD / 4; q++) {
            const uint4 t4 = tp[q];
            tw[4 * q] = t4.x; tw[4 * q + 1] = t4.y; tw[4 * q + 2] = t4.z; tw[4 * q + 3] = t4.w;
        }
    } else {
        tw[0] = tw_s[C::tw_off(K)];
        tw[1] = 0;
    }
#pragma unroll
    for (int i = 0; i < TRIPS; i++) {
#pragma unroll
        for (int t = 0; t < r.LR; t++) {
            constexpr int dummy = 0; (void)dummy;
            const int half = RAD >> (t + 1), off = RAD - (RAD >> t);
#pragma unroll
            for (int j = 0; j < half; j++) {
                const bool trivial = (r.LS == 0 && j == 0);           // w^0
#pragma unroll
                for (int blk = 0; blk < RAD / (2 * half); blk++) {
                    const int i0 = i * RAD + blk * 2 * half + j, i1 = i0 + half;
                    const uint32_t u = x[i0], v = x[i1];
                    x[i0] = fadd(u, v);
                    x[i1] = trivial ? fsub(u, v) : mont_mul(fsub(u, v), tw[off + j]);
                }
            }
        }
    }
}

// scatter the 16 results of round K into the chunk order of round K + 1
template <int LOG_R, int K>
__device__ __forceinline__ void n2_scatter(uint32_t* tile, const uint32_t (&x)[N2_E], uint32_t row_rt, uint32_t col)
{
    constexpr N2Round r = n2_round(LOG_R, K);
    constexpr int RAD = 1 << r.LR;
    uint32_t* base = tile + (n2_slot_of<LOG_R, K + 1>(row_rt) * 8 + col) * N2_PITCH + n2_pos_of<LOG_R, K + 1>(row_rt);
#pragma unroll
    for (int u = 0; u < N2_E; u++) {
        const uint32_t rc = n2_rows<LOG_R, K>(0u, (uint32_t)(u / RAD)) | ((uint32_t)(u % RAD) << r.LS);   // compile-time row bits
        base[n2_slot_of<LOG_R, K + 1>(rc) * 8 * N2_PITCH + n2_pos_of<LOG_R, K + 1>(rc)] = x[u];
    }
}
__device__ __forceinline__ void n2_gather(const uint32_t* tile, uint32_t (&x)[N2_E], uint32_t rslot, uint32_t col)
{
    const uint4* p = reinterpret_cast<const uint4*>(tile + (rslot * 8 + col) * N2_PITCH);
#pragma unroll
    for (int q = 0; q < N2_E / 4; q++) {
        const uint4 v = p[q];
        x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
    }
}
// rounds 1 .. NR-1 of the CTA-local transform; x holds round 0's results on entry and the last round's on exit
template <int LOG_R, int K = 0>
__device__ __forceinline__ void n2_rounds_after_first(uint32_t* tile, const uint32_t* tw_s, uint32_t (&x)[N2_E], uint32_t rslot,
                                                      uint32_t col)
{
    if constexpr (K + 1 < Ntt2Cfg<LOG_R>::NR) {
        n2_scatter<LOG_R, K>(tile, x, n2_rows<LOG_R, K>(rslot, 0u), col);
        __syncthreads();
        n2_gather(tile, x, rslot, col);
        if constexpr (K + 2 < Ntt2Cfg<LOG_R>::NR) __syncthreads();     // the tile is overwritten by this round's scatter
        n2_butterflies<LOG_R, K + 1>(x, tw_s, n2_rows<LOG_R, K + 1>(rslot, 0u));
        n2_rounds_after_first<LOG_R, K + 1>(tile, tw_s, x, rslot, col);
    }
}

// Strided pass, narrow tile.  Same arithmetic as ntt_strided_kernel; the coset factor c_k sigma^k, k = a S + b, is split
// as sigma^(a S) (per row: a table in shared memory, one multiplication per point) times sigma^b (per column: folded
// into the start of the output twiddle chain, free), and so is the n^-1 of the inverse transform.
template <int LOG_R, bool COSET>
__global__ void __launch_bounds__(Ntt2Cfg<LOG_R>::THREADS) ntt2_strided_kernel(const NttPassArgs a)
{
    pdl_enter();
    using C = Ntt2Cfg<LOG_R>;
    constexpr int NR = C::NR;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* tile = smem;
    uint32_t* tw_s = smem + C::TILE_WORDS;
    uint32_t* rf = tw_s + C::TW_WORDS;                            // [NSLOT][N2_PITCH] (COSET)
    const uint32_t col = threadIdx.x & 7, rslot = threadIdx.x >> 3;
    const uint32_t log_tpb = a.log_S - 3;                        // tiles of 8 columns per block of M_sub
    const uint32_t log_per_array = a.log_M - LOG_R - 3;
    uint32_t tix = blockIdx.x;
    const uint32_t r = tix >> log_per_array;                     // coset
    tix &= (1u << log_per_array) - 1u;
    const uint32_t blk = tix >> log_tpb, b = ((tix & ((1u << log_tpb) - 1u)) << 3) + col;
    const size_t M = (size_t)1 << a.log_M;
    const uint32_t* src = a.src + (size_t)blockIdx.y * a.src_stride + (COSET ? 0 : r * M);
    uint32_t* dst = a.dst + (size_t)blockIdx.y * a.dst_stride + r * M;
    const size_t base = ((size_t)blk << a.log_Msub) + b;
    const uint32_t rr = r + a.coset_base;

    // round 0 straight from global memory: rows row0 | (u << LS0)
    constexpr N2Round r0 = n2_round(LOG_R, 0);
    static_assert(r0.LR == 4, "the first round is radix 16: one mini-transform per thread");
    const uint32_t row0 = n2_rows<LOG_R, 0>(rslot, 0u);
    uint32_t x[N2_E];
    {
        const uint32_t* p = src + base + ((size_t)row0 << a.log_S);
        const size_t step = (size_t)1 << (a.log_S + r0.LS);
#pragma unroll
        for (int u = 0; u < N2_E; u++, p += step) x[u] = *p;
    }
    n2_load_twiddles<LOG_R>(tw_s, a.tw2);
    // per-COLUMN constants of the output twiddle chain, computed once per column (by row slot 0) and shared: the step
    // w^(b 2^(LOG_R-4)) and its powers, and the column's share sigma^b of the coset factor (times the final scale)
    uint32_t* colc = rf + (COSET ? C::RF_WORDS : 0);            // [8][8]
    if (rslot == 0) {
        const uint32_t N1 = (1u << a.hpow.log_N) - 1u, sh = a.hpow.log_N - a.log_Msub;
        uint32_t e1 = ((b << (LOG_R - 4)) << sh) & N1;
        if (a.inverse) e1 = ((N1 + 1u) - e1) & N1;
        const uint32_t w1 = pw(a.hpow, e1), w2 = mont_mul(w1, w1);
        uint32_t f = R1;                                         // Montgomery form of 1
        if (COSET) f = sigma_pow_m(a, rr, b);
        if (a.use_scale) f = mont_mul(f, a.scale_m);
        uint32_t* cc = colc + col * 8;
        cc[0] = w1; cc[1] = w2; cc[2] = mont_mul(w2, w1); cc[3] = mont_mul(w2, w2); cc[4] = f;
    }
    if (COSET) {
        // first pass of the LDE (one block: blk == 0, k = a S + b): sigma^(a S) in the order round 0 reads it.  A row is
        // (slot bits) | (digit << LS0) with disjoint bits, so the factor is a product of a per-slot and a per-digit power:
        // NSLOT + 16 table look-ups per CTA, then one multiplication per entry.
        uint32_t* fa = colc + 64;                                 // [NSLOT] sigma^(rows(s) S)   [16] sigma^((u << LS0) S)
        for (uint32_t e = threadIdx.x; e < (uint32_t)C::NSLOT + 16; e += C::THREADS) {
            const uint32_t row = e < (uint32_t)C::NSLOT ? n2_rows<LOG_R, 0>(e, 0u) : ((e - C::NSLOT) << r0.LS);
            fa[e] = sigma_pow_m(a, rr, row << a.log_S);
        }
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < (uint32_t)C::NSLOT * N2_E; e += C::THREADS) {
            const uint32_t s = e >> 4, u = e & 15;
            rf[s * N2_PITCH + u] = mont_mul(fa[s], fa[C::NSLOT + u]);
        }
    }
    __syncthreads();
    if (COSET) {
        const uint4* rp = reinterpret_cast<const uint4*>(rf + rslot * N2_PITCH);
#pragma unroll
        for (int q = 0; q < N2_E / 4; q++) {
            const uint4 f = rp[q];
            x[4 * q] = mont_mul(x[4 * q], f.x); x[4 * q + 1] = mont_mul(x[4 * q + 1], f.y);
            x[4 * q + 2] = mont_mul(x[4 * q + 2], f.z); x[4 * q + 3] = mont_mul(x[4 * q + 3], f.w);
        }
    }
    n2_butterflies<LOG_R, 0>(x, tw_s, row0);
    n2_rounds_after_first<LOG_R>(tile, tw_s, x, rslot, col);
    // The thread now holds rows rslot * 16 + u of the bit-reversed result, i.e. outputs c = brev(rslot) + (brev4(u) << (LOG_R - 4)):
    // twiddle w_Msub^(b c) (times the column's share of the coset factor and the final scale), stores straight from registers.
    {
        constexpr N2Round rl = n2_round(LOG_R, NR - 1);
        static_assert(n2_rows<LOG_R, NR - 1>(1u, 0u) == 16u && rl.LS == 0, "last round: 16 consecutive rows per thread");
        const uint32_t c0 = brev_c<LOG_R>(rslot << 4);           // < 2^(LOG_R - 4)
        const uint32_t N1 = (1u << a.hpow.log_N) - 1u, sh = a.hpow.log_N - a.log_Msub;
        uint32_t e0 = ((b * c0) << sh) & N1;                     // start: w^(b c0)
        if (a.inverse) e0 = ((N1 + 1u) - e0) & N1;
        const uint32_t* cc = colc + col * 8;
        const uint32_t w1 = cc[0], w2 = cc[1], w3 = cc[2], w4 = cc[3];
        uint32_t t[4];
        t[0] = mont_mul(pw(a.hpow, e0), cc[4]);
        t[1] = mont_mul(t[0], w1); t[2] = mont_mul(t[0], w2); t[3] = mont_mul(t[0], w3);
        uint32_t* q = dst + ((size_t)blk << a.log_Msub) + ((size_t)c0 << a.log_S) + b;
        const size_t step = (size_t)1 << (a.log_S + LOG_R - 4);
        // outputs in the order m = brev4(u) = 0 .. 15 so that the four multiplicative chains advance in step
#pragma unroll
        for (int mm = 0; mm < N2_E; mm++, q += step) {
            const int u = (int)brevc((uint32_t)mm, 4);
            *q = mont_mul(x[u], t[mm & 3]);
            if (mm + 4 < N2_E) t[mm & 3] = mont_mul(t[mm & 3], w4);
        }
    }
}

// Final pass, narrow tile: the CTA takes 8 contiguous rows of R points -- the 8 cosets of one row index (LOG_CS = 3) or
// 8 consecutive row indices c1 of one array (LOG_CS = 0) -- as the 8 columns of its tile.  Outputs of one digit c are 8
// consecutive words of the natural-order result: the 8 lanes of a row slot store one sector.
// SINGLE (whole transform in this pass, M = R <= 2^10): an array is ONE row, so LOG_CS = 3 fills the 8 columns with the 8
// cosets of the LDE (scaled here by sigma_r^k) and LOG_CS = 0 uses one column of the tile -- still one short launch
// instead of a 32-column tile kernel that leaves 24 or 31 of its 32 lanes idle (2^10: 27 -> 6 us, 21 -> 5 us).
template <int LOG_R, int LOG_CS, bool SINGLE = false>
__global__ void __launch_bounds__(Ntt2Cfg<LOG_R>::THREADS) ntt2_final_kernel(const NttPassArgs a)
{
    pdl_enter();
    using C = Ntt2Cfg<LOG_R>;
    constexpr int CS = 1 << LOG_CS, LOG_U = SINGLE ? 0 : 3 - LOG_CS;
    extern __shared__ __align__(16) uint32_t smem[];
    uint32_t* tile = smem;
    uint32_t* tw_s = smem + C::TILE_WORDS;
    const uint32_t col = threadIdx.x & 7, rslot = threadIdx.x >> 3;
    const uint32_t log_tiles_c1 = a.log_R1 - LOG_U;
    const uint32_t c2 = blockIdx.x >> log_tiles_c1, c1_0 = (blockIdx.x & ((1u << log_tiles_c1) - 1u)) << LOG_U;
    const size_t M = (size_t)1 << a.log_M;
    const uint32_t* srcb = a.src + (size_t)blockIdx.y * a.src_stride;
    uint32_t* dstb = a.dst + (size_t)blockIdx.y * a.dst_stride;
    const bool live = !SINGLE || col < (uint32_t)CS;          // SINGLE with one array: only column 0 carries data

    constexpr N2Round r0 = n2_round(LOG_R, 0);
    const uint32_t row0 = n2_rows<LOG_R, 0>(rslot, 0u);
    uint32_t x[N2_E];
    if (live) {
        // column `col` of the tile is one contiguous row of R points; the other lanes of the CTA pick up the rest of each line
        const uint32_t rr = col & (CS - 1), uu = col >> LOG_CS;
        const size_t qrow = ((size_t)(c1_0 + uu) << a.log_Q2) + c2;
        // SINGLE LDE: the 8 cosets are transforms of the SAME coefficient array, scaled below by sigma_r^k (prover.rs:69-70)
        const uint32_t* p = srcb + ((SINGLE && a.use_coset) ? 0 : rr * M) + (qrow << LOG_R) + row0;
#pragma unroll
        for (int u = 0; u < N2_E; u++) x[u] = p[u << r0.LS];      // immediates
        if (SINGLE && a.use_coset) {
            uint32_t t = sigma_pow_m(a, rr + a.coset_base, row0);
            const uint32_t step = sigma_pow_m(a, rr + a.coset_base, 1u << r0.LS);
#pragma unroll
            for (int u = 0; u < N2_E; u++) {
                x[u] = mont_mul(x[u], t);
                if (u + 1 < N2_E) t = mont_mul(t, step);
            }
        }
    } else {
#pragma unroll
        for (int u = 0; u < N2_E; u++) x[u] = 0u;
    }
    n2_load_twiddles<LOG_R>(tw_s, a.tw2);
    __syncthreads();
    n2_butterflies<LOG_R, 0>(x, tw_s, row0);
    n2_rounds_after_first<LOG_R>(tile, tw_s, x, rslot, col);
    {
        const uint32_t log_rows = a.log_M - LOG_R;
        const size_t rho = (size_t)c1_0 + ((size_t)c2 << a.log_R1);      // first of the tile's row indices
        const uint32_t c0 = brev_c<LOG_R>(rslot << 4);
        // outputs c = c0 + (m << (LOG_R - 4)), m = brev4(u) = 0 .. 15: a constant stride in the natural-order result
        const size_t o0 = ((rho + ((size_t)c0 << log_rows)) << LOG_CS) + col;
        const size_t step = (size_t)1 << (LOG_R - 4 + log_rows + LOG_CS);
        if (live) {
            if (LOG_CS == 0 && a.pack.log_jb != 0) {
#pragma unroll
                for (int mm = 0; mm < N2_E; mm++) {
                    uint32_t v = x[brevc((uint32_t)mm, 4)];
                    if (a.use_scale) v = mont_mul(v, a.scale_m);
                    dstb[pack_index(a.pack, (uint32_t)(o0 + mm * step))] = v;
                }
            } else {
                uint32_t* q = dstb + o0;
#pragma unroll
                for (int mm = 0; mm < N2_E; mm++, q += step) {
                    uint32_t v = x[brevc((uint32_t)mm, 4)];
                    if (a.use_scale) v = mont_mul(v, a.scale_m);
                    *q = v;
                }
            }
        }
    }
}

// After the size-n inverse transform of (a_0..a_{n-2}, 0): c'_k.  The interpolant of the n-1
// trace points (prover.rs:60: zip truncates to 1023 pairs) has degree <= n-2, i.e. is
//   c_k = c'_k - t g^(k+1),  t = c'_{n-1}
// (equivalently: the transform of the trace extended by the phantom value -g sum a_i g^i).
__global__ void phantom_fix_kernel(uint32_t* c, uint64_t stride, uint32_t log_n, PowTab hpow)
{
    pdl_enter();
    const uint32_t n = 1u << log_n;
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t* cb = c + (size_t)blockIdx.y * stride;
    const uint32_t t = cb[n - 1];
    const uint32_t sh = hpow.log_N - log_n;  // g = h^(N/n)
    uint32_t e = ((k + 1u) << sh) & ((1u << hpow.log_N) - 1u);
    uint32_t v = fsub(cb[k], mont_mul(t, pw(hpow, e)));
    if (k != n - 1) cb[k] = v;  // c_{n-1} = t - t g^n = 0 is written by phantom_clear_kernel afterwards
}
__global__ void phantom_clear_kernel(uint32_t* c, uint64_t stride, uint32_t log_n)
{
    pdl_enter();
    c[(size_t)blockIdx.x * stride + (1u << log_n) - 1] = 0u;
}

// The phantom point up front (used when the caller provides an accumulator): the interpolant of the n-1 trace points
// is the size-n inverse transform of (a_0 .. a_{n-2}, y) with y chosen so that the top coefficient vanishes,
//   c_{n-1} = n^-1 sum_i y_i g^(-i(n-1)) = n^-1 sum_i y_i g^i = 0   =>   y = -g sum_{i<n-1} a_i g^i,
// so one reduction over the trace replaces the post-transform correction pass (read + write of all n coefficients).
// Partial sums are canonical residues < 2^32, at most 2^19 of them: their plain 64-bit sum cannot overflow, so CTAs
// combine with one integer atomicAdd and the last kernel reduces mod p.
constexpr int PH_CH = 8;  // trace values per thread
__global__ void __launch_bounds__(256) phantom_sum_kernel(const uint32_t* y, uint64_t stride, uint32_t log_n, PowTab hpow,
                                                          unsigned long long* acc)
{
    pdl_enter();
    __shared__ uint32_t red[8];
    const uint32_t n1 = (1u << log_n) - 1u;                    // points 0 .. n-2
    const uint32_t* yb = y + (size_t)blockIdx.y * stride;
    const uint32_t i0 = (blockIdx.x * 256 + threadIdx.x) * PH_CH;
    uint32_t sum = 0;
    if (i0 < n1) {
        const uint32_t sh = hpow.log_N - log_n;                // g = h^(N/n)
        const uint32_t g_m = pw(hpow, 1u << sh);
        uint32_t gp = pw(hpow, (i0 << sh) & ((1u << hpow.log_N) - 1u));
#pragma unroll
        for (int j = 0; j < PH_CH; j++) {
            if (i0 + j < n1) sum = fadd(sum, mont_mul(yb[i0 + j], gp));
            gp = mont_mul(gp, g_m);
        }
    }
    // block sum of canonical residues (64-bit plain sum of <= 256 values, reduced once)
    unsigned long long w = sum;
#pragma unroll
    for (int o = 16; o; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (uint32_t)(w % P);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int i = 0; i < 8; i++) t += red[i];
        atomicAdd(acc + blockIdx.y, t % P);
    }
}
__global__ void phantom_set_kernel(uint32_t* y, uint64_t stride, uint32_t log_n, PowTab hpow, unsigned long long* acc, uint32_t batch)
{
    pdl_enter();
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const uint32_t sh = hpow.log_N - log_n;
    const uint32_t g_m = pw(hpow, 1u << sh);
    const uint32_t S = (uint32_t)(acc[b] % P);
    acc[b] = 0;                                                // re-arm for the next call on this stream
    y[(size_t)b * stride + (1u << log_n) - 1u] = fneg(mont_mul(S, g_m));
}

// ---------------------------------------------------------------- host side
struct NttPlan {
    uint32_t n_pass;
    uint32_t log_r[3];
};
// Radix choice: CTA-local radix <= 2^10 (shared-memory tile [R][33] words), and small enough that a pass
// has at least ~2^NTT_LOG_TILES tiles of 32 columns so every SM holds several CTAs (one CTA's global
// load/store phase overlaps another's butterflies).  At most three passes (mu <= 30).
#ifndef NTT_LOG_TILES
#define NTT_LOG_TILES 10
#endif
static NttPlan make_plan(uint32_t mu, uint32_t log_arrays)
{
    NttPlan p{};
    static const int log_tiles = [] { const char* e = getenv("ZK_NTT_LOG_TILES"); return e ? atoi(e) : NTT_LOG_TILES; }();
    int cap = (int)(mu + log_arrays) - 5 - log_tiles;
    if (cap > 10) cap = 10;
    if (cap < 5) cap = 5;
    uint32_t np = mu <= 10 ? 1 : (mu + cap - 1) / cap;
    if (np > 3) np = 3;
    if (np > mu / 5) np = mu / 5;  // every pass needs radix >= 32: strides of >= 32 words (128 B rows)
    if (np < 1) np = 1;
    p.n_pass = np;
    for (uint32_t i = 0; i < np; i++) p.log_r[i] = (mu + np - 1 - i) / np;  // as even as possible, larger first
    return p;
}
// tw tables: for radix 2^m the R/2 entries start at offset 2^(m-1) - 1 (m >= 1)
static inline size_t tw_off(uint32_t m) { return m ? ((size_t)1 << (m - 1)) - 1 : 0; }

struct NttTables {
    const uint32_t* tw_fwd;
    const uint32_t* tw_inv;
    PowTab hpow, wpow;
};

template <typename K>
static cudaError_t launch_k(K kernel, dim3 grid, int threads, size_t smem, cudaStream_t s, const NttPassArgs& a)
{
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute((const void*)kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return zk_launch(kernel, grid, dim3(threads), smem, s, a);
}

template <int LOG_R>
static cudaError_t launch_strided(bool coset, dim3 grid, cudaStream_t s, const NttPassArgs& a)
{
    const size_t smem = ((size_t)NttCfg<LOG_R>::R * TP + NttCfg<LOG_R>::R / 2 + 4) * 4;
    if (coset) return launch_k(ntt_strided_kernel<LOG_R, true>, grid, NttCfg<LOG_R>::THREADS, smem, s, a);
    return launch_k(ntt_strided_kernel<LOG_R, false>, grid, NttCfg<LOG_R>::THREADS, smem, s, a);
}
template <int LOG_R>
static cudaError_t launch_final(bool coset, uint32_t n_cosets, dim3 grid, cudaStream_t s, const NttPassArgs& a)
{
    const size_t smem = ((size_t)NttCfg<LOG_R>::R * TP + NttCfg<LOG_R>::R / 2 + 1) * 4;
    const int th = NttCfg<LOG_R>::THREADS;
    if (n_cosets == 8) {
        if (coset) return launch_k(ntt_final_kernel<LOG_R, 3, true>, grid, th, smem, s, a);
        return launch_k(ntt_final_kernel<LOG_R, 3, false>, grid, th, smem, s, a);
    }
    if (coset) return launch_k(ntt_final_kernel<LOG_R, 0, true>, grid, th, smem, s, a);
    return launch_k(ntt_final_kernel<LOG_R, 0, false>, grid, th, smem, s, a);
}

template <int LOG_R>
static cudaError_t launch_strided2(bool coset, dim3 grid, cudaStream_t s, const NttPassArgs& a)
{
    using C = Ntt2Cfg<LOG_R>;
    const size_t smem = ((size_t)C::TILE_WORDS + C::TW_WORDS + (coset ? C::RF_WORDS + C::NSLOT + 16 : 0) + 64) * 4;
    if (coset) return launch_k(ntt2_strided_kernel<LOG_R, true>, grid, Ntt2Cfg<LOG_R>::THREADS, smem, s, a);
    return launch_k(ntt2_strided_kernel<LOG_R, false>, grid, Ntt2Cfg<LOG_R>::THREADS, smem, s, a);
}
template <int LOG_R>
static cudaError_t launch_final2(uint32_t n_cosets, dim3 grid, cudaStream_t s, const NttPassArgs& a)
{
    const size_t smem = ((size_t)Ntt2Cfg<LOG_R>::TILE_WORDS + Ntt2Cfg<LOG_R>::TW_WORDS) * 4;
    if (n_cosets == 8) return launch_k(ntt2_final_kernel<LOG_R, 3>, grid, Ntt2Cfg<LOG_R>::THREADS, smem, s, a);
    return launch_k(ntt2_final_kernel<LOG_R, 0>, grid, Ntt2Cfg<LOG_R>::THREADS, smem, s, a);
}
template <int LOG_R>
static cudaError_t launch_single2(uint32_t n_cosets, dim3 grid, cudaStream_t s, const NttPassArgs& a)
{
    const size_t smem = ((size_t)Ntt2Cfg<LOG_R>::TILE_WORDS + Ntt2Cfg<LOG_R>::TW_WORDS) * 4;
    if (n_cosets == 8) return launch_k(ntt2_final_kernel<LOG_R, 3, true>, grid, Ntt2Cfg<LOG_R>::THREADS, smem, s, a);
    return launch_k(ntt2_final_kernel<LOG_R, 0, true>, grid, Ntt2Cfg<LOG_R>::THREADS, smem, s, a);
}
// which kernel family: 2 (narrow tiles, default from 2^12) or 1; ZK_NTT_V=1 forces the wide-tile kernels (experiments)
static int ntt_version()
{
    static const int v = [] { const char* e = getenv("ZK_NTT_V"); return e ? atoi(e) : 2; }();
    return v;
}
// smallest transform that takes the narrow-tile kernels.  Measured on B200 (LDE, us, narrow vs wide tiles): 2^12 9.5 / 10.6,
// 2^14 10.5 / 12.7, 2^16 12.8 / 24.0, 2^18 24.8 / 30.0, 2^20 70.1 / 83.4, 2^22 267 / 316, 2^24 1180 / 1386: they win from the
// first size they can run (two passes of radix >= 2^6).  ZK_NTT2_MIN_MU overrides (>= 12).
static int ntt2_min_mu()
{
    static const int v = [] { const char* e = getenv("ZK_NTT2_MIN_MU"); int m = e ? atoi(e) : 12; return m < 12 ? 12 : m; }();
    return v;
}
// narrow-tile plan: ceil(mu / 10) passes of radix 2^6 .. 2^10, as even as possible, larger first
static NttPlan make_plan2(uint32_t mu)
{
    NttPlan p{};
    static const uint32_t maxr = [] { const char* e = getenv("ZK_NTT2_MAXR"); uint32_t v = e ? (uint32_t)atoi(e) : 10u; return v < 6 ? 6u : (v > 10 ? 10u : v); }();
    p.n_pass = (mu + maxr - 1) / maxr;
    if (p.n_pass > 3) p.n_pass = 3;
    for (uint32_t i = 0; i < p.n_pass; i++) p.log_r[i] = (mu + p.n_pass - 1 - i) / p.n_pass;
    return p;
}

// measurement support: ZK_NTT_TIMING=1 brackets every pass with CUDA events; tools read them back through ntt_pass_times()
static cudaEvent_t g_pass_ev[5];
static int g_pass_n = -1;   // -1: off
static bool pass_timing()
{
    static const bool on = [] {
        const char* e = getenv("ZK_NTT_TIMING");
        if (!(e && e[0] == '1')) return false;
        for (auto& ev : g_pass_ev) cudaEventCreate(&ev);
        return true;
    }();
    return on;
}
static void pass_mark(cudaStream_t s, int i)
{
    if (pass_timing() && i < 5) { cudaEventRecord(g_pass_ev[i], s); g_pass_n = i; }
}
int ntt_pass_times(float ms[4])
{
    if (!pass_timing() || g_pass_n < 1) return 0;
    cudaEventSynchronize(g_pass_ev[g_pass_n]);
    for (int i = 0; i < g_pass_n; i++) cudaEventElapsedTime(&ms[i], g_pass_ev[i], g_pass_ev[i + 1]);
    return g_pass_n;
}

// Transform `n_cosets` (1 or 8) arrays of length 2^mu.  Passes before the last run in place on `work`
// (n_cosets * M words per proof); the last writes natural order (interleaved over cosets) to dst.
// With coset_scale the first pass reads the single coefficient array `coeffs` instead of `work`.
static cudaError_t run_ntt(cudaStream_t s, const NttTables& T, uint32_t mu, uint32_t n_cosets, bool coset_scale,
                           bool inverse, const uint32_t* coeffs, uint64_t coeffs_stride, uint32_t* work,
                           uint64_t work_stride, uint32_t* dst, uint64_t dst_stride, bool use_scale,
                           uint32_t scale_m, uint32_t batch, int* n_launches, uint32_t coset_base = 0,
                           const NttPack* pack = nullptr)
{
    if (n_cosets != 1 && n_cosets != 8) return cudaErrorInvalidValue;
    uint32_t log_arrays = 0;
    while ((1u << log_arrays) < n_cosets * batch) log_arrays++;
    // a single array below 2^22 is too few narrow tiles to fill the machine (2^20 interpolation out of DRAM: 39 vs 28 us)
    const bool v2 = mu >= (uint32_t)ntt2_min_mu() && ntt_version() == 2 && (n_cosets * batch >= 8 || mu >= 22);
    NttPlan p = v2 ? make_plan2(mu) : make_plan(mu, log_arrays);
    NttPassArgs a{};
    if (mu >= 6 && mu <= 10 && ntt_version() == 2) {
        // the whole transform in one narrow-tile CTA per proof: arrays are single rows
        a.log_M = mu; a.inverse = inverse; a.hpow = T.hpow; a.wpow = T.wpow; a.coset_base = coset_base;
        if (pack) a.pack = *pack;
        a.tw = (inverse ? T.tw_inv : T.tw_fwd) + tw_off(mu);
        a.tw2 = (inverse ? T.tw_inv : T.tw_fwd) + ntt2_tab_off((int)mu);
        a.src = coeffs; a.src_stride = coeffs_stride; a.dst = dst; a.dst_stride = dst_stride;
        a.use_scale = use_scale; a.scale_m = scale_m; a.use_coset = coset_scale;
        a.log_Msub = mu;
        dim3 grid(1, batch);
        cudaError_t e1;
        switch (mu) {
            case 6: e1 = launch_single2<6>(n_cosets, grid, s, a); break;
            case 7: e1 = launch_single2<7>(n_cosets, grid, s, a); break;
            case 8: e1 = launch_single2<8>(n_cosets, grid, s, a); break;
            case 9: e1 = launch_single2<9>(n_cosets, grid, s, a); break;
            default: e1 = launch_single2<10>(n_cosets, grid, s, a); break;
        }
        if (e1 == cudaSuccess && n_launches) (*n_launches)++;
        return e1;
    }
    a.log_M = mu;
    a.inverse = inverse;
    a.hpow = T.hpow;
    a.wpow = T.wpow;
    a.coset_base = coset_base;
    if (pack) a.pack = *pack;
    const uint32_t* twb = inverse ? T.tw_inv : T.tw_fwd;
    uint32_t log_S = mu;
    cudaError_t e;
    pass_mark(s, 0);
    for (uint32_t i = 0; i + 1 < p.n_pass; i++) {
        const uint32_t log_R = p.log_r[i];
        a.log_Msub = log_S;
        log_S -= log_R;
        a.log_S = log_S;
        a.tw = twb + tw_off(log_R);
        a.tw2 = twb + ntt2_tab_off((int)log_R);
        const bool cs = (i == 0 && coset_scale);
        a.src = (i == 0) ? coeffs : work;
        a.src_stride = (i == 0) ? coeffs_stride : work_stride;
        a.dst = work;
        a.dst_stride = work_stride;
        if (v2) {
            a.use_scale = (i == 0 && use_scale);   // the final scale rides on the first pass's twiddle chain
            a.scale_m = scale_m;
            dim3 grid2(((1u << (mu - 3)) >> log_R) * n_cosets, batch);   // (M / (R*S)) * (S/8)
            switch (log_R) {
                case 6: e = launch_strided2<6>(cs, grid2, s, a); break;
                case 7: e = launch_strided2<7>(cs, grid2, s, a); break;
                case 8: e = launch_strided2<8>(cs, grid2, s, a); break;
                case 9: e = launch_strided2<9>(cs, grid2, s, a); break;
                case 10: e = launch_strided2<10>(cs, grid2, s, a); break;
                default: return cudaErrorInvalidValue;
            }
            if (e != cudaSuccess) return e;
            if (n_launches) (*n_launches)++;
            pass_mark(s, i + 1);
            continue;
        }
        uint32_t tiles = (1u << (mu - 5)) >> log_R;  // (M / (R*S)) * (S/32)
        dim3 grid(tiles * n_cosets, batch);
        switch (log_R) {
            case 5: e = launch_strided<5>(cs, grid, s, a); break;
            case 6: e = launch_strided<6>(cs, grid, s, a); break;
            case 7: e = launch_strided<7>(cs, grid, s, a); break;
            case 8: e = launch_strided<8>(cs, grid, s, a); break;
            case 9: e = launch_strided<9>(cs, grid, s, a); break;
            case 10: e = launch_strided<10>(cs, grid, s, a); break;
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    const uint32_t log_R = p.log_r[p.n_pass - 1];
    a.log_Msub = log_R;
    a.log_S = 0;
    a.tw = twb + tw_off(log_R);
    a.tw2 = twb + ntt2_tab_off((int)log_R);
    const bool cs = (p.n_pass == 1 && coset_scale);
    a.src = (p.n_pass == 1) ? coeffs : work;
    a.src_stride = (p.n_pass == 1) ? coeffs_stride : work_stride;
    a.dst = dst;
    a.dst_stride = dst_stride;
    a.use_scale = use_scale;
    a.scale_m = scale_m;
    a.log_R1 = p.n_pass >= 2 ? p.log_r[0] : 0;
    a.log_Q2 = p.n_pass == 3 ? p.log_r[1] : 0;
    if (v2) {
        a.use_scale = 0;   // applied by the first pass
        const uint32_t log_U = n_cosets == 8 ? 0 : 3;
        dim3 grid(1u << (a.log_R1 - log_U + a.log_Q2), batch);
        switch (log_R) {
            case 6: e = launch_final2<6>(n_cosets, grid, s, a); break;
            case 7: e = launch_final2<7>(n_cosets, grid, s, a); break;
            case 8: e = launch_final2<8>(n_cosets, grid, s, a); break;
            case 9: e = launch_final2<9>(n_cosets, grid, s, a); break;
            case 10: e = launch_final2<10>(n_cosets, grid, s, a); break;
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
        pass_mark(s, p.n_pass);
        return cudaSuccess;
    }
    {
        uint32_t log_U = n_cosets == 8 ? 2 : 5;
        if (log_U > a.log_R1) log_U = a.log_R1;
        uint32_t tiles = 1u << (a.log_R1 - log_U + a.log_Q2);
        dim3 grid(tiles, batch);
        switch (log_R) {
            case 3: e = launch_final<3>(cs, n_cosets, grid, s, a); break;
            case 4: e = launch_final<4>(cs, n_cosets, grid, s, a); break;
            case 5: e = launch_final<5>(cs, n_cosets, grid, s, a); break;
            case 6: e = launch_final<6>(cs, n_cosets, grid, s, a); break;
            case 7: e = launch_final<7>(cs, n_cosets, grid, s, a); break;
            case 8: e = launch_final<8>(cs, n_cosets, grid, s, a); break;
            case 9: e = launch_final<9>(cs, n_cosets, grid, s, a); break;
            case 10: e = launch_final<10>(cs, n_cosets, grid, s, a); break;
            default: return cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    return cudaSuccess;
}

// lagrange() over the subgroup (polynomial.rs:337; prover.rs:60-61).  y: n words per proof with
// y[n-1] ignored when n_points == n-1.  coeffs may alias y (in place).
cudaError_t launch_interpolate(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow,
                               PowTab wpow, uint32_t log_n, bool phantom, uint32_t* y, uint64_t y_stride,
                               uint32_t* coeffs, uint64_t coeffs_stride, uint32_t batch, int* n_launches,
                               unsigned long long* acc)
{
    NttTables T{tw_fwd, tw_inv, hpow, wpow};
    cudaError_t e;
    const uint32_t n = 1u << log_n;
    if (phantom && acc && log_n >= 22) {
        // phantom value first, then a plain inverse transform (acc: one zeroed 64-bit word per proof).  Pays from 2^22
        // (57 vs 71 us; 193 vs 224 us at 2^24); below that the two small extra launches cost more than the correction pass
        // over an L2-resident array (2^20: 35 vs 29 us).
        const uint32_t per_cta = 256 * PH_CH;
        dim3 grid((n - 1 + per_cta - 1) / per_cta, batch);
        if ((e = zk_launch(phantom_sum_kernel, grid, dim3(256), 0, s, (const uint32_t*)y, y_stride, log_n, hpow, acc)) != cudaSuccess) return e;
        if ((e = zk_launch(phantom_set_kernel, dim3((batch + 63) / 64), dim3(64), 0, s, y, y_stride, log_n, hpow, acc, batch)) != cudaSuccess) return e;
        if (n_launches) (*n_launches) += 2;
        uint32_t ninv_m = h_to_mont(h_inv(n % P));
        return run_ntt(s, T, log_n, 1, false, true, y, y_stride, y, y_stride, coeffs, coeffs_stride, true, ninv_m, batch, n_launches);
    }
    if (phantom) {
        if ((e = zk_launch(phantom_clear_kernel, dim3(batch), dim3(1), 0, s, y, y_stride, log_n)) != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    uint32_t ninv_m = h_to_mont(h_inv(n % P));
    e = run_ntt(s, T, log_n, 1, false, true, y, y_stride, y, y_stride, coeffs, coeffs_stride, true, ninv_m,
                batch, n_launches);
    if (e != cudaSuccess) return e;
    if (phantom) {
        uint32_t th = n < 256 ? n : 256;
        dim3 grid(n / th, batch);
        // one CTA when n <= 256; for larger n the value t = c'[n-1] is only rewritten by the last CTA,
        // which is why the kernel never stores k = n-1 and a second tiny launch zeroes it.
        if ((e = zk_launch(phantom_fix_kernel, grid, dim3(th), 0, s, coeffs, coeffs_stride, log_n, hpow)) != cudaSuccess) return e;
        if ((e = zk_launch(phantom_clear_kernel, dim3(batch), dim3(1), 0, s, coeffs, coeffs_stride, log_n)) != cudaSuccess) return e;
        if (n_launches) (*n_launches) += 2;
    }
    return cudaSuccess;
}

// f_poly.solve over the coset 5 h^i (prover.rs:69-70).  work: 8 n words per proof.
cudaError_t launch_lde(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow, PowTab wpow,
                       uint32_t log_n, const uint32_t* coeffs, uint64_t coeffs_stride, uint32_t* work,
                       uint64_t work_stride, uint32_t* evals, uint64_t evals_stride, uint32_t batch,
                       int* n_launches)
{
    NttTables T{tw_fwd, tw_inv, hpow, wpow};
    return run_ntt(s, T, log_n, 1u << LOG_BLOWUP, true, false, coeffs, coeffs_stride, work, work_stride, evals,
                   evals_stride, false, 0, batch, n_launches);
}

// One coset of the LDE on its own: out[j] = f(5 h^(8 j + r)), j < n (sharded LDE: ranks split the cosets).
cudaError_t launch_lde_one_coset(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow, PowTab wpow,
                                 uint32_t log_n, uint32_t r, const uint32_t* coeffs, uint32_t* work, uint32_t* out,
                                 int* n_launches, const NttPack* pack)
{
    NttTables T{tw_fwd, tw_inv, hpow, wpow};
    return run_ntt(s, T, log_n, 1, true, false, coeffs, 0, work, 0, out, 0, false, 0, 1, n_launches, r, pack);
}

}  // namespace zk
