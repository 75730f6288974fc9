// K3 + K4 + K5 -- bottom-up SHA-256 Merkle trees with fused leaf sources.
//
// Reproduces Merkle::new (merkle.rs:14-51): implicit heap, root at 0, children of i at
// 2i+1 / 2i+2, leaves at [size-1, 2size-1); leaf = SHA256(BE32(v)), node = SHA256(L||R).
//
// Work decomposition (B200: 148 SMs, hashing is INT-pipe bound, ~3 compressions / leaf):
//   * pass kernel: ONE THREAD owns 2^k consecutive inputs and hashes the whole 2^k-subtree
//     above them serially (binary-counter merge, left siblings parked in a small per-thread
//     stack).  No inter-thread communication, every lane busy on every level; k adapts so
//     that the grid still fills the machine.  Inputs are leaf values (hashed lane-per-leaf)
//     or digests of a previous pass.
//   * top kernel: one CTA finishes the last <= 2^TOP_LOG inputs level by level in shared
//     memory and then runs the Fiat-Shamir step (channel.rs:19-32) on the root, so the
//     next stage's challenge never leaves the device.
// Leaf values can be produced in-kernel (north_star items 3/4): the FRI fold of the previous
// layer (polynomial.rs:385-400 in evaluation form, proof.rs:110-112) or the composition
// polynomial (prover.rs:101-173); the produced layer is written once and hashed from L1.
#include <cstdlib>
#include <mutex>
#include "common.cuh"
#define ZK_SHA_DEVICE
#include "sha256.cuh"

namespace zk {

// ---------------------------------------------------------------- helpers
// cudaFuncSetAttribute once per (kernel instantiation, device); contexts of several threads may race here
static std::mutex g_attr_mutex;
template <typename K>
static cudaError_t set_max_smem_once(bool (&done)[64], K kernel, int bytes)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    std::lock_guard<std::mutex> lk(g_attr_mutex);
    if (done[dev]) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess) done[dev] = true;
    return e;
}
__device__ __forceinline__ uint32_t pow_h_m(const PowTab& t, uint32_t e)
{
    return mont_mul(__ldg(t.lo + (e & ((1u << t.lo_bits) - 1u))), __ldg(t.hi + (e >> t.lo_bits)));
}

// ---------------------------------------------------------------- fused sources
// FRI fold, evaluation form.  next[i] = 2^-1 * [ (u+v) + beta * x_i^-1 * (u-v) ],
//   u = cur[i], v = cur[i+half], x_i = (5 h^i)^(2^k)  =>  x_i^-1 = 5^-(2^k) * h^(-i 2^k).
struct FoldCoef { uint32_t cb_m, step_m; };   // beta * 5^-(2^k) and h^-(2^k), Montgomery form: per layer, not per point
__device__ __forceinline__ FoldCoef fold_coef(const FoldArgs& f, const PowTab& pw, uint32_t beta_raw)
{
    const uint32_t N = 1u << pw.log_N;
    FoldCoef k;
    k.cb_m = mont_mul(to_mont(reduce_u32(beta_raw)), f.winv2k_m);            // field.rs:20-24
    k.step_m = pow_h_m(pw, (N - (1u << f.layer_k)) & (N - 1));
    return k;
}
template <int C>
__device__ __forceinline__ void fold_chunk(const FoldArgs& f, const PowTab& pw, const uint32_t* cur,
                                           uint32_t* next, uint32_t base, const FoldCoef& k)
{
    const uint32_t half = 1u << f.log_half;
    const uint32_t N = 1u << pw.log_N;
    const uint32_t step_m = k.step_m;
    const uint32_t gbase = shard_global(f.shard, base);                   // chunks never straddle a shard block
    uint32_t e = (N - ((gbase << f.layer_k) & (N - 1))) & (N - 1);
    uint32_t t_m = mont_mul(k.cb_m, pow_h_m(pw, e));                      // beta * x_base^-1
    uint32_t u[C], v[C];
    if (C >= 4) {
#pragma unroll
        for (int j = 0; j < C; j += 4) {
            uint4 a = *reinterpret_cast<const uint4*>(cur + base + j);
            uint4 b = *reinterpret_cast<const uint4*>(cur + base + half + j);
            u[j] = a.x; u[j + 1] = a.y; u[j + 2] = a.z; u[j + 3] = a.w;
            v[j] = b.x; v[j + 1] = b.y; v[j + 2] = b.z; v[j + 3] = b.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < C; j++) { u[j] = cur[base + j]; v[j] = cur[base + half + j]; }
    }
    uint32_t o[C];
#pragma unroll
    for (int j = 0; j < C; j++) {
        uint32_t s = fadd(u[j], v[j]);
        uint32_t d = mont_mul(fsub(u[j], v[j]), t_m);
        o[j] = mont_mul(fadd(s, d), f.inv2_m);
        t_m = mont_mul(t_m, step_m);
    }
    if (C >= 4) {
#pragma unroll
        for (int j = 0; j < C; j += 4)
            *reinterpret_cast<uint4*>(next + base + j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < C; j++) next[base + j] = o[j];
    }
}

// Composition polynomial at x = 5 h^i (prover.rs:101-173):
//  cp = a0 (f(x)-a_first)/(x-1) + a1 (f(x)-a_last)/(x-g^(n-2))
//     + a2 (f(g^2 x) - f(g x)^2 - f(x)^2) (x-g^(n-3))(x-g^(n-2))(x-g^(n-1)) / (x^n-1)
// f(g x) = f_eval[i+8], f(g^2 x) = f_eval[i+16] (indices mod N: cp_poly is evaluated on the
// whole coset, prover.rs:172-173).  x^n - 1 depends on i mod 8 only.  The two linear denominators are merged,
//  a0 (f-a_first)/(x-1) + a1 (f-a_last)/(x-g2) = [a0 (f-a_first)(x-g2) + a1 (f-a_last)(x-1)] / D,  D = (x-1)(x-g2),
// and D is inverted with one batched inversion per C points.
//
// Montgomery bookkeeping: mont_mul(u, v) = u v / R, so a value is carried as (true value) * R^e and every product
// lowers e by one; nothing is converted, the exponents are arranged to cancel:
//   x, f, a0, a1: e = 0;  D: e = -1;  prefix products: e = 1 - 2j;  1/D_j out of the backward pass: e = 3 for all j;
//   [..]: e = -2, times 1/D: e = 0.   squares: e = -1, numerator: e = -1, cubic: e = -2, product: e = -4,
//   times Z[r] = a2 / (x^n - 1) * R^5: e = 0.
// 17 multiplications per point + 43 / C for the inversion (was 33 + 63 / C with per-value conversions).
__device__ __forceinline__ uint32_t mont_inv_chain(uint32_t x)   // x^(p-2); p - 2 = 0b10 followed by thirty 1s
{
    const uint32_t x2 = mont_mul(x, x);
    const uint32_t x3 = mont_mul(x2, x);
    uint32_t t = mont_mul(x3, x3);
    t = mont_mul(t, t);
    const uint32_t x15 = mont_mul(t, x3);
    const uint32_t x31 = mont_mul(mont_mul(x15, x15), x);
    uint32_t r = x2;
#pragma unroll 1
    for (int g = 0; g < 6; g++) {
#pragma unroll
        for (int q = 0; q < 5; q++) r = mont_mul(r, r);
        r = mont_mul(r, x31);
    }
    return r;
}

// transition numerator over R at one point: (f(g^2 x) - step(f(x), f(g x))) / R, values canonical (see AirCoef)
__device__ __forceinline__ uint32_t air_num(const AirCoef& A, uint32_t fx, uint32_t fg, uint32_t fgg)
{
    if (!A.generic) return fsub(fsub(mont_mul(fgg, 1u), mont_mul(fg, fg)), mont_mul(fx, fx));   // prover.rs:134-138
    uint32_t r = fadd(mont_mul(mont_mul(fg, fg), A.c20m), mont_mul(mont_mul(fx, fx), A.c02m));
    r = fadd(r, mont_mul(mont_mul(fx, fg), A.c11m));
    r = fadd(r, fadd(mont_mul(fg, A.c10), mont_mul(fx, A.c01)));
    r = fadd(r, A.c00r);
    return fsub(mont_mul(fgg, 1u), r);
}

struct CpCoef {
    uint32_t a0, a1;       // alpha_0, alpha_1 reduced (field.rs:20-24), e = 0
    uint32_t z[8];         // alpha_2 / (x^n - 1) for i mod 8, e = 5
    uint32_t a_first, a_last;
};
__device__ __forceinline__ void cp_coef(const CpArgs& c, const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last,
                                        CpCoef& k)
{
    k.a0 = reduce_u32(alpha_raw[0]);
    k.a1 = reduce_u32(alpha_raw[1]);
    const uint32_t a2 = reduce_u32(alpha_raw[2]);
#pragma unroll
    for (int r = 0; r < 8; r++) k.z[r] = mont_mul(c.zinv_r6[r], a2);
    k.a_first = a_first;
    k.a_last = a_last;
}

template <int C>
__device__ __forceinline__ void cp_chunk(const CpArgs& c, const PowTab& pw, const uint32_t* f,
                                         uint32_t* out, uint32_t base, const CpCoef& k)
{
    const uint32_t B = 1u << LOG_BLOWUP;
    const uint32_t len_mask = (1u << c.log_len) - 1u;
    const bool sharded = c.shard.n_ranks > 1 || c.halo != nullptr;
    const uint32_t blk_mask = (1u << c.shard.log_block) - 1u;
    const uint32_t h_m = pow_h_m(pw, 1);
    uint32_t x = mont_mul(GEN, pow_h_m(pw, shard_global(c.shard, base)));  // canonical 5 h^i, i global
    uint32_t xs[C], D[C], pre[C];
    uint32_t acc = R1;
#pragma unroll
    for (int j = 0; j < C; j++) {
        xs[j] = x;
        D[j] = mont_mul(fsub(x, 1u), fsub(x, c.g2));
        pre[j] = acc;
        acc = mont_mul(acc, D[j]);
        x = mont_mul(x, h_m);
    }
    acc = mont_inv_chain(acc);  // no denominator vanishes on the coset
#pragma unroll
    for (int j = C - 1; j >= 0; j--) {
        const uint32_t i = base + j;
        const uint32_t inv_d = mont_mul(acc, pre[j]);        // 1 / D_j, e = 3
        if (j) acc = mont_mul(acc, D[j]);
        uint32_t fx = f[i], fgx, fggx;
        if (!sharded) {
            fgx = f[(i + B) & len_mask];
            fggx = f[(i + 2 * B) & len_mask];
        } else {
            // inside the local block, else in the 16-value head of the block that follows it globally
            const uint32_t t = i & blk_mask, bl = ((i >> c.shard.log_block) + c.halo_roll) & c.halo_mask;
            fgx = (t + B <= blk_mask) ? f[i + B] : c.halo[bl * 16 + (t + B - blk_mask - 1)];
            fggx = (t + 2 * B <= blk_mask) ? f[i + 2 * B] : c.halo[bl * 16 + (t + 2 * B - blk_mask - 1)];
        }
        const uint32_t d0 = fsub(xs[j], 1u), d1 = fsub(xs[j], c.g2);
        const uint32_t s = fadd(mont_mul(mont_mul(fsub(fx, k.a_first), d1), k.a0),
                                mont_mul(mont_mul(fsub(fx, k.a_last), d0), k.a1));
        const uint32_t p01 = mont_mul(s, inv_d);
        const uint32_t num = air_num(c.air, fx, fgx, fggx);
        const uint32_t cub = mont_mul(mont_mul(fsub(xs[j], c.g3), d1), fsub(xs[j], c.g1));
        const uint32_t p2 = mont_mul(mont_mul(num, cub), k.z[shard_global(c.shard, i) & 7]);
        out[i] = fadd(p01, p2);
    }
}

// The same polynomial from the per-size constant tables (CpArgs::tab): everything that depends on the point only
// is read back instead of recomputed -- 9 multiplications per point, no inversion, half the registers.
//   cp = a0 (f(x)-a_first) A_i + a1 (f(x)-a_last) B_i + a2 (f(g^2 x) - f(g x)^2 - f(x)^2) C_i
struct CpCoefTab { uint32_t a0m, a1m, a2m, a_first, a_last; };
__device__ __forceinline__ void cp_coef_tab(const uint32_t alpha_raw[3], uint32_t a_first, uint32_t a_last, CpCoefTab& k)
{
    k.a0m = to_mont(reduce_u32(alpha_raw[0]));                          // field.rs:20-24
    k.a1m = to_mont(reduce_u32(alpha_raw[1]));
    k.a2m = to_mont(reduce_u32(alpha_raw[2]));
    k.a_first = a_first;
    k.a_last = a_last;
}
// f at local index i + off (off = 8 or 16): inside the local block, else in the 16-value head of the block that follows it
// globally (CpArgs::halo), else -- not sharded -- modulo the domain
__device__ __forceinline__ uint32_t cp_neighbor(const CpArgs& c, const uint32_t* f, uint32_t i, uint32_t off)
{
    if (c.halo == nullptr) return f[(i + off) & ((1u << c.log_len) - 1u)];
    const uint32_t blk_mask = (1u << c.shard.log_block) - 1u, t = i & blk_mask;
    if (t + off <= blk_mask) return f[i + off];
    const uint32_t bl = ((i >> c.shard.log_block) + c.halo_roll) & c.halo_mask;
    return c.halo[bl * 16 + (t + off - blk_mask - 1)];
}
// the same for four consecutive points starting at a multiple of four: the four neighbours are in one place
__device__ __forceinline__ uint4 cp_neighbor4(const CpArgs& c, const uint32_t* f, uint32_t i, uint32_t off)
{
    if (c.halo == nullptr) return *reinterpret_cast<const uint4*>(f + ((i + off) & ((1u << c.log_len) - 1u)));
    const uint32_t blk_mask = (1u << c.shard.log_block) - 1u, t = i & blk_mask;
    if (t + off <= blk_mask) return *reinterpret_cast<const uint4*>(f + i + off);
    const uint32_t bl = ((i >> c.shard.log_block) + c.halo_roll) & c.halo_mask;
    return *reinterpret_cast<const uint4*>(c.halo + bl * 16 + (t + off - blk_mask - 1));
}
template <int C>
__device__ __forceinline__ void cp_chunk_tab(const CpArgs& c, const uint32_t* f, uint32_t* out, uint32_t base,
                                             const CpCoefTab& k)
{
    const uint32_t B = 1u << LOG_BLOWUP;
    const uint32_t N = 1u << c.log_len;      // local points: the tables are in local order
    uint32_t fx[C], fg[C], fgg[C], ta[C], tb[C], tc[C];
    if (C >= 4) {
#pragma unroll
        for (int j = 0; j < C; j += 4) {
            const uint4 v = *reinterpret_cast<const uint4*>(f + base + j);
            const uint4 a = *reinterpret_cast<const uint4*>(c.tab + base + j);
            const uint4 b = *reinterpret_cast<const uint4*>(c.tab + N + base + j);
            const uint4 q = *reinterpret_cast<const uint4*>(c.tab + 2 * (size_t)N + base + j);
            const uint4 v1 = cp_neighbor4(c, f, base + j, B);
            const uint4 v2 = cp_neighbor4(c, f, base + j, 2 * B);
            fx[j] = v.x; fx[j + 1] = v.y; fx[j + 2] = v.z; fx[j + 3] = v.w;
            ta[j] = a.x; ta[j + 1] = a.y; ta[j + 2] = a.z; ta[j + 3] = a.w;
            tb[j] = b.x; tb[j + 1] = b.y; tb[j + 2] = b.z; tb[j + 3] = b.w;
            tc[j] = q.x; tc[j + 1] = q.y; tc[j + 2] = q.z; tc[j + 3] = q.w;
            fg[j] = v1.x; fg[j + 1] = v1.y; fg[j + 2] = v1.z; fg[j + 3] = v1.w;
            fgg[j] = v2.x; fgg[j + 1] = v2.y; fgg[j + 2] = v2.z; fgg[j + 3] = v2.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < C; j++) {
            const uint32_t i = base + j;
            fx[j] = f[i]; fg[j] = cp_neighbor(c, f, i, B); fgg[j] = cp_neighbor(c, f, i, 2 * B);
            ta[j] = c.tab[i]; tb[j] = c.tab[N + i]; tc[j] = c.tab[2 * (size_t)N + i];
        }
    }
    uint32_t o[C];
#pragma unroll
    for (int j = 0; j < C; j++) {
        const uint32_t p0 = mont_mul(mont_mul(fsub(fx[j], k.a_first), ta[j]), k.a0m);
        const uint32_t p1 = mont_mul(mont_mul(fsub(fx[j], k.a_last), tb[j]), k.a1m);
        const uint32_t num = air_num(c.air, fx[j], fg[j], fgg[j]);   // / R
        const uint32_t p2 = mont_mul(mont_mul(num, tc[j]), k.a2m);
        o[j] = fadd(fadd(p0, p1), p2);
    }
    if (C >= 4) {
#pragma unroll
        for (int j = 0; j < C; j += 4)
            *reinterpret_cast<uint4*>(out + base + j) = make_uint4(o[j], o[j + 1], o[j + 2], o[j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < C; j++) out[base + j] = o[j];
    }
}

// Builds CpArgs::tab for one size (one thread per point, run once per context and size).
__global__ void __launch_bounds__(256) cp_table_kernel(uint32_t* tab, CpArgs c, PowTab pw)
{
    pdl_enter();
    const uint32_t N = 1u << c.log_len;                                  // local points (== 2^log_N when not sharded)
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const uint32_t x = mont_mul(GEN, pow_h_m(pw, shard_global(c.shard, i)));   // canonical 5 h^i, i the GLOBAL index
    const uint32_t xm = to_mont(x);
    const uint32_t d0m = to_mont(fsub(x, 1u)), d1m = to_mont(fsub(x, c.g2));
    const uint32_t xn_m = mont_pow(xm, (uint64_t)1 << c.log_n);          // x^n, one of 8 values
    const uint32_t zinv_m = mont_inv(fsub(xn_m, R1));                    // 1 / (x^n - 1), never singular on the coset
    const uint32_t cub_m = mont_mul(mont_mul(to_mont(fsub(x, c.g3)), d1m), to_mont(fsub(x, c.g1)));
    tab[i] = mont_inv(d0m);                                              // A_i R
    tab[N + i] = mont_inv(d1m);                                          // B_i R
    tab[2 * (size_t)N + i] = to_mont(mont_mul(cub_m, zinv_m));           // C_i R^2
}
cudaError_t launch_cp_tables(cudaStream_t s, uint32_t* tab, CpArgs c, PowTab pow)
{
    const uint32_t N = 1u << c.log_len;
    return zk_launch(cp_table_kernel, dim3((N + 255) / 256), dim3(256), 0, s, tab, c, pow);
}

template <int SRC>
__device__ __forceinline__ void produce_values(const TreeArgs& a, uint32_t b, uint32_t base, uint32_t count)
{
    if (SRC == SRC_FOLD) {
        const uint32_t* cur = a.fold.cur + (size_t)b * a.fold.cur_stride;
        uint32_t* next = a.vals + (size_t)b * a.vals_stride;
        uint32_t beta = a.fold.beta_from_state ? a.state[b].beta_raw[a.fold.layer_k] : a.fold.beta_raw;
        const FoldCoef fk = fold_coef(a.fold, a.pow, beta);
        if (count >= 8) { for (uint32_t o = 0; o < count; o += 8) fold_chunk<8>(a.fold, a.pow, cur, next, base + o, fk); }
        else if (count == 4) fold_chunk<4>(a.fold, a.pow, cur, next, base, fk);
        else if (count == 2) fold_chunk<2>(a.fold, a.pow, cur, next, base, fk);
        else fold_chunk<1>(a.fold, a.pow, cur, next, base, fk);
    } else if (SRC == SRC_CP_TAB) {
        const uint32_t* f = a.cp.f + (size_t)b * a.cp.f_stride;
        uint32_t* out = a.vals + (size_t)b * a.vals_stride;
        CpCoefTab kt;
        if (a.cp.alpha_from_state) cp_coef_tab(a.state[b].alpha_raw, a.state[b].a_first, a.state[b].a_last, kt);
        else cp_coef_tab(a.cp.alpha_raw, a.cp.a_first, a.cp.a_last, kt);
        if (count >= 4) {
#pragma unroll 1
            for (uint32_t o = 0; o < count; o += 4) cp_chunk_tab<4>(a.cp, f, out, base + o, kt);   // 4 per trip: 24 loads in flight, low register count
        }
        else if (count == 2) cp_chunk_tab<2>(a.cp, f, out, base, kt);
        else cp_chunk_tab<1>(a.cp, f, out, base, kt);
    } else if (SRC == SRC_CP) {
        const uint32_t* f = a.cp.f + (size_t)b * a.cp.f_stride;
        uint32_t* out = a.vals + (size_t)b * a.vals_stride;
        CpCoef k;
        if (a.cp.alpha_from_state) cp_coef(a.cp, a.state[b].alpha_raw, a.state[b].a_first, a.state[b].a_last, k);
        else cp_coef(a.cp, a.cp.alpha_raw, a.cp.a_first, a.cp.a_last, k);
        if (count >= 8) { for (uint32_t o = 0; o < count; o += 8) cp_chunk<8>(a.cp, a.pow, f, out, base + o, k); }
        else if (count == 4) cp_chunk<4>(a.cp, a.pow, f, out, base, k);
        else if (count == 2) cp_chunk<2>(a.cp, a.pow, f, out, base, k);
        else cp_chunk<1>(a.cp, a.pow, f, out, base, k);
    }
}

// ---------------------------------------------------------------- one out-of-line SHA-256 for latency-bound code
// The tile kernel and the channel hash a handful of messages one after the other; with every call site inlined
// the kernel carried ~6 copies of the 17 KB compression body and each copy was fetched cold exactly once
// (measured: ~2.7 us per copy on a 45 us kernel).  One out-of-line copy stays hot in the instruction cache.
// Message = 16 words (one block, already padded) or 64 bytes + the constant padding block (`two_blocks`).
struct Dig8 { uint32_t v[8]; };
template <int V>
__device__ __noinline__ Dig8 sha_msg(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                                     uint32_t w6, uint32_t w7, uint32_t w8, uint32_t w9, uint32_t w10, uint32_t w11,
                                     uint32_t w12, uint32_t w13, uint32_t w14, uint32_t w15, uint32_t two_blocks)
{
    uint32_t w[16] = {w0, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11, w12, w13, w14, w15};
    Dig8 d;
    sha::init_iv(d.v);
    sha::compress<V>(d.v, w);
    if (two_blocks) sha::compress_pad64<V>(d.v);
    return d;
}
template <int V>
__device__ __forceinline__ void msg_node(const uint32_t l[8], const uint32_t r[8], uint32_t out[8])   // merkle.rs:40-46
{
    Dig8 d = sha_msg<V>(l[0], l[1], l[2], l[3], l[4], l[5], l[6], l[7], r[0], r[1], r[2], r[3], r[4], r[5], r[6], r[7], 1u);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = d.v[i];
}
template <int V>
__device__ __forceinline__ void msg_leaf(uint32_t v, uint32_t out[8])                                 // merkle.rs:30-34
{
    Dig8 d = sha_msg<V>(v, 0x80000000u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u, 32u, 0u);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = d.v[i];
}

// CTA epilogue of the pass kernel: the PASS_THREADS subtree roots of a CTA (level `level`, indices first .. first + 255)
// are reduced `levels` more levels through shared memory, every node stored at its heap position.  The hashes of
// these narrow levels are latency bound, but unlike in a separate kernel the other resident CTAs of the SM keep the
// ALU pipe busy meanwhile, so the upper-tree kernel that follows starts from 2^-8 as many digests (1,024 instead of
// 262,144 at 2^23 leaves): 80 -> ~30 us per large tree.
template <int V>
__device__ __forceinline__ void pass_epilogue(uint8_t* heap, uint32_t level, uint32_t first, uint32_t levels, uint32_t d[8])
{
    __shared__ __align__(16) uint32_t epA[8][PASS_THREADS];
    __shared__ __align__(16) uint32_t epB[8][PASS_THREADS / 2];
    const uint32_t t = threadIdx.x;
#pragma unroll
    for (int w = 0; w < 8; w++) epA[w][t] = d[w];
    __syncthreads();
    uint32_t cnt = PASS_THREADS / 2, lvl = level - 1, base = first >> 1;
    bool inA = true;
    for (uint32_t e = 0; e < levels; e++) {
        if (t < cnt) {
            uint32_t l[8], r[8], o[8];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                const uint2 pr = inA ? *reinterpret_cast<const uint2*>(&epA[w][2 * t]) : *reinterpret_cast<const uint2*>(&epB[w][2 * t]);
                l[w] = pr.x; r[w] = pr.y;
            }
            msg_node<V>(l, r, o);                                           // merkle.rs:40-46
            sha::store_digest(heap + 32 * ((((size_t)1 << lvl) - 1) + base + t), o);
#pragma unroll
            for (int w = 0; w < 8; w++) { if (inA) epB[w][t] = o[w]; else epA[w][t] = o[w]; }
        }
        __syncthreads();
        inA = !inA; cnt >>= 1; lvl--; base >>= 1;
    }
}

// ---------------------------------------------------------------- pass kernel
// Thread t of proof b owns inputs [t*2^k, (t+1)*2^k) at heap level `in_level` and writes
// every node of the subtree above them (levels in_level-1 .. in_level-k); for value
// sources also the leaf digests themselves.
template <int SRC, int V = ZK_SHA_V, int UNI = ZK_SHA_UNI>
__global__ void __launch_bounds__(PASS_THREADS, ZK_PASS_MINBLOCKS) tree_pass_kernel(const TreeArgs a)
{
    pdl_wait();
    const uint32_t b = blockIdx.y;
    const uint32_t t = blockIdx.x * PASS_THREADS + threadIdx.x;
    const uint32_t n_threads = 1u << (a.in_level - a.k);
    const uint32_t cnt = 1u << a.k;
    const uint32_t base = t << a.k;
    uint8_t* heap = a.heap + (size_t)b * a.heap_stride;

    if (SRC == SRC_FOLD || SRC == SRC_CP_TAB) {
        // The CTA produces the values of all its threads cooperatively, four consecutive points per thread and trip, so
        // the reads of the previous layer / of f and the constant tables and the writes of the new layer are coalesced
        // 128-bit accesses (a thread walking its own 2^k points touches 32 different lines per warp instruction);
        // the hashing below then reads the CTA's own 2^k * 256 values back from L1 / L2.
        const uint32_t cta_base = (blockIdx.x * PASS_THREADS) << a.k;
        const uint32_t total = 1u << a.in_level;
        if (cnt >= 4) {
            const uint32_t end = (PASS_THREADS << a.k) < total - cta_base ? (PASS_THREADS << a.k) : total - cta_base;
            if (SRC == SRC_FOLD) {
                const uint32_t* cur = a.fold.cur + (size_t)b * a.fold.cur_stride;
                uint32_t* next = a.vals + (size_t)b * a.vals_stride;
                const FoldCoef fk = fold_coef(a.fold, a.pow, a.fold.beta_from_state ? a.state[b].beta_raw[a.fold.layer_k] : a.fold.beta_raw);
                for (uint32_t o = threadIdx.x * 4; o < end; o += PASS_THREADS * 4) fold_chunk<4>(a.fold, a.pow, cur, next, cta_base + o, fk);
            } else {
                const uint32_t* f = a.cp.f + (size_t)b * a.cp.f_stride;
                uint32_t* out = a.vals + (size_t)b * a.vals_stride;
                CpCoefTab kt;
                if (a.cp.alpha_from_state) cp_coef_tab(a.state[b].alpha_raw, a.state[b].a_first, a.state[b].a_last, kt);
                else cp_coef_tab(a.cp.alpha_raw, a.cp.a_first, a.cp.a_last, kt);
                for (uint32_t o = threadIdx.x * 4; o < end; o += PASS_THREADS * 4) cp_chunk_tab<4>(a.cp, f, out, cta_base + o, kt);
            }
        } else if (t < n_threads) {
            produce_values<SRC>(a, b, base, cnt);
        }
        __syncthreads();
    }
    if (t >= n_threads) return;
    if (SRC == SRC_CP) produce_values<SRC>(a, b, base, cnt);
    const uint32_t* vals = a.vals + (size_t)b * a.vals_stride;

    uint32_t stack[MAX_PASS_K][8];
    uint32_t d[8];
    const size_t lvl0 = ((size_t)1 << a.in_level) - 1;
    if (UNI && SRC != SRC_DIGESTS) {
        // One shared copy of the round loop for leaves and nodes (small instruction footprint): the message is
        // set up first, then the same code hashes it; nodes add the constant padding block.
        for (uint32_t i = 0; i < cnt; i++) {
            uint32_t w[16];
            w[0] = vals[base + i];                                          // merkle.rs:30-34
            if (UNI == 1) {
                w[1] = 0x80000000u;
#pragma unroll
                for (int j = 2; j < 15; j++) w[j] = 0u;
                w[15] = 32u;
            }
            uint32_t m = i, lvl = 0, idx = base + i;
            bool node = false;
            for (;;) {
                if (UNI == 1) {
                    sha::init_iv(d);
                    sha::compress<V>(d, w);
                } else {
                    sha::compress_leaf_or_node<V>(d, w, !node);
                }
                if (node) sha::compress_pad64<V>(d);                        // merkle.rs:40-46
                sha::store_digest(heap + 32 * ((((size_t)1 << (a.in_level - lvl)) - 1) + idx), d);
                if (!(m & 1u)) break;
#pragma unroll
                for (int j = 0; j < 8; j++) { w[j] = stack[lvl][j]; w[8 + j] = d[j]; }
                node = true;
                idx >>= 1; m >>= 1; lvl++;
            }
            if (lvl < a.k) {
#pragma unroll
                for (int j = 0; j < 8; j++) stack[lvl][j] = d[j];
            }
        }
        // d: root of this thread's 2^k-leaf subtree (heap level in_level - k, index t)
        if (a.ep_levels) pass_epilogue<V>(heap, a.in_level - a.k, blockIdx.x * PASS_THREADS, a.ep_levels, d);
        return;
    }
    for (uint32_t i = 0; i < cnt; i++) {
        if (SRC == SRC_DIGESTS) {
            sha::load_digest(heap + 32 * (lvl0 + base + i), d);
        } else {
            sha::hash_leaf<V>(vals[base + i], d);                           // merkle.rs:30-34
            sha::store_digest(heap + 32 * (lvl0 + base + i), d);
        }
        uint32_t m = i, lvl = 0, idx = base + i;
        while (m & 1u) {                                                    // right child: merge with parked left
            uint32_t l[8];
#pragma unroll
            for (int w = 0; w < 8; w++) l[w] = stack[lvl][w];
            sha::hash_node<V>(l, d, d);                                     // merkle.rs:40-46
            idx >>= 1; m >>= 1; lvl++;
            sha::store_digest(heap + 32 * ((((size_t)1 << (a.in_level - lvl)) - 1) + idx), d);
        }
        if (lvl < a.k) {
#pragma unroll
            for (int w = 0; w < 8; w++) stack[lvl][w] = d[w];
        }
    }
    if (a.ep_levels) pass_epilogue<V>(heap, a.in_level - a.k, blockIdx.x * PASS_THREADS, a.ep_levels, d);
}

// ---------------------------------------------------------------- device channel
// channel.rs:19-26 for a 32-byte value: state = SHA256(state || v) -- a 64-byte message,
// i.e. exactly the node hash.
__device__ __forceinline__ void chan_commit_digest(uint32_t st[8], const uint32_t v[8])
{
    uint32_t out[8];
    msg_node<ZK_TILE_SHA_V>(st, v, out);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = out[i];
}
// channel.rs:19-26 for a u32: bincode writes it little-endian (4 bytes); the 36-byte message
// state || le(v) is one padded block.
__device__ __forceinline__ void chan_commit_u32(uint32_t st[8], uint32_t v)
{
    Dig8 d = sha_msg<ZK_TILE_SHA_V>(st[0], st[1], st[2], st[3], st[4], st[5], st[6], st[7], sha::bswap32(v), 0x80000000u,
                                    0u, 0u, 0u, 0u, 0u, 36u * 8u, 0u);
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = d.v[i];
}
// channel.rs:28-32
__device__ __forceinline__ uint32_t chan_get_u32(uint32_t st[8])
{
    uint32_t f = st[0];  // be32(state[0..4]) == H0
    chan_commit_u32(st, f);
    return f;
}
__device__ __forceinline__ void tr_put_u32(uint8_t* tr, size_t off, uint32_t v)
{
    *reinterpret_cast<uint32_t*>(tr + off) = v;  // little-endian host/device, 4-byte aligned offsets
}
__device__ __forceinline__ void tr_put_digest(uint8_t* tr, size_t off, const uint32_t d[8])
{
#pragma unroll
    for (int i = 0; i < 8; i++) tr_put_u32(tr, off + 4 * i, sha::bswap32(d[i]));
}

__device__ void post_root(const TreeArgs& a, uint32_t b, const uint32_t root[8])
{
    ProofState* ps = a.state + b;
    uint8_t* tr = a.transcript + (size_t)b * a.transcript_stride;
    uint32_t st[8];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = ps->chan[i];
    const uint32_t L = a.n_layers;
    if (a.post == POST_F) {
        tr_put_digest(tr, 0, root);
        chan_commit_digest(st, root);                                       // prover.rs:85
        for (int j = 0; j < 3; j++) {                                       // prover.rs:163-165
            uint32_t v = chan_get_u32(st);
            ps->alpha_raw[j] = v;
            tr_put_u32(tr, 32 + 4 * j, v);
        }
    } else if (a.post == POST_CP) {
        tr_put_digest(tr, 44, root);
        chan_commit_digest(st, root);                                       // prover.rs:180
        uint32_t v = chan_get_u32(st);                                      // prover.rs:200 (first layer)
        ps->beta_raw[0] = v;
        tr_put_u32(tr, 76, v);
    } else if (a.post == POST_FRI) {
        const uint32_t i = a.fri_layer;
        tr_put_digest(tr, 80 + 36 * (size_t)(i - 1), root);
        chan_commit_digest(st, root);                                       // prover.rs:224
        if (i < L) {
            uint32_t v = chan_get_u32(st);                                  // prover.rs:200
            ps->beta_raw[i] = v;
            tr_put_u32(tr, 76 + 36 * (size_t)i, v);
        } else {
            uint32_t ft = a.vals[(size_t)b * a.vals_stride];                // prover.rs:254
            ps->free_term = ft;
            chan_commit_u32(st, ft);
            tr_put_u32(tr, 76 + 36 * (size_t)L, ft);
            uint32_t q = chan_get_u32(st);                                  // prover.rs:263
            ps->query_raw = q;
            tr_put_u32(tr, 80 + 36 * (size_t)L, q);
#pragma unroll
            for (int w = 0; w < 8; w++) ps->chan_at_query[w] = st[w];
        }
    }
#pragma unroll
    for (int i = 0; i < 8; i++) ps->chan[i] = st[i];
}

// ---------------------------------------------------------------- tile kernel (upper tree in ONE launch)
// A tile is up to 2^TILE_LOG consecutive inputs of level `in_level` -- leaf values (hashed here, after the fused
// fold / composition source produced them) or digests left by a pass kernel.  One CTA reduces its tile to a
// single digest level by level (two inputs per thread, then word-major ping-pong buffers in shared memory:
// conflict-free 64-bit reads, one barrier per level), writing every node into the heap.  The CTA that finishes
// LAST (per-proof arrival counter, threadfence-reduction pattern: no co-residency requirement, no grid sync)
// reduces the tile roots to the tree root and runs the Fiat-Shamir step (channel.rs:19-32), so a whole upper
// tree + challenge draw is a single launch instead of four digest passes + a top kernel.
//
// Narrow levels are LATENCY bound: a lone warp needs ~3.0 us per node (measured in-kernel, profiles/README r1e) because
// its 18 ALU-pipe instructions per round issue at one per 2 cycles, while the dependency chain of a round is only
// ~20 cycles.  Levels with <= 32*COOP_PAIRS nodes therefore hash every node with TWO warps on two different SM
// sub-partitions: the helper warp computes the message schedule K[t] + W[t], t = 16..63, into shared memory, 16 words
// at a time, and the main warp runs the 64 rounds consuming them (named barriers 1..12, helper arrives / main
// syncs), then the constant padding block.  For leaf sources the pair also hashes the left / right leaf in parallel.
constexpr uint32_t COOP_PAIRS = 4;              // named barriers 1 + 3*pair + {0,1,2} <= 12
constexpr uint32_t COOP_WORDS = 64 * 32;        // per pair: 48 schedule words + 16 exchange words, one column per lane
extern __shared__ __align__(16) uint32_t tile_smem[];

__device__ __forceinline__ void bar_sync(uint32_t id, uint32_t n)
{
    asm volatile("barrier.sync.aligned %0, %1;" ::"r"(id), "r"(n) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint32_t id, uint32_t n)
{
    asm volatile("barrier.arrive.aligned %0, %1;" ::"r"(id), "r"(n) : "memory");
}

// main warp of a pair: node hash (merkle.rs:40-46) with the schedule of rounds 16..63 read from tile_smem[kw + ..]
template <int V>
__device__ __noinline__ Dig8 coop_main(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                                       uint32_t w6, uint32_t w7, uint32_t w8, uint32_t w9, uint32_t w10, uint32_t w11,
                                       uint32_t w12, uint32_t w13, uint32_t w14, uint32_t w15, uint32_t kw, uint32_t bar)
{
    constexpr sha::Consts C = sha::make_consts();
    using sha::addv; using sha::bsig0; using sha::bsig1; using sha::ch; using sha::maj;
    const uint32_t w[16] = {w0, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11, w12, w13, w14, w15};
    uint32_t a = ZK_SHA_IV0, b = ZK_SHA_IV1, c = ZK_SHA_IV2, d = ZK_SHA_IV3, e = ZK_SHA_IV4, f = ZK_SHA_IV5,
             g = ZK_SHA_IV6, h = ZK_SHA_IV7;
#define ZK_KW_FIRST(j) (addv<V, 32>(w[j], C.k[j]))
    ZK_SHA_16ROUNDS(V, ZK_KW_FIRST)
#undef ZK_KW_FIRST
#pragma unroll 1
    for (uint32_t ck = 0; ck < 3; ck++) {
        bar_sync(bar + ck, 64);                                    // the helper has published words 16*ck .. 16*ck+15
        uint32_t kk[16];
#pragma unroll
        for (int j = 0; j < 16; j++) kk[j] = tile_smem[kw + (16 * ck + j) * 32];
#define ZK_KW_COOP(j) (kk[j])
        ZK_SHA_16ROUNDS(V, ZK_KW_COOP)
#undef ZK_KW_COOP
    }
    Dig8 o;
    o.v[0] = a + ZK_SHA_IV0; o.v[1] = b + ZK_SHA_IV1; o.v[2] = c + ZK_SHA_IV2; o.v[3] = d + ZK_SHA_IV3;
    o.v[4] = e + ZK_SHA_IV4; o.v[5] = f + ZK_SHA_IV5; o.v[6] = g + ZK_SHA_IV6; o.v[7] = h + ZK_SHA_IV7;
    sha::compress_pad64<V>(o.v);
    return o;
}
// helper warp of a pair: K[t] + W[t] for t = 16..63 of the same 16-word message
template <int V>
__device__ __noinline__ void coop_helper(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t w4, uint32_t w5,
                                         uint32_t w6, uint32_t w7, uint32_t w8, uint32_t w9, uint32_t w10, uint32_t w11,
                                         uint32_t w12, uint32_t w13, uint32_t w14, uint32_t w15, uint32_t kw, uint32_t bar)
{
    using sha::addv; using sha::ssig0; using sha::ssig1;
    uint32_t w[16] = {w0, w1, w2, w3, w4, w5, w6, w7, w8, w9, w10, w11, w12, w13, w14, w15};
#pragma unroll 1
    for (uint32_t ck = 0; ck < 3; ck++) {
        uint32_t kk[16];
        {
            const uint4* kp = reinterpret_cast<const uint4*>(sha::c_k + 16 + 16 * ck);
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 t = kp[q];
                kk[4 * q] = t.x; kk[4 * q + 1] = t.y; kk[4 * q + 2] = t.z; kk[4 * q + 3] = t.w;
            }
        }
#pragma unroll
        for (int j = 0; j < 16; j++) {
            w[j] = addv<V, 2>(addv<V, 2>(addv<V, 2>(w[j], ssig1<V>(w[(j + 14) & 15])), w[(j + 9) & 15]),
                              ssig0<V>(w[(j + 1) & 15]));
            tile_smem[kw + (16 * ck + j) * 32] = addv<V, 2>(w[j], kk[j]);
        }
        bar_arrive(bar + ck, 64);
    }
}

struct TileBufs {
    uint32_t* A;       // [8][P]
    uint32_t* B;       // [8][P/2]
    uint32_t P;        // == blockDim.x
    uint32_t coop;     // word offset of the cooperative area inside tile_smem
    uint32_t coop_cnt; // levels with at most this many nodes are hashed cooperatively (0: never)
};

template <int SRC, int V>
__device__ __forceinline__ void tile_input(const TreeArgs& a, uint8_t* heap, const uint32_t* vals, uint32_t level,
                                           uint32_t idx, uint32_t d[8], bool coherent)
{
    const size_t node = (((size_t)1 << level) - 1) + idx;
    if (SRC == SRC_DIGESTS) {
        if (coherent) sha::load_digest_cg(heap + 32 * node, d);   // written by other CTAs of this launch
        else sha::load_digest(heap + 32 * node, d);
    } else {
        msg_leaf<V>(vals[idx], d);                                 // merkle.rs:30-34
        sha::store_digest(heap + 32 * node, d);
    }
}

// One level hashed by warp pairs.  `cnt` nodes (<= 32 * COOP_PAIRS, 2 * ceil(cnt/32) warps exist); node j of the
// level is column j.  FIRST: the children are tile inputs (heap digests / leaf values), else they are columns 2j, 2j+1
// of `in` (pitch in_p).  Results go to `out` (pitch out_p) and to the heap at `heap_lvl + j`.
template <int SRC, int V, bool FIRST>
__device__ __forceinline__ void coop_level(const TreeArgs& a, uint8_t* heap, const uint32_t* vals, uint32_t level,
                                           uint32_t first, bool coherent, uint32_t cnt, const uint32_t* in, uint32_t in_p,
                                           uint32_t* out, uint32_t out_p, size_t heap_lvl, uint32_t coop, uint32_t d[8])
{
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nm = (cnt + 31) >> 5;
    if (warp >= 2 * nm) return;
    const bool is_main = warp < nm;
    const uint32_t pair = is_main ? warp : warp - nm;
    const uint32_t j = pair * 32 + lane, jc = j < cnt ? j : 0;      // lanes past the level hash a duplicate, store nothing
    const uint32_t area = coop + pair * COOP_WORDS + lane, bar = 1 + 3 * pair;
    uint32_t m[16];
    if (FIRST && SRC != SRC_DIGESTS) {
        // the pair hashes the two leaves side by side and swaps digests through shared memory (barrier bar+2 is free
        // until the last schedule chunk)
        uint32_t x[8];
        tile_input<SRC, V>(a, heap, vals, level, first + 2 * jc + (is_main ? 0 : 1), x, coherent);
#pragma unroll
        for (int w = 0; w < 8; w++) tile_smem[area + (48 + (is_main ? 0 : 8) + w) * 32] = x[w];
        bar_sync(bar + 2, 64);
#pragma unroll
        for (int w = 0; w < 16; w++) m[w] = tile_smem[area + (48 + w) * 32];
    } else if (FIRST) {
        tile_input<SRC, V>(a, heap, vals, level, first + 2 * jc, m, coherent);
        tile_input<SRC, V>(a, heap, vals, level, first + 2 * jc + 1, m + 8, coherent);
    } else {
#pragma unroll
        for (int w = 0; w < 8; w++) {
            uint2 p = *reinterpret_cast<const uint2*>(in + w * in_p + 2 * jc);
            m[w] = p.x; m[8 + w] = p.y;
        }
    }
    if (is_main) {
        Dig8 r = coop_main<ZK_COOP_V>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13],
                              m[14], m[15], area, bar);
#pragma unroll
        for (int w = 0; w < 8; w++) d[w] = r.v[w];
        if (j < cnt) {
            sha::store_digest(heap + 32 * (heap_lvl + j), d);
#pragma unroll
            for (int w = 0; w < 8; w++) out[w * out_p + j] = d[w];
        }
    } else {
        coop_helper<ZK_COOP_HELPER_V>(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8], m[9], m[10], m[11], m[12], m[13], m[14],
                       m[15], area, bar);
    }
}

// Reduces inputs [first, first + 2^log_cnt) of heap level `level` to one digest; every produced node is stored at
// its heap position.  Returns with the subtree root in `root` of thread 0.
template <int SRC, int V>
__device__ __forceinline__ void tile_reduce(const TreeArgs& a, uint8_t* heap, const uint32_t* vals, uint32_t level,
                                            uint32_t first, uint32_t log_cnt, bool coherent, const TileBufs& tb,
                                            uint32_t root[8])
{
    const uint32_t t = threadIdx.x;
    if (log_cnt == 0) {                                            // a single input: it is the subtree root
        if (t == 0) tile_input<SRC, V>(a, heap, vals, level, first, root, coherent);
        return;
    }
    uint32_t cnt = 1u << (log_cnt - 1);                            // nodes of the level being produced
    uint32_t lvl = level - 1;
    uint32_t d[8];
    uint32_t base = first >> 1;
    if (cnt <= tb.coop_cnt && 64 * ((cnt + 31) >> 5) <= tb.P) {
        coop_level<SRC, V, true>(a, heap, vals, level, first, coherent, cnt, nullptr, 0, tb.A, tb.P,
                                 (((size_t)1 << lvl) - 1) + base, tb.coop, d);
    } else if (t < cnt) {
        uint32_t l[8], r[8];
        tile_input<SRC, V>(a, heap, vals, level, first + 2 * t, l, coherent);
        tile_input<SRC, V>(a, heap, vals, level, first + 2 * t + 1, r, coherent);
        msg_node<V>(l, r, d);                                      // merkle.rs:40-46
        sha::store_digest(heap + 32 * ((((size_t)1 << lvl) - 1) + base + t), d);
#pragma unroll
        for (int w = 0; w < 8; w++) tb.A[w * tb.P + t] = d[w];
    }
    __syncthreads();
    bool inA = true;
    while (cnt > 1) {
        cnt >>= 1; lvl--; base >>= 1;
        const uint32_t* in = inA ? tb.A : tb.B;
        uint32_t* out = inA ? tb.B : tb.A;
        const uint32_t in_p = inA ? tb.P : tb.P / 2, out_p = inA ? tb.P / 2 : tb.P;
        if (cnt <= tb.coop_cnt && 64 * ((cnt + 31) >> 5) <= tb.P) {
            coop_level<SRC, V, false>(a, heap, vals, level, first, coherent, cnt, in, in_p, out, out_p,
                                      (((size_t)1 << lvl) - 1) + base, tb.coop, d);
        } else if (t < cnt) {
            uint32_t l[8], r[8];
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint2 p = *reinterpret_cast<const uint2*>(in + w * in_p + 2 * t);
                l[w] = p.x; r[w] = p.y;
            }
            msg_node<V>(l, r, d);
            sha::store_digest(heap + 32 * ((((size_t)1 << lvl) - 1) + base + t), d);
#pragma unroll
            for (int w = 0; w < 8; w++) out[w * out_p + t] = d[w];
        }
        __syncthreads();
        inA = !inA;
    }
    if (t == 0) {
#pragma unroll
        for (int w = 0; w < 8; w++) root[w] = d[w];
    }
}

template <int SRC, int V = ZK_TILE_SHA_V>
__global__ void __launch_bounds__(TILE_PAIRS) tree_tile_kernel(const TreeArgs a)
{
    pdl_wait();
    TileBufs tb;
    tb.P = blockDim.x;
    tb.A = tile_smem;
    tb.B = tile_smem + 8 * tb.P;
    tb.coop = 12 * tb.P;
    tb.coop_cnt = a.coop_cnt;
    const uint32_t b = blockIdx.y, tile = blockIdx.x, n_tiles = gridDim.x, t = threadIdx.x;
    const uint32_t log_tile = a.in_level < a.log_tile ? a.in_level : a.log_tile;   // n_tiles == 2^(in_level - log_tile)
    uint8_t* heap = a.heap + (size_t)b * a.heap_stride;
    const uint32_t first = tile << log_tile;

    if (SRC == SRC_FOLD || SRC == SRC_CP || SRC == SRC_CP_TAB) {
        // every layer has an even number (>= 8) of points: two consecutive points per thread
        if (2 * t < (1u << log_tile)) produce_values<SRC>(a, b, first + 2 * t, 2);
        __syncthreads();                                           // CP reads neighbours' f only; fold reads cur only
    }
    const uint32_t* vals = a.vals + (size_t)b * a.vals_stride;
    uint32_t root[8];
    tile_reduce<SRC, V>(a, heap, vals, a.in_level, first, log_tile, false, tb, root);
    if (n_tiles > 1) {
        __threadfence();                                           // this tile's nodes are visible device-wide ...
        __syncthreads();
        int last = 0;
        if (t == 0) {
            const uint32_t prev = atomicAdd(a.counters + b, 1u);   // ... before the arrival is counted
            last = (prev == n_tiles - 1);
            if (last) a.counters[b] = 0;                           // re-arm for the next launch on this stream
        }
        if (!__syncthreads_or(last)) return;
        __threadfence();
        tile_reduce<SRC_DIGESTS, V>(a, heap, vals, a.in_level - log_tile, 0, a.in_level - log_tile, true, tb, root);
    }
    pdl_trigger();                                                 // every other CTA has exited: the next kernel may move in
    if (t == 0 && a.post != POST_NONE) post_root(a, b, root);
}

// ---------------------------------------------------------------- FRI tail: the small layers of a proof in ONE launch
// Once a layer fits a single tile (<= 2^TILE_MIN_LOG values) nothing is left to spread over SMs: fold, leaf hashes, the
// levels, the channel step that draws the next beta (prover.rs:200-224) -- and the same for every following layer down to
// the 8-value last one with its free term and query draw (prover.rs:254,263) -- run back to back in one CTA per proof.
// Saves a launch (and its dependent-launch hand-over) per layer on the latency path of every proof.
template <int V = ZK_TILE_SHA_V>
__global__ void __launch_bounds__(TILE_PAIRS) fri_tail_kernel(const TreeArgs a0, uint32_t last_layer)
{
    pdl_wait();
    TreeArgs a = a0;                                               // the layer-dependent fields advance in the loop
    TileBufs tb;
    tb.P = blockDim.x;
    tb.A = tile_smem;
    tb.B = tile_smem + 8 * tb.P;
    tb.coop = 12 * tb.P;
    tb.coop_cnt = a.coop_cnt;
    const uint32_t b = blockIdx.y, t = threadIdx.x;
    for (;;) {
        uint8_t* heap = a.heap + (size_t)b * a.heap_stride;
        if (2 * t < (1u << a.log_size)) produce_values<SRC_FOLD>(a, b, 2 * t, 2);   // polynomial.rs:385-400 in evaluation form
        __syncthreads();
        const uint32_t* vals = a.vals + (size_t)b * a.vals_stride;
        uint32_t root[8];
        tile_reduce<SRC_FOLD, V>(a, heap, vals, a.log_size, 0, a.log_size, false, tb, root);
        const bool done = a.fri_layer >= last_layer;
        if (done) pdl_trigger();
        if (t == 0) post_root(a, b, root);                         // commit the root; next beta, or free term + query
        if (done) break;
        __syncthreads();                                           // beta and the layer's values are visible to the whole CTA
        const uint32_t len = 1u << a.log_size;
        a.fold.cur = a.vals;
        a.fold.cur_stride = a.vals_stride;
        a.fold.layer_k += 1;
        a.fold.winv2k_m = mont_mul(a.fold.winv2k_m, a.fold.winv2k_m);   // (5^(2^(k+1)))^-1, Montgomery form
        a.vals += len;
        a.heap += 64 * (size_t)len;
        a.log_size -= 1;
        a.in_level = a.log_size;
        a.fold.log_half = a.log_size;
        a.fri_layer += 1;
    }
}

// ---------------------------------------------------------------- stand-alone stage kernels
__global__ void __launch_bounds__(256) fold_kernel(const uint32_t* cur, uint32_t* next, FoldArgs f, PowTab pw,
                                                   uint64_t next_stride, const ProofState* st)
{
    pdl_enter();
    const uint32_t b = blockIdx.y;
    const uint32_t half = 1u << f.log_half;
    const uint32_t* c = cur + (size_t)b * f.cur_stride;
    uint32_t* o = next + (size_t)b * next_stride;
    const uint32_t beta = (f.beta_from_state && st) ? st[b].beta_raw[f.layer_k] : f.beta_raw;
    if (half >= 4) {
        uint32_t base = (blockIdx.x * 256 + threadIdx.x) * 4;
        if (base < half) fold_chunk<4>(f, pw, c, o, base, fold_coef(f, pw, beta));
    } else {
        uint32_t base = blockIdx.x * 256 + threadIdx.x;
        if (base < half) fold_chunk<1>(f, pw, c, o, base, fold_coef(f, pw, beta));
    }
}

template <bool TAB>
__global__ void __launch_bounds__(256) cp_kernel(uint32_t* out, CpArgs c, PowTab pw, uint64_t out_stride,
                                                 const ProofState* st)
{
    pdl_enter();
    const uint32_t b = blockIdx.y;
    const uint32_t N = 1u << c.log_len;
    const uint32_t base = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (base >= N) return;
    const bool from_state = c.alpha_from_state && st;
    const uint32_t* al = from_state ? st[b].alpha_raw : c.alpha_raw;
    const uint32_t a_first = reduce_u32(from_state ? st[b].a_first : c.a_first);
    const uint32_t a_last = reduce_u32(from_state ? st[b].a_last : c.a_last);
    if (TAB) {
        CpCoefTab k;
        cp_coef_tab(al, a_first, a_last, k);
        cp_chunk_tab<4>(c, c.f + (size_t)b * c.f_stride, out + (size_t)b * out_stride, base, k);
    } else {
        CpCoef k;
        cp_coef(c, al, a_first, a_last, k);
        cp_chunk<4>(c, pw, c.f + (size_t)b * c.f_stride, out + (size_t)b * out_stride, base, k);
    }
}

// ---------------------------------------------------------------- decommit gather (prover.rs:263-289)
// One CTA per proof.  Writes the bincode records of Part 4 straight into the transcript:
//   4 x (u32 value | u64 len | len digests)          f[x], f[x+8], f[x+16], cp[x]
//   L x (u32 cp_i[x_i] | u32 cp_i[nx_i] | u64 len | path(x_i) | u64 len | path(nx_i))
// merkle.rs:54-71 `trace`: sibling of heap index i is i-1 (i even) or i+1 (i odd), leaf -> root.
__device__ __forceinline__ size_t layer_vals_off_dev(uint32_t log_N, uint32_t i)   // sum_{j<i} 2^(log_N-j)
{
    return ((size_t)2 << log_N) - ((size_t)2 << (log_N - i));
}
__device__ __forceinline__ size_t layer_heap_off_dev(uint32_t log_N, uint32_t i)   // common.cuh layer_heap_off: sum_{j<i} 64*2^(log_N-j)
{
    return 64 * (((size_t)2 << log_N) - ((size_t)2 << (log_N - i)));
}
struct DecommitArgs {
    uint32_t log_n;
    const uint32_t* f_eval; uint64_t f_stride;
    const uint8_t* f_heap; uint64_t f_heap_stride;
    const uint32_t* layers; uint64_t layers_stride;
    const uint8_t* layer_heaps; uint64_t layer_heaps_stride;
    ProofState* state;
    uint8_t* transcript; uint64_t transcript_stride;
    // Extra openings (zk_open_queries): when xs != null the kernel serves query positions xs[blockIdx.y] of ONE proof
    // (the base pointers already point at it) and writes each record set to out + blockIdx.y * out_stride instead of
    // the transcript.
    const uint32_t* xs;
    uint8_t* out; uint64_t out_stride;
};

// `step` levels above leaf `leaf` of a 2^log_size-leaf heap, the sibling (merkle.rs:58-66)
__device__ __forceinline__ size_t sibling_node(uint32_t log_size, uint32_t leaf, uint32_t step)
{
    const size_t i = (((size_t)1 << (log_size - step)) - 1) + (leaf >> step);  // ancestor, heap index
    return (i & 1) ? i + 1 : i - 1;
}

// Every sibling digest of every record is an independent 32-byte gather, so the work is a flat list of
// (record, path, step) items spread over the CTA: all loads are in flight at once (the first version walked
// the 4 + L records one after the other, ~24 dependent DRAM round trips = 60 us at k = 20).
__global__ void __launch_bounds__(256) decommit_kernel(const DecommitArgs a)
{
    pdl_enter();
    const bool extra = a.xs != nullptr;                          // openings of further positions of one proof
    const uint32_t b = extra ? 0u : blockIdx.y;                  // proof; blockIdx.x: slice of the gather items
    const uint32_t k = a.log_n, L = k, log_N = k + LOG_BLOWUP, N = 1u << log_N, B = 1u << LOG_BLOWUP;
    ProofState* ps = a.state + b;
    const size_t off0 = 84 + 36 * (size_t)L;
    uint8_t* tr = extra ? a.out + (size_t)blockIdx.y * a.out_stride - off0 : a.transcript + (size_t)b * a.transcript_stride;
    const uint32_t* f = a.f_eval + (size_t)b * a.f_stride;
    const uint8_t* fh = a.f_heap + (size_t)b * a.f_heap_stride;
    const uint32_t* lay = a.layers + (size_t)b * a.layers_stride;
    const uint8_t* lh = a.layer_heaps + (size_t)b * a.layer_heaps_stride;
    const uint32_t x = extra ? a.xs[blockIdx.y] : (uint32_t)((size_t)ps->query_raw % (size_t)(N - 2 * B));  // prover.rs:263
    const size_t rec0 = 12 + 32 * (size_t)log_N;
    if (!extra && blockIdx.x == 0 && threadIdx.x == 0) {
        // prover.rs:238: the last FRI polynomial has degree 0, i.e. its 8 evaluations are equal
        const size_t voff = layer_vals_off_dev(log_N, L);
        uint32_t bad = 0;
        for (uint32_t j = 1; j < (N >> L); j++) bad |= (lay[voff + j] != lay[voff]);
        ps->pad[0] = bad;
    }
    // values and length prefixes: one thread per record
    for (uint32_t r = threadIdx.x; blockIdx.x == 0 && r < 4 + L; r += blockDim.x) {
        if (r < 4) {                                             // prover.rs:266-277
            uint8_t* rec = tr + off0 + r * rec0;
            *reinterpret_cast<uint32_t*>(rec) = (r < 3) ? f[x + r * B] : lay[x];
            *reinterpret_cast<uint32_t*>(rec + 4) = log_N;       // u64 LE length
            *reinterpret_cast<uint32_t*>(rec + 8) = 0;
        } else {                                                 // prover.rs:280-289
            const uint32_t i = r - 4, lg = log_N - i, len = 1u << lg;
            const uint32_t xi = x & (len - 1), nx = (xi + len / 2) & (len - 1);
            const uint32_t* vals = lay + layer_vals_off_dev(log_N, i);
            // records before this one: sum_{j<i} (24 + 64 (log_N - j))
            uint8_t* rec = tr + off0 + 4 * rec0 + 24 * (size_t)i + 64 * ((size_t)i * log_N - (size_t)i * (i - 1) / 2);
            *reinterpret_cast<uint32_t*>(rec) = vals[xi];
            *reinterpret_cast<uint32_t*>(rec + 4) = vals[nx];
            *reinterpret_cast<uint32_t*>(rec + 8) = lg;
            *reinterpret_cast<uint32_t*>(rec + 12) = 0;
            *reinterpret_cast<uint32_t*>(rec + 16 + 32 * (size_t)lg) = lg;
            *reinterpret_cast<uint32_t*>(rec + 20 + 32 * (size_t)lg) = 0;
        }
    }
    // sibling digests: item = (record r, path p, step s); 8 lanes move one digest (4 bytes each)
    const uint32_t per_rec_f = log_N;
    const uint32_t n_items = 4 * per_rec_f + 2 * (L * log_N - L * (L - 1) / 2);
    const uint32_t sub = threadIdx.x & 7;
    for (uint32_t it = blockIdx.x * (blockDim.x >> 3) + (threadIdx.x >> 3); it < n_items; it += gridDim.x * (blockDim.x >> 3)) {
        const uint8_t* heap;
        uint8_t* dst;
        uint32_t lg, leaf, s;
        if (it < 4 * per_rec_f) {
            const uint32_t r = it / per_rec_f;
            s = it - r * per_rec_f;
            lg = log_N;
            leaf = (r < 3) ? x + r * B : x;
            heap = (r < 3) ? fh : lh;
            dst = tr + off0 + r * rec0 + 12 + 32 * (size_t)s;
        } else {
            uint32_t rem = it - 4 * per_rec_f, i = 0;
            while (rem >= 2 * (log_N - i)) { rem -= 2 * (log_N - i); i++; }   // <= L iterations, no memory traffic
            lg = log_N - i;
            const uint32_t len = 1u << lg, xi = x & (len - 1), nx = (xi + len / 2) & (len - 1);
            const uint32_t path = rem >= lg;
            s = path ? rem - lg : rem;
            leaf = path ? nx : xi;
            heap = lh + layer_heap_off_dev(log_N, i);
            uint8_t* rec = tr + off0 + 4 * rec0 + 24 * (size_t)i + 64 * ((size_t)i * log_N - (size_t)i * (i - 1) / 2);
            dst = rec + (path ? 24 + 32 * (size_t)lg : 16) + 32 * (size_t)s;
        }
        const uint32_t* src = reinterpret_cast<const uint32_t*>(heap + 32 * sibling_node(lg, leaf, s));
        reinterpret_cast<uint32_t*>(dst)[sub] = src[sub];        // transcript offsets are 4-byte aligned
    }
}

// The same records for a proof whose arrays are owned blocked-cyclically by several ranks (common.cuh ShardMap):
// every rank runs this kernel over its own arrays and writes only what it owns -- a value, the log_block rank-local
// siblings below its block root and the siblings above it from the replicated upper heap; rank 0 adds the length
// prefixes and the layers of the replicated FRI tail.  Everything else stays zero, so the byte-wise maximum over the
// ranks is the decommitment part of the transcript (one small all-reduce instead of per-record host round trips).
struct ShardSrc { const uint32_t* val; const uint8_t* heap; uint32_t log_size; uint32_t leaf; const uint8_t* upper; uint32_t block; };
__device__ __forceinline__ bool shard_locate(const ShardDecommitArgs& a, uint32_t tree /*0 = f, 1 + i = layer i*/, uint32_t layer,
                                             uint32_t lg, uint32_t leaf, ShardSrc& o)
{
    const uint32_t log_N = a.log_n + LOG_BLOWUP;
    if (tree == 0 || layer <= a.first_tail) {
        const uint32_t G = a.n_ranks, lgG = 31 - __clz(G), lb = a.log_block;
        const uint32_t b = leaf >> lb;
        if ((b & (G - 1)) != a.rank) return false;
        const uint32_t loc = ((b >> lgG) << lb) | (leaf & ((1u << lb) - 1u));
        o.val = a.vals[tree] + loc;
        o.heap = a.heap[tree]; o.log_size = lg - lgG; o.leaf = loc;
        o.upper = a.upper[tree]; o.block = b;
        return true;
    }
    if (a.rank != 0) return false;
    const size_t voff = layer_vals_off_dev(log_N, layer) - layer_vals_off_dev(log_N, a.first_tail);
    o.val = a.tail_vals + voff + leaf;
    o.heap = a.tail_heaps + 64 * voff; o.log_size = lg; o.leaf = leaf;
    o.upper = nullptr; o.block = 0;
    return true;
}
__global__ void __launch_bounds__(256) shard_decommit_kernel(const ShardDecommitArgs a)
{
    pdl_enter();
    const uint32_t k = a.log_n, L = k, log_N = k + LOG_BLOWUP, N = 1u << log_N, B = 1u << LOG_BLOWUP;
    uint8_t* tr = a.out;
    const uint32_t x = (uint32_t)((size_t)a.state->query_raw % (size_t)(N - 2 * B));  // prover.rs:263
    const size_t off0 = 84 + 36 * (size_t)L;
    const size_t rec0 = 12 + 32 * (size_t)log_N;
    if (threadIdx.x == 0 && a.flag_off && a.tail_vals) {
        // prover.rs:238: the last FRI polynomial has degree 0, i.e. its 8 evaluations are equal (the tail is replicated)
        const uint32_t* last = a.tail_vals + (layer_vals_off_dev(log_N, L) - layer_vals_off_dev(log_N, a.first_tail));
        uint32_t bad = 0;
        for (uint32_t j = 1; j < (N >> L); j++) bad |= (last[j] != last[0]);
        *reinterpret_cast<uint32_t*>(tr + a.flag_off) = bad;
    }
    for (uint32_t r = threadIdx.x; r < 4 + L; r += blockDim.x) {
        ShardSrc sp;
        if (r < 4) {                                             // prover.rs:266-277
            uint8_t* rec = tr + off0 + r * rec0;
            const uint32_t leaf = (r < 3) ? x + r * B : x;
            if (shard_locate(a, r < 3 ? 0 : 1, 0, log_N, leaf, sp)) *reinterpret_cast<uint32_t*>(rec) = *sp.val;
            if (a.rank == 0) *reinterpret_cast<uint32_t*>(rec + 4) = log_N;   // u64 LE length (high word stays 0)
        } else {                                                 // prover.rs:280-289
            const uint32_t i = r - 4, lg = log_N - i, len = 1u << lg;
            const uint32_t xi = x & (len - 1), nx = (xi + len / 2) & (len - 1);
            uint8_t* rec = tr + off0 + 4 * rec0 + 24 * (size_t)i + 64 * ((size_t)i * log_N - (size_t)i * (i - 1) / 2);
            if (shard_locate(a, 1 + i, i, lg, xi, sp)) *reinterpret_cast<uint32_t*>(rec) = *sp.val;
            if (shard_locate(a, 1 + i, i, lg, nx, sp)) *reinterpret_cast<uint32_t*>(rec + 4) = *sp.val;
            if (a.rank == 0) {
                *reinterpret_cast<uint32_t*>(rec + 8) = lg;
                *reinterpret_cast<uint32_t*>(rec + 16 + 32 * (size_t)lg) = lg;
            }
        }
    }
    const uint32_t per_rec_f = log_N;
    const uint32_t n_items = 4 * per_rec_f + 2 * (L * log_N - L * (L - 1) / 2);
    const uint32_t sub = threadIdx.x & 7;
    for (uint32_t it = threadIdx.x >> 3; it < n_items; it += blockDim.x >> 3) {
        uint8_t* dst;
        uint32_t lg, leaf, s, tree, layer;
        if (it < 4 * per_rec_f) {
            const uint32_t r = it / per_rec_f;
            s = it - r * per_rec_f;
            lg = log_N;
            leaf = (r < 3) ? x + r * B : x;
            tree = (r < 3) ? 0 : 1;
            layer = 0;
            dst = tr + off0 + r * rec0 + 12 + 32 * (size_t)s;
        } else {
            uint32_t rem = it - 4 * per_rec_f, i = 0;
            while (rem >= 2 * (log_N - i)) { rem -= 2 * (log_N - i); i++; }
            lg = log_N - i;
            const uint32_t len = 1u << lg, xi = x & (len - 1), nx = (xi + len / 2) & (len - 1);
            const uint32_t path = rem >= lg;
            s = path ? rem - lg : rem;
            leaf = path ? nx : xi;
            tree = 1 + i;
            layer = i;
            uint8_t* rec = tr + off0 + 4 * rec0 + 24 * (size_t)i + 64 * ((size_t)i * log_N - (size_t)i * (i - 1) / 2);
            dst = rec + (path ? 24 + 32 * (size_t)lg : 16) + 32 * (size_t)s;
        }
        ShardSrc sp;
        if (!shard_locate(a, tree, layer, lg, leaf, sp)) continue;
        const uint8_t* node;
        if (sp.upper == nullptr || s < a.log_block) node = sp.heap + 32 * sibling_node(sp.log_size, sp.leaf, s);
        else node = sp.upper + 32 * sibling_node(lg - a.log_block, sp.block, s - a.log_block);
        reinterpret_cast<uint32_t*>(dst)[sub] = reinterpret_cast<const uint32_t*>(node)[sub];
    }
}
cudaError_t launch_shard_decommit(cudaStream_t s, const ShardDecommitArgs& a)
{
    return zk_launch(shard_decommit_kernel, dim3(1), dim3(256), 0, s, a);
}

// channel.rs:19-26 over the decommitment records (prover.rs:274-288): 4 + L sequential commits, each
// state = SHA256(state || record) -- ~340 dependent compressions at k = 20, used for batches (the chains of all proofs
// advance in parallel; for a single proof the host does it after the copy).  One CTA per proof:
//   1. all threads stage the records in shared memory as big-endian message words;
//   2. only the FIRST block of a record contains the running state, so the message schedules K[t] + W[t] of every
//      other block (93 % of them) are computed up front, one block per thread, into shared memory;
//   3. thread 0 walks the chain: a full compression for the first block of a record, rounds only for the rest.
// The first version (one thread per proof, message words fetched from global memory block by block) took 0.9 ms.
__device__ __forceinline__ uint32_t tail_rec_len(uint32_t r, uint32_t log_N)    // bytes of record r
{
    return r < 4 ? 12 + 32 * log_N : 24 + 64 * (log_N - (r - 4));
}
// message word i of record r's commit (state || record, padded): i >= 8 only (the state words are handled by the caller)
__device__ __forceinline__ uint32_t tail_msg_word(const uint32_t* recw, uint32_t nw, uint32_t i, uint32_t last)
{
    if (i < nw) return recw[i - 8];
    if (i == nw) return 0x80000000u;
    if (i == last) return 32u * nw;  // message < 2^29 bytes: the high length word is 0
    return 0u;
}
template <int V>
__device__ __forceinline__ void compress_kw(uint32_t st[8], const uint32_t* kw)   // rounds only, K[t] + W[t] given
{
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    using sha::addv; using sha::bsig0; using sha::bsig1; using sha::ch; using sha::maj;
#pragma unroll 1
    for (int i = 0; i < 64; i += 16) {
        uint32_t kk[16];
#pragma unroll
        for (int q = 0; q < 16; q++) kk[q] = kw[i + q];
#define ZK_KW_T(j) (kk[j])
        ZK_SHA_16ROUNDS(V, ZK_KW_T)
#undef ZK_KW_T
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
__global__ void __launch_bounds__(256) chan_tail_kernel(ProofState* states, const uint8_t* transcripts, uint64_t stride,
                                                        uint32_t log_n)
{
    pdl_enter();
    extern __shared__ __align__(16) uint32_t tail_smem[];
    __shared__ uint32_t rec_word0[40], rec_blk0[40];            // first word / first precomputed block of record r
    const uint32_t L = log_n, log_N = log_n + LOG_BLOWUP, n_rec = 4 + L;
    ProofState* ps = states + blockIdx.x;
    const uint32_t* tr = reinterpret_cast<const uint32_t*>(transcripts + (size_t)blockIdx.x * stride + 84 + 36 * (size_t)L);
    if (threadIdx.x == 0) {
        uint32_t w = 0, bk = 0;
        for (uint32_t r = 0; r < n_rec; r++) {
            rec_word0[r] = w; rec_blk0[r] = bk;
            const uint32_t nw = 8 + tail_rec_len(r, log_N) / 4;
            w += nw - 8;
            bk += (nw + 3 + 15) / 16 - 1;                        // blocks after the first
        }
        rec_word0[n_rec] = w; rec_blk0[n_rec] = bk;
    }
    __syncthreads();
    const uint32_t n_words = rec_word0[n_rec], n_blk = rec_blk0[n_rec];
    uint32_t* recw = tail_smem;                                  // [n_words] big-endian message words of all records
    uint32_t* kw = tail_smem + ((n_words + 3) & ~3u);            // [n_blk][64]
    for (uint32_t i = threadIdx.x; i < n_words; i += blockDim.x) recw[i] = sha::bswap32(tr[i]);
    __syncthreads();
    for (uint32_t bk = threadIdx.x; bk < n_blk; bk += blockDim.x) {
        uint32_t r = 0;
        while (rec_blk0[r + 1] <= bk) r++;
        const uint32_t j = bk - rec_blk0[r] + 1;                 // block index inside the record's message (>= 1)
        const uint32_t nw = 8 + tail_rec_len(r, log_N) / 4, last = 16 * ((nw + 3 + 15) / 16) - 1;
        const uint32_t* rw = recw + rec_word0[r];
        uint32_t w[16];
#pragma unroll
        for (int q = 0; q < 16; q++) w[q] = tail_msg_word(rw, nw, 16 * j + q, last);
        uint32_t* o = kw + 64 * (size_t)bk;
#pragma unroll
        for (int q = 0; q < 16; q++) o[q] = w[q] + sha::c_k[q];
#pragma unroll 1
        for (int i = 16; i < 64; i += 16) {
#pragma unroll
            for (int q = 0; q < 16; q++) {
                w[q] = w[q] + sha::ssig1<0>(w[(q + 14) & 15]) + w[(q + 9) & 15] + sha::ssig0<0>(w[(q + 1) & 15]);
                o[i + q] = w[q] + sha::c_k[i + q];
            }
        }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    uint32_t st[8];
#pragma unroll
    for (int i = 0; i < 8; i++) st[i] = ps->chan_at_query[i];
    for (uint32_t r = 0; r < n_rec; r++) {
        const uint32_t nw = 8 + tail_rec_len(r, log_N) / 4, n_blocks = (nw + 3 + 15) / 16, last = 16 * n_blocks - 1;
        const uint32_t* rw = recw + rec_word0[r];
        uint32_t out[8], w[16];
        sha::init_iv(out);
#pragma unroll
        for (int q = 0; q < 8; q++) { w[q] = st[q]; w[8 + q] = tail_msg_word(rw, nw, 8 + q, last); }
        sha::compress<0>(out, w);
        const uint32_t* kb = kw + 64 * (size_t)rec_blk0[r];
        for (uint32_t j = 1; j < n_blocks; j++) compress_kw<64>(out, kb + 64 * (size_t)(j - 1));
#pragma unroll
        for (int i = 0; i < 8; i++) st[i] = out[i];
    }
#pragma unroll
    for (int i = 0; i < 8; i++) ps->chan[i] = st[i];
}
cudaError_t launch_chan_tail(cudaStream_t s, ProofState* states, const uint8_t* transcripts, uint64_t stride,
                             uint32_t log_n, uint32_t batch)
{
    const uint32_t log_N = log_n + LOG_BLOWUP;
    size_t words = 0, blocks = 0;
    for (uint32_t r = 0; r < 4 + log_n; r++) {
        const uint32_t len = r < 4 ? 12 + 32 * log_N : 24 + 64 * (log_N - (r - 4));
        const uint32_t nw = 8 + len / 4;
        words += nw - 8;
        blocks += (nw + 3 + 15) / 16 - 1;
    }
    const size_t smem = 4 * (((words + 3) & ~(size_t)3) + 64 * blocks);
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    {
        static bool attr_done[64] = {};
        cudaError_t e = set_max_smem_once(attr_done, chan_tail_kernel, 200 * 1024);
        if (e != cudaSuccess) return e;
    }
    return zk_launch(chan_tail_kernel, dim3(batch), dim3(256), smem, s, states, transcripts, stride, log_n);
}

// channel.rs:12-17 + public inputs: zero state, a[0], a[n-2]
__global__ void state_init_kernel(ProofState* st, const uint32_t* traces, uint64_t trace_stride,
                                  uint32_t log_n, uint32_t batch, ChanSeed seed)
{
    pdl_enter();
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    ProofState s;
    uint32_t* w = reinterpret_cast<uint32_t*>(&s);
    for (int i = 0; i < 64; i++) w[i] = 0;
    const uint32_t* tr = traces + (size_t)b * trace_stride;
    s.a_first = tr[0];
    s.a_last = tr[(1u << log_n) - 2];
    for (int i = 0; i < 8; i++) s.chan[i] = seed.w[i];    // Channel::new() is all zeros (channel.rs:12-17)
    st[b] = s;
}

// ---------------------------------------------------------------- launchers
static const TreeTune& tune_of(const TreeArgs& a)
{
    static const TreeTune dflt = default_tree_tune();
    return a.tune ? *a.tune : dflt;
}
template <int SRC>
static cudaError_t launch_pass(cudaStream_t s, const TreeArgs& a, uint32_t batch)
{
    uint32_t n_threads = 1u << (a.in_level - a.k);
    dim3 grid((n_threads + PASS_THREADS - 1) / PASS_THREADS, batch);
    return zk_launch(tree_pass_kernel<SRC>, grid, dim3(PASS_THREADS), 0, s, a);
}
// a.in_level inputs per proof -> root (+ channel step); in_level <= 2 * TILE_LOG
template <int SRC>
static cudaError_t launch_tile(cudaStream_t s, const TreeArgs& a, uint32_t batch)
{
    if (a.in_level > 2 * TILE_LOG) return cudaErrorInvalidValue;
    // Tile size: a lone warp hashes a node in ~2.7 us and already keeps its sub-partition's ALU pipe ~75 % busy
    // (one warp instruction per 2 cycles), so the latency-bound levels want FEW warps on MANY SMs and no two CTAs
    // on one SM (their surviving warps would share sub-partition 0): 2^tile_shift = 128 tiles <= 148 SMs,
    // tiles of 128 inputs (2 warps) for small trees, up to 2^TILE_LOG inputs above that.
    TreeArgs b = a;
    const TreeTune& T = tune_of(a);
    int lt = (int)a.in_level - (int)T.tile_shift;
    if (lt < (int)TILE_MIN_LOG) lt = TILE_MIN_LOG;
    if (lt > (int)TILE_LOG) lt = TILE_LOG;
    if (T.tile_log_override) lt = T.tile_log_override;
    if (lt > (int)a.in_level) lt = a.in_level;
    if (a.in_level - lt > TILE_LOG) return cudaErrorInvalidValue;
    b.log_tile = lt;
    const uint32_t n_tiles = 1u << (a.in_level - lt);
    if (n_tiles > 1 && !a.counters) return cudaErrorInvalidValue;
    uint32_t threads = lt >= 1 ? 1u << (lt - 1) : 1u;                       // pairs of a tile ...
    if (n_tiles > 2 * threads) threads = n_tiles / 2;                        // ... or of the tile roots, in the closing CTA
    if (threads < 32) threads = 32;
    // cooperative levels need two warps per 32 nodes
    uint32_t coop = T.coop_cnt > 32 * COOP_PAIRS ? 32 * COOP_PAIRS : T.coop_cnt;
    if (coop) {
        uint32_t need = 2 * (threads < coop ? threads : coop);
        need = (need + 63) & ~63u;
        if (need > threads) threads = need;
    }
    b.coop_cnt = coop;
    const uint32_t npair = coop ? (threads / 64 < COOP_PAIRS ? threads / 64 : COOP_PAIRS) : 0;
    const size_t smem = 4 * ((size_t)12 * threads + (size_t)COOP_WORDS * npair);
    {
        static bool attr_done[64] = {};
        cudaError_t e = set_max_smem_once(attr_done, tree_tile_kernel<SRC>, (int)(4 * (12 * TILE_PAIRS + COOP_WORDS * COOP_PAIRS)));
        if (e != cudaSuccess) return e;
    }
    return zk_launch(tree_tile_kernel<SRC>, dim3(n_tiles, batch), dim3(threads), smem, s, b);
}
static cudaError_t launch_tile_src(cudaStream_t s, int src, const TreeArgs& a, uint32_t batch)
{
    switch (src) {
        case SRC_VALUES: return launch_tile<SRC_VALUES>(s, a, batch);
        case SRC_FOLD: return launch_tile<SRC_FOLD>(s, a, batch);
        case SRC_CP: return a.cp.tab ? launch_tile<SRC_CP_TAB>(s, a, batch) : launch_tile<SRC_CP>(s, a, batch);
        case SRC_DIGESTS: return launch_tile<SRC_DIGESTS>(s, a, batch);
        default: return cudaErrorInvalidValue;
    }
}
// layers fri_layer .. last_layer of a proof (all single tiles: a.log_size <= TILE_MIN_LOG), one CTA per proof
cudaError_t launch_fri_tail(cudaStream_t s, TreeArgs a, uint32_t batch, uint32_t last_layer)
{
    if (a.log_size > TILE_MIN_LOG || a.log_size < 1) return cudaErrorInvalidValue;
    const TreeTune& T = tune_of(a);
    a.in_level = a.log_size;
    a.log_tile = a.log_size;
    uint32_t threads = 1u << (a.log_size - 1);
    if (threads < 32) threads = 32;
    uint32_t coop = T.coop_cnt > 32 * COOP_PAIRS ? 32 * COOP_PAIRS : T.coop_cnt;
    if (coop) {
        uint32_t need = 2 * (threads < coop ? threads : coop);
        need = (need + 63) & ~63u;
        if (need > threads) threads = need;
    }
    a.coop_cnt = coop;
    const uint32_t npair = coop ? (threads / 64 < COOP_PAIRS ? threads / 64 : COOP_PAIRS) : 0;
    const size_t smem = 4 * ((size_t)12 * threads + (size_t)COOP_WORDS * npair);
    {
        static bool attr_done[64] = {};
        cudaError_t e = set_max_smem_once(attr_done, fri_tail_kernel<>, (int)(4 * (12 * TILE_PAIRS + COOP_WORDS * COOP_PAIRS)));
        if (e != cudaSuccess) return e;
    }
    return zk_launch(fri_tail_kernel<>, dim3(1, batch), dim3(threads), smem, s, a, last_layer);
}
// CTA epilogue levels of a pass over 2^in_level inputs with 2^k per thread, never going above heap level `floor`
// (needs full CTAs: at least PASS_THREADS threads)
static uint32_t pick_ep(const TreeTune& T, uint32_t in_level, uint32_t k, uint32_t floor, uint32_t batch)
{
    uint32_t log_pt = 0;
    while ((1u << (log_pt + 1)) <= PASS_THREADS) log_pt++;
    if (!T.pass_epilogue || in_level < k + log_pt) return 0;
    const uint32_t room = in_level - k - floor;
    uint32_t ep = room < log_pt ? room : log_pt;
    // A CTA in its epilogue holds its registers and thread slots while one warp works: with a single proof that
    // costs more bulk throughput than the shorter upper-tree kernel wins back beyond one level (measured, 2^20:
    // 5.111 / 5.092 / 5.100 / 5.142 / 5.258 ms at 0 / 1 / 2 / 4 / 8 levels); in a batch the other proofs' CTAs fill the
    // gap and the deep epilogue wins (238.9 / 240.2 / 240.6 / 241.1 proofs/s at 0 / 2 / 4 / 8).
    const uint32_t cap = batch >= 4 ? T.pass_epilogue : (T.pass_epilogue < 1 ? T.pass_epilogue : 1u);
    return ep < cap ? ep : cap;
}
// digest passes until the tile kernel can take over, then the tile kernel
static cudaError_t finish_upper(cudaStream_t s, TreeArgs a, uint32_t batch, int* launches)
{
    cudaError_t e = cudaSuccess;
    while (a.in_level > 2 * TILE_LOG) {
        uint32_t rem = a.in_level - 2 * TILE_LOG;
        a.k = rem >= 4 ? 4 : rem;
        a.ep_levels = pick_ep(tune_of(a), a.in_level, a.k, 0, batch);
        e = launch_pass<SRC_DIGESTS>(s, a, batch);
        (*launches)++;
        if (e != cudaSuccess) return e;
        a.in_level -= a.k + a.ep_levels;
        a.ep_levels = 0;
    }
    e = launch_tile<SRC_DIGESTS>(s, a, batch);
    (*launches)++;
    return e;
}

static uint32_t ilog2(uint64_t v)
{
    uint32_t l = 0;
    while (((uint64_t)1 << (l + 1)) <= v) l++;
    return l;
}

cudaError_t launch_merkle_tree(cudaStream_t s, int src, TreeArgs a, uint32_t batch, int* n_launches)
{
    cudaError_t e = cudaSuccess;
    int launches = 0;
    a.in_level = a.log_size;
    const uint32_t lb = ilog2(batch);
    const TreeTune& T = tune_of(a);
    if (a.log_size + lb <= T.tile_direct_log || a.log_size <= TILE_LOG) {
        // small and mid-size trees: leaves -> root in one launch
        e = launch_tile_src(s, src, a, batch);
        launches++;
    } else {
        // leaf pass: enough threads to fill 148 SMs several times over, else fewer levels per thread
        int k = (int)a.log_size + (int)lb - (int)T.pass_log_threads;
        if (k < 1) k = 1;
        if (k > (int)T.pass_max_k) k = T.pass_max_k;
        if (k > (int)(a.log_size - TILE_LOG)) k = a.log_size - TILE_LOG;
        a.k = k;
        a.ep_levels = pick_ep(tune_of(a), a.in_level, k, 0, batch);
        switch (src) {
            case SRC_VALUES: e = launch_pass<SRC_VALUES>(s, a, batch); break;
            case SRC_FOLD: e = launch_pass<SRC_FOLD>(s, a, batch); break;
            case SRC_CP: e = a.cp.tab ? launch_pass<SRC_CP_TAB>(s, a, batch) : launch_pass<SRC_CP>(s, a, batch); break;
            default: return cudaErrorInvalidValue;
        }
        launches++;
        if (e != cudaSuccess) return e;
        a.in_level -= k + a.ep_levels;
        a.ep_levels = 0;
        e = finish_upper(s, a, batch, &launches);
    }
    if (n_launches) *n_launches += launches;
    return e;
}

cudaError_t launch_merkle_partial(cudaStream_t s, int src, TreeArgs a, uint32_t batch, uint32_t stop_level, int* n_launches)
{
    if (a.log_size <= stop_level) return cudaErrorInvalidValue;
    cudaError_t e = cudaSuccess;
    int launches = 0;
    a.in_level = a.log_size;
    a.post = POST_NONE;
    const uint32_t lb = ilog2(batch);
    int k = (int)a.log_size + (int)lb - 18;
    if (k < 1) k = 1;
    if (k > 5) k = 5;
    if (k > (int)(a.log_size - stop_level)) k = a.log_size - stop_level;
    a.k = k;
    a.ep_levels = pick_ep(tune_of(a), a.in_level, k, stop_level, batch);
    switch (src) {
        case SRC_VALUES: e = launch_pass<SRC_VALUES>(s, a, batch); break;
        case SRC_FOLD: e = launch_pass<SRC_FOLD>(s, a, batch); break;
        case SRC_CP: e = a.cp.tab ? launch_pass<SRC_CP_TAB>(s, a, batch) : launch_pass<SRC_CP>(s, a, batch); break;
        default: return cudaErrorInvalidValue;
    }
    launches++;
    a.in_level -= k + a.ep_levels;
    while (e == cudaSuccess && a.in_level > stop_level) {
        uint32_t rem = a.in_level - stop_level;
        int kk = (int)a.in_level + (int)lb - 18;   // still plenty of subtrees: keep several levels per thread
        if (kk < 1) kk = 1;
        if (kk > 5) kk = 5;
        a.k = rem < (uint32_t)kk ? rem : (uint32_t)kk;
        a.ep_levels = pick_ep(tune_of(a), a.in_level, a.k, stop_level, batch);
        e = launch_pass<SRC_DIGESTS>(s, a, batch);
        launches++;
        a.in_level -= a.k + a.ep_levels;
    }
    if (n_launches) *n_launches += launches;
    return e;
}

cudaError_t launch_merkle_upper(cudaStream_t s, TreeArgs a, uint32_t batch, int* n_launches)
{
    int launches = 0;
    cudaError_t e = finish_upper(s, a, batch, &launches);
    if (n_launches) *n_launches += launches;
    return e;
}

TreeTune default_tree_tune()
{
    TreeTune t;
    if (const char* e = getenv("ZK_EPILOGUE_LEVELS")) t.pass_epilogue = (uint32_t)atoi(e);
    if (const char* e = getenv("ZK_PASS_MAX_K")) {
        const uint32_t k = (uint32_t)atoi(e);
        t.pass_max_k = k < 1 ? 1u : (k > MAX_PASS_K ? MAX_PASS_K : k);
    }
    return t;
}
bool pdl_enabled()
{
    static const bool on = [] { const char* e = getenv("ZK_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}
// bench knob: bits 0-7 direct threshold, 8-15 forced tile size (0 = adaptive), 16-23 tile shift, 24-31 leaf-pass
// log-threads (0 = keep the default)
void set_tile_direct_log(TreeTune& t, uint32_t v)
{
    t.tile_direct_log = v & 0xffu;
    t.tile_log_override = (v >> 8) & 0xffu;
    t.tile_shift = ((v >> 16) & 0xffu) ? ((v >> 16) & 0xffu) : ZK_TILE_SHIFT;
    t.pass_log_threads = (v >> 24) ? (v >> 24) : ZK_PASS_LOG_THREADS;
}

__global__ void state_set_kernel(ProofState* st, uint32_t a_first, uint32_t a_last)
{
    pdl_enter();
    uint32_t* w = reinterpret_cast<uint32_t*>(st);
    for (int i = 0; i < 64; i++) w[i] = 0;
    st->a_first = a_first;
    st->a_last = a_last;
}
cudaError_t launch_state_set(cudaStream_t s, ProofState* st, uint32_t a_first, uint32_t a_last)
{
    return zk_launch(state_set_kernel, dim3(1), dim3(1), 0, s, st, a_first, a_last);
}

cudaError_t launch_tree_leaf_pass(cudaStream_t s, TreeArgs a, uint32_t batch, int variant)
{
    uint32_t n_threads = 1u << (a.in_level - a.k);
    dim3 grid((n_threads + PASS_THREADS - 1) / PASS_THREADS, batch);
    switch (variant) {
#ifdef ZK_SHA_EXPERIMENTS
#define ZK_VCASE(v) \
        case v: tree_pass_kernel<SRC_VALUES, v, 0><<<grid, PASS_THREADS, 0, s>>>(a); break; \
        case 100 + v: tree_pass_kernel<SRC_VALUES, v, 1><<<grid, PASS_THREADS, 0, s>>>(a); break; \
        case 200 + v: tree_pass_kernel<SRC_VALUES, v, 2><<<grid, PASS_THREADS, 0, s>>>(a); break;
        ZK_VCASE(0) ZK_VCASE(1) ZK_VCASE(2) ZK_VCASE(3) ZK_VCASE(35) ZK_VCASE(131) ZK_VCASE(259) ZK_VCASE(387) ZK_VCASE(385) ZK_VCASE(129) ZK_VCASE(515) ZK_VCASE(4611)
#endif
        default: tree_pass_kernel<SRC_VALUES><<<grid, PASS_THREADS, 0, s>>>(a); break;
    }
    return cudaGetLastError();
}

// Latency probe (bench support): `warps` warps of one CTA each hash a chain of `reps` nodes; returns cycles.
template <int V>
__global__ void node_latency_kernel(uint32_t* out, uint32_t reps, long long* cycles)
{
    uint32_t l[8], r[8], d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { l[i] = threadIdx.x * 8 + i; r[i] = out[i]; }
    long long t0 = clock64();
    for (uint32_t k = 0; k < reps; k++) {
        sha::hash_node<V>(l, r, d);
#pragma unroll
        for (int i = 0; i < 8; i++) r[i] = d[i];
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cycles = t1 - t0;
    if (d[0] == 0x12345678u) out[0] = d[1];
}
cudaError_t launch_node_latency(cudaStream_t s, uint32_t* out, uint32_t reps, uint32_t warps, int variant, long long* cycles)
{
    switch (variant) {
#ifdef ZK_SHA_EXPERIMENTS
#define ZK_LCASE(v) case v: node_latency_kernel<v><<<1, 32 * warps, 0, s>>>(out, reps, cycles); break;
        ZK_LCASE(0) ZK_LCASE(1) ZK_LCASE(3) ZK_LCASE(7) ZK_LCASE(11) ZK_LCASE(15) ZK_LCASE(19) ZK_LCASE(23) ZK_LCASE(31) ZK_LCASE(28) ZK_LCASE(12)
#endif
        default: node_latency_kernel<ZK_SHA_V><<<1, 32 * warps, 0, s>>>(out, reps, cycles); break;
    }
    return cudaGetLastError();
}

// multipliers of the FMA-pipe forms in sha256.cuh (2^(32-k) for a right shift / rotate by k)
cudaError_t sha_init_constants()
{
    uint32_t h[16] = {0};
    h[sha::SC_ONE] = 1u;
    h[sha::SC_R6] = 1u << 26; h[sha::SC_R11] = 1u << 21; h[sha::SC_R2] = 1u << 30; h[sha::SC_R13] = 1u << 19;
    h[sha::SC_S3] = 1u << 29; h[sha::SC_S10] = 1u << 22; h[sha::SC_R18] = 1u << 14; h[sha::SC_R19] = 1u << 13;
    cudaError_t e = cudaMemcpyToSymbol(sha::c_sha, h, sizeof h);
    if (e != cudaSuccess) return e;
    constexpr sha::Consts C = sha::make_consts();
    return cudaMemcpyToSymbol(sha::c_kwpad, C.kw_pad, sizeof C.kw_pad);
}

cudaError_t launch_decommit(cudaStream_t s, uint32_t log_n, const uint32_t* f_eval, uint64_t f_stride,
                            const uint8_t* f_heap, uint64_t f_heap_stride, const uint32_t* layers,
                            uint64_t layers_stride, const uint8_t* layer_heaps, uint64_t layer_heaps_stride,
                            ProofState* state, uint8_t* transcript, uint64_t transcript_stride, uint32_t batch)
{
    DecommitArgs a{log_n, f_eval, f_stride, f_heap, f_heap_stride, layers, layers_stride,
                   layer_heaps, layer_heaps_stride, state, transcript, transcript_stride, nullptr, nullptr, 0};
    // 32 digests per CTA pass: enough CTAs that every sibling load of a proof is in flight at once
    const uint32_t log_N = log_n + LOG_BLOWUP, n_items = 4 * log_N + 2 * (log_n * log_N - log_n * (log_n - 1) / 2);
    uint32_t slices = (n_items + 31) / 32;
    if (batch > 16) slices = 1 + (slices - 1) / (batch / 16);        // large batches fill the machine on their own
    return zk_launch(decommit_kernel, dim3(slices, batch), dim3(256), 0, s, a);
}

// further openings of ONE proof (pointers already offset to it): n_q positions xs_dev, records to out_dev
cudaError_t launch_open_queries(cudaStream_t s, uint32_t log_n, const uint32_t* f_eval, const uint8_t* f_heap,
                                const uint32_t* layers, const uint8_t* layer_heaps, ProofState* state,
                                const uint32_t* xs_dev, uint32_t n_q, uint8_t* out_dev, uint64_t out_stride)
{
    DecommitArgs a{log_n, f_eval, 0, f_heap, 0, layers, 0, layer_heaps, 0, state, nullptr, 0, xs_dev, out_dev, out_stride};
    const uint32_t log_N = log_n + LOG_BLOWUP, n_items = 4 * log_N + 2 * (log_n * log_N - log_n * (log_n - 1) / 2);
    uint32_t slices = (n_items + 31) / 32;
    if (n_q > 16) slices = 1 + (slices - 1) / (n_q / 16);
    return zk_launch(decommit_kernel, dim3(slices, n_q), dim3(256), 0, s, a);
}

cudaError_t launch_state_init(cudaStream_t s, ProofState* st, const uint32_t* traces, uint64_t trace_stride,
                              uint32_t log_n, uint32_t batch, ChanSeed seed)
{
    return zk_launch(state_init_kernel, dim3((batch + 127) / 128), dim3(128), 0, s, st, traces, trace_stride, log_n, batch, seed);
}

cudaError_t launch_fold(cudaStream_t s, const uint32_t* cur, uint32_t* next, uint32_t log_half, FoldArgs f,
                        PowTab pow, uint32_t batch, uint64_t next_stride, const ProofState* st)
{
    uint32_t half = 1u << log_half;
    uint32_t work = half >= 4 ? half / 4 : half;
    dim3 grid((work + 255) / 256, batch);
    f.log_half = log_half;
    return zk_launch(fold_kernel, grid, dim3(256), 0, s, cur, next, f, pow, next_stride, st);
}

cudaError_t launch_cp(cudaStream_t s, uint32_t* cp_out, CpArgs c, PowTab pow, uint32_t batch, uint64_t out_stride,
                      const ProofState* st)
{
    if (c.log_len == 0) c.log_len = pow.log_N;
    uint32_t N = 1u << c.log_len;
    dim3 grid((N / 4 + 255) / 256, batch);
    return c.tab ? zk_launch(cp_kernel<true>, grid, dim3(256), 0, s, cp_out, c, pow, out_stride, st)
                 : zk_launch(cp_kernel<false>, grid, dim3(256), 0, s, cp_out, c, pow, out_stride, st);
}

}  // namespace zk
