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
#include "common.cuh"

namespace zk {

struct NttPassArgs {
    const uint32_t* src; uint64_t src_stride;   // per proof (elements)
    uint32_t* dst; uint64_t dst_stride;
    uint32_t log_M;      // transform length
    uint32_t log_R;      // radix of this pass
    uint32_t log_S;      // stride of this pass (strided kernel)
    uint32_t log_Msub;   // log_R + log_S
    uint32_t n_cosets;   // 1 (iNTT) or 8 (LDE): coset r lives at + r*M in the work buffer
    int coset_scale;     // load c_k (5 h^r)^k from the single coefficient array `src`
    int inverse;
    int use_scale;
    uint32_t scale_m;    // Montgomery constant multiplied into the final outputs (n^-1)
    const uint32_t* tw;  // w_R^j (or inverse), j < R/2, Montgomery form
    PowTab hpow;         // powers of h, order N = 2^hpow.log_N  (M divides N)
    PowTab wpow;         // powers of 5, exponents < 2^wpow.log_N
    uint32_t log_R1;     // final pass: first-pass radix (0 when single pass)
    uint32_t log_Q2;     // final pass: second-pass radix (0 unless three passes)
    uint32_t coset_base; // index of the first coset of this launch (sharded LDE: one coset per launch)
};

constexpr int NTT_MAX_THREADS = 1024;  // block size is chosen per radix: 32 * clamp(R/16, 2, 32)
constexpr int NTT_UNR = 4;             // independent global loads in flight per thread
constexpr int TP = 33;  // padded tile pitch

__device__ __forceinline__ uint32_t pw(const PowTab& t, uint32_t e)
{
    return mont_mul(__ldg(t.lo + (e & ((1u << t.lo_bits) - 1u))), __ldg(t.hi + (e >> t.lo_bits)));
}

__device__ __forceinline__ uint32_t brev(uint32_t x, uint32_t bits) { return bits ? (__brev(x) >> (32 - bits)) : 0u; }

// One register-blocked round of the in-place DIF: stages s .. s+LR-1 of the size-R transform along
// the row index of sm[R][TP].  A thread owns one column (its lane) of one radix-2^LR group: it pulls
// the 2^LR rows into registers, runs LR butterfly stages there, and writes them back -- one shared
// memory round trip per LR stages.  Twiddles are warp-uniform (they depend on the row group only):
// stage s' butterfly with in-group position j uses w_R^((j*stride + lo) << s'), a broadcast LDS.
template <int LR>
__device__ __forceinline__ void dif_round(uint32_t* sm, const uint32_t* tws, uint32_t log_R, uint32_t s)
{
    constexpr int RAD = 1 << LR;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t log_stride = log_R - s - LR;
    const uint32_t n_mini = 1u << (log_R - LR);
    for (uint32_t q = warp; q < n_mini; q += nw) {
        const uint32_t lo = q & ((1u << log_stride) - 1u), hi = q >> log_stride;
        const uint32_t base = (hi << (log_R - s)) + lo;
        uint32_t x[RAD];
#pragma unroll
        for (int d = 0; d < RAD; d++) x[d] = sm[(base + ((uint32_t)d << log_stride)) * TP + lane];
#pragma unroll
        for (int t = 0; t < LR; t++) {
            const int half = RAD >> (t + 1);
#pragma unroll
            for (int j = 0; j < half; j++) {
                const uint32_t tw = tws[(((uint32_t)j << log_stride) + lo) << (s + t)];
#pragma unroll
                for (int blk = 0; blk < RAD / (2 * half); blk++) {
                    const int i0 = blk * 2 * half + j, i1 = i0 + half;
                    const uint32_t u = x[i0], v = x[i1];
                    x[i0] = fadd(u, v);
                    x[i1] = mont_mul(fsub(u, v), tw);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < RAD; d++) sm[(base + ((uint32_t)d << log_stride)) * TP + lane] = x[d];
    }
    __syncthreads();
}

// In-place DIF over the row index of sm[R][TP], all 32 columns; leaves X[c] at row brev(c).
__device__ __forceinline__ void cta_dif(uint32_t* sm, const uint32_t* tws, uint32_t log_R)
{
    uint32_t s = 0;
    while (s < log_R) {
        const uint32_t rem = log_R - s;
        const uint32_t r = rem > 4 ? (rem == 5 ? 3u : 4u) : rem;
        switch (r) {
            case 4: dif_round<4>(sm, tws, log_R, s); break;
            case 3: dif_round<3>(sm, tws, log_R, s); break;
            case 2: dif_round<2>(sm, tws, log_R, s); break;
            default: dif_round<1>(sm, tws, log_R, s); break;
        }
        s += r;
    }
}

// c_k * 5^k * h^(r k): coefficient scaling that moves the evaluation onto coset r (prover.rs:69)
__device__ __forceinline__ uint32_t coset_scale(const NttPassArgs& a, uint32_t c, uint32_t k, uint32_t r)
{
    uint32_t v = mont_mul(c, pw(a.wpow, k));
    r += a.coset_base;
    if (r) v = mont_mul(v, pw(a.hpow, (uint32_t)(((uint64_t)r * k) & ((1u << a.hpow.log_N) - 1u))));
    return v;
}

__global__ void __launch_bounds__(NTT_MAX_THREADS) ntt_strided_kernel(const NttPassArgs a)
{
    extern __shared__ uint32_t smem[];
    const uint32_t R = 1u << a.log_R, S = 1u << a.log_S;
    uint32_t* sm = smem;
    uint32_t* tws = smem + R * TP;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t tiles_per_blk = S >> 5, n_blk = 1u << (a.log_M - a.log_Msub);
    uint32_t tile = blockIdx.x;
    const uint32_t per_array = tiles_per_blk * n_blk;
    const uint32_t r = tile / per_array;  // coset
    tile -= r * per_array;
    const uint32_t blk = tile / tiles_per_blk, b = ((tile % tiles_per_blk) << 5) + lane;
    const size_t M = (size_t)1 << a.log_M;
    const uint32_t* src = a.src + (size_t)blockIdx.y * a.src_stride + (a.coset_scale ? 0 : r * M);
    uint32_t* dst = a.dst + (size_t)blockIdx.y * a.dst_stride + r * M;
    const size_t base = ((size_t)blk << a.log_Msub) + b;

    for (uint32_t i = threadIdx.x; i < R / 2; i += blockDim.x) tws[i] = a.tw[i];
    // load (rows of 32 consecutive words = 128 B), NTT_UNR independent loads in flight per thread
    for (uint32_t row0 = warp * NTT_UNR; row0 < R; row0 += nw * NTT_UNR) {
        uint32_t v[NTT_UNR];
#pragma unroll
        for (int u = 0; u < NTT_UNR; u++) v[u] = src[base + ((size_t)(row0 + u) << a.log_S)];
#pragma unroll
        for (int u = 0; u < NTT_UNR; u++) {
            if (a.coset_scale) v[u] = coset_scale(a, v[u], (uint32_t)(base + ((size_t)(row0 + u) << a.log_S)), r);
            sm[(row0 + u) * TP + lane] = v[u];
        }
    }
    __syncthreads();
    cta_dif(sm, tws, a.log_R);
    // twiddle w_Msub^(b c) and store (c,b)
    const uint32_t N1 = (1u << a.hpow.log_N) - 1u, sh = a.hpow.log_N - a.log_Msub;
    for (uint32_t c0 = warp * NTT_UNR; c0 < R; c0 += nw * NTT_UNR) {
        uint32_t v[NTT_UNR], t[NTT_UNR];
#pragma unroll
        for (int u = 0; u < NTT_UNR; u++) {
            const uint32_t c = c0 + u;
            uint32_t e = (b * c) << sh;
            if (a.inverse) e = ((N1 + 1u) - e) & N1;
            t[u] = pw(a.hpow, e);
            v[u] = sm[brev(c, a.log_R) * TP + lane];
        }
#pragma unroll
        for (int u = 0; u < NTT_UNR; u++)
            dst[((size_t)blk << a.log_Msub) + ((size_t)(c0 + u) << a.log_S) + b] = mont_mul(v[u], t[u]);
    }
}

// Final pass.  Lane l of the tile owns one contiguous row of R elements; lanes are arranged so
// that for every output digit c the 32 lanes write 32 consecutive words:
//   lane = u * CS + r   (r = coset, u = index along the first-pass digit c1)
//   out[(rho_nat + (M/R) c) * CS + r],  rho_nat = c1 + R1 c2
__global__ void __launch_bounds__(NTT_MAX_THREADS) ntt_final_kernel(const NttPassArgs a)
{
    extern __shared__ uint32_t smem[];
    const uint32_t R = 1u << a.log_R;
    uint32_t* sm = smem;
    uint32_t* tws = smem + R * TP;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const uint32_t CS = a.n_cosets, R1 = 1u << a.log_R1, Q2 = 1u << a.log_Q2;
    uint32_t U = 32 / CS;           // lanes along c1
    if (U > R1) U = R1;
    const uint32_t tiles_c1 = R1 / U;
    const uint32_t c2 = blockIdx.x / tiles_c1, c1_0 = (blockIdx.x % tiles_c1) * U;
    const size_t M = (size_t)1 << a.log_M;
    const uint32_t* srcb = a.src + (size_t)blockIdx.y * a.src_stride;
    uint32_t* dstb = a.dst + (size_t)blockIdx.y * a.dst_stride;

    for (uint32_t i = threadIdx.x; i < R / 2; i += blockDim.x) tws[i] = a.tw[i];
    // load: the tile is n_rows contiguous rows of R words; flat index f = l * R + a0, a warp reads 32
    // consecutive words of one row per access (R >= 32), NTT_UNR accesses in flight
    const uint32_t n_rows = U * CS, total = n_rows << a.log_R;
    for (uint32_t f0 = threadIdx.x; f0 < total; f0 += blockDim.x * NTT_UNR) {
        uint32_t v[NTT_UNR];
#pragma unroll
        for (int u = 0; u < NTT_UNR; u++) {
            const uint32_t f = f0 + u * blockDim.x;
            if (f < total) {
                const uint32_t l = f >> a.log_R, a0 = f & (R - 1);
                const uint32_t rr = l % CS, uu = l / CS;
                const size_t q = (size_t)(c1_0 + uu) * Q2 + c2;
                v[u] = srcb[(a.coset_scale ? 0 : rr * M) + (q << a.log_R) + a0];
            }
        }
#pragma unroll
        for (int u = 0; u < NTT_UNR; u++) {
            const uint32_t f = f0 + u * blockDim.x;
            if (f < total) {
                const uint32_t l = f >> a.log_R, a0 = f & (R - 1);
                if (a.coset_scale) {
                    const uint32_t rr = l % CS, uu = l / CS;
                    const size_t q = (size_t)(c1_0 + uu) * Q2 + c2;
                    v[u] = coset_scale(a, v[u], (uint32_t)((q << a.log_R) + a0), rr);
                }
                sm[a0 * TP + l] = v[u];
            }
        }
    }
    __syncthreads();
    cta_dif(sm, tws, a.log_R);
    if (lane < n_rows) {
        const uint32_t r = lane % CS, u = lane / CS;
        const size_t rho = (size_t)(c1_0 + u) + (size_t)R1 * c2;
        const uint32_t log_rows = a.log_M - a.log_R;
        for (uint32_t c = warp; c < R; c += nw) {
            uint32_t v = sm[brev(c, a.log_R) * TP + lane];
            if (a.use_scale) v = mont_mul(v, a.scale_m);
            dstb[(rho + ((size_t)c << log_rows)) * CS + r] = v;
        }
    }
}

// After the size-n inverse transform of (a_0..a_{n-2}, 0): c'_k.  The interpolant of the n-1
// trace points (prover.rs:60: zip truncates to 1023 pairs) has degree <= n-2, i.e. is
//   c_k = c'_k - t g^(k+1),  t = c'_{n-1}
// (equivalently: the transform of the trace extended by the phantom value -g sum a_i g^i).
__global__ void phantom_fix_kernel(uint32_t* c, uint64_t stride, uint32_t log_n, PowTab hpow)
{
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
    c[(size_t)blockIdx.x * stride + (1u << log_n) - 1] = 0u;
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
    int cap = (int)(mu + log_arrays) - 5 - NTT_LOG_TILES;
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
static inline uint32_t ntt_threads(uint32_t log_R)
{
    uint32_t nw = (1u << log_R) >> 4;
    if (nw < 2) nw = 2;
    if (nw > 32) nw = 32;
    return nw * 32;
}

// tw tables: for radix 2^m the R/2 entries start at offset 2^(m-1) - 1 (m >= 1)
static inline size_t tw_off(uint32_t m) { return m ? ((size_t)1 << (m - 1)) - 1 : 0; }

struct NttTables {
    const uint32_t* tw_fwd;
    const uint32_t* tw_inv;
    PowTab hpow, wpow;
};

static cudaError_t set_smem(const void* fn, size_t bytes)
{
    if (bytes > 48 * 1024) return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return cudaSuccess;
}

// Transform `n_cosets` arrays of length 2^mu.  Passes before the last run in place on `work`
// (n_cosets * M words per proof); the last writes natural order (interleaved over cosets) to dst.
// With coset_scale the first pass reads the single coefficient array `coeffs` instead of `work`.
static cudaError_t run_ntt(cudaStream_t s, const NttTables& T, uint32_t mu, uint32_t n_cosets, bool coset_scale,
                           bool inverse, const uint32_t* coeffs, uint64_t coeffs_stride, uint32_t* work,
                           uint64_t work_stride, uint32_t* dst, uint64_t dst_stride, bool use_scale,
                           uint32_t scale_m, uint32_t batch, int* n_launches, uint32_t coset_base = 0)
{
    uint32_t log_arrays = 0;
    while ((1u << log_arrays) < n_cosets * batch) log_arrays++;
    NttPlan p = make_plan(mu, log_arrays);
    NttPassArgs a{};
    a.log_M = mu;
    a.n_cosets = n_cosets;
    a.inverse = inverse;
    a.hpow = T.hpow;
    a.wpow = T.wpow;
    a.coset_base = coset_base;
    const uint32_t* twb = inverse ? T.tw_inv : T.tw_fwd;
    uint32_t log_S = mu;
    cudaError_t e;
    for (uint32_t i = 0; i + 1 < p.n_pass; i++) {
        a.log_R = p.log_r[i];
        a.log_Msub = log_S;
        log_S -= a.log_R;
        a.log_S = log_S;
        a.tw = twb + tw_off(a.log_R);
        a.coset_scale = (i == 0 && coset_scale);
        a.src = (i == 0) ? coeffs : work;
        a.src_stride = (i == 0) ? coeffs_stride : work_stride;
        a.dst = work;
        a.dst_stride = work_stride;
        size_t smem = ((size_t)(1u << a.log_R) * TP + (1u << a.log_R) / 2) * 4;
        if ((e = set_smem((const void*)ntt_strided_kernel, smem)) != cudaSuccess) return e;
        uint32_t tiles = (1u << (mu - 5)) >> a.log_R;  // (M / (R*S)) * (S/32)
        dim3 grid(tiles * n_cosets, batch);
        ntt_strided_kernel<<<grid, ntt_threads(a.log_R), smem, s>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    a.log_R = p.log_r[p.n_pass - 1];
    a.log_Msub = a.log_R;
    a.log_S = 0;
    a.tw = twb + tw_off(a.log_R);
    a.coset_scale = (p.n_pass == 1 && coset_scale);
    a.src = (p.n_pass == 1) ? coeffs : work;
    a.src_stride = (p.n_pass == 1) ? coeffs_stride : work_stride;
    a.dst = dst;
    a.dst_stride = dst_stride;
    a.use_scale = use_scale;
    a.scale_m = scale_m;
    a.log_R1 = p.n_pass >= 2 ? p.log_r[0] : 0;
    a.log_Q2 = p.n_pass == 3 ? p.log_r[1] : 0;
    {
        size_t smem = ((size_t)(1u << a.log_R) * TP + (1u << a.log_R) / 2 + 1) * 4;
        if ((e = set_smem((const void*)ntt_final_kernel, smem)) != cudaSuccess) return e;
        uint32_t R1 = 1u << a.log_R1, U = 32 / n_cosets;
        if (U > R1) U = R1;
        uint32_t tiles = (R1 / U) << a.log_Q2;
        dim3 grid(tiles, batch);
        ntt_final_kernel<<<grid, ntt_threads(a.log_R), smem, s>>>(a);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        if (n_launches) (*n_launches)++;
    }
    return cudaSuccess;
}

// lagrange() over the subgroup (polynomial.rs:337; prover.rs:60-61).  y: n words per proof with
// y[n-1] ignored when n_points == n-1.  coeffs may alias y (in place).
cudaError_t launch_interpolate(cudaStream_t s, const uint32_t* tw_fwd, const uint32_t* tw_inv, PowTab hpow,
                               PowTab wpow, uint32_t log_n, bool phantom, uint32_t* y, uint64_t y_stride,
                               uint32_t* coeffs, uint64_t coeffs_stride, uint32_t batch, int* n_launches)
{
    NttTables T{tw_fwd, tw_inv, hpow, wpow};
    cudaError_t e;
    const uint32_t n = 1u << log_n;
    if (phantom) {
        phantom_clear_kernel<<<batch, 1, 0, s>>>(y, y_stride, log_n);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
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
        phantom_fix_kernel<<<grid, th, 0, s>>>(coeffs, coeffs_stride, log_n, hpow);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
        phantom_clear_kernel<<<batch, 1, 0, s>>>(coeffs, coeffs_stride, log_n);
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
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
                                 int* n_launches)
{
    NttTables T{tw_fwd, tw_inv, hpow, wpow};
    return run_ntt(s, T, log_n, 1, true, false, coeffs, 0, work, 0, out, 0, false, 0, 1, n_launches, r);
}

}  // namespace zk
