// Measurement support for bench.py (zk_bench_kernel): times single hot-path kernels with CUDA
// events on the ctx stream.  Not on the proving path.
#include <cuda_runtime.h>

#include <vector>

#include "common.cuh"

namespace zk {

__global__ void fill_kernel(uint32_t* p, size_t n, uint32_t seed)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        uint64_t x = (i + 1) * 0x9E3779B97F4A7C15ull + seed;
        x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 32;
        p[i] = (uint32_t)(x % P);
    }
}

__global__ void flush_kernel(uint4* p, size_t n)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_uint4((uint32_t)i, 1u, 2u, 3u);
}

// K1 on its own (field.cuh): every device field operation on arbitrary u32 operands, reduced first as F::from(u32) does
// (field.rs:20-24).  out: 6 rows of n -- a + b, a - b, a * b, -a, a^-1 (0 for a = 0), a^b  (field.rs:96-211, 26-38).
__global__ void __launch_bounds__(256) field_ops_kernel(const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out)
{
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = reduce_u32(a[i]), y = reduce_u32(b[i]);
    const uint32_t xm = to_mont(x);
    out[i] = fadd(x, y);
    out[n + i] = fsub(x, y);
    out[2 * n + i] = mont_mul(xm, y);                       // (x R)(y) / R
    out[3 * n + i] = fneg(x);
    out[4 * n + i] = x ? from_mont(mont_inv(xm)) : 0u;
    out[5 * n + i] = from_mont(mont_pow(xm, (uint64_t)b[i]));   // the exponent is the raw word (Pow<u32>, field.rs:26-38)
}
cudaError_t launch_field_ops(cudaStream_t s, const uint32_t* a, const uint32_t* b, size_t n, uint32_t* out)
{
    field_ops_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, b, n, out);
    return cudaGetLastError();
}

// Butterfly microbenchmark (instruction roofline of the transforms): every thread runs radix-16 register rounds -- 32
// DIF butterflies  (u, v) -> (u + v, (u - v) w)  with p = 3 * 2^30 + 1 > 2^31 -- on 16 values it keeps in registers; no
// memory traffic.  mode 0: whole butterflies, 1: only the Montgomery products, 2: only the additions / subtractions.
constexpr int BF_ITERS = 512;
template <int MODE>
__global__ void __launch_bounds__(256) bfly_kernel(uint32_t* out, uint32_t seed)
{
    uint32_t x[16], tw[15];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = reduce_u32(seed * (i + 1) + threadIdx.x * 7919u + blockIdx.x);
#pragma unroll
    for (int i = 0; i < 15; i++) tw[i] = reduce_u32(seed ^ (0x9E3779B9u * (i + 3)));
#pragma unroll 1
    for (int it = 0; it < BF_ITERS; it++) {
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int half = 8 >> t, off = 16 - (16 >> t);
#pragma unroll
            for (int j = 0; j < half; j++) {
#pragma unroll
                for (int blk = 0; blk < 16 / (2 * half); blk++) {
                    const int i0 = blk * 2 * half + j, i1 = i0 + half;
                    const uint32_t u = x[i0], v = x[i1];
                    if (MODE == 0) { x[i0] = fadd(u, v); x[i1] = mont_mul(fsub(u, v), tw[off + j]); }
                    else if (MODE == 1) { x[i0] = v; x[i1] = mont_mul(u, tw[off + j]); }
                    else { x[i0] = fadd(u, v); x[i1] = fsub(u, v); }
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) acc ^= x[i];
    if (acc == 0x12345u) out[0] = acc;
}
// returns butterflies per launch
double bench_bfly(cudaStream_t s, uint32_t* out, int n_sm, int mode)
{
    const int blocks = n_sm * 8;
    if (mode == 1) bfly_kernel<1><<<blocks, 256, 0, s>>>(out, 12345u);
    else if (mode == 2) bfly_kernel<2><<<blocks, 256, 0, s>>>(out, 12345u);
    else bfly_kernel<0><<<blocks, 256, 0, s>>>(out, 12345u);
    return (double)blocks * 256 * BF_ITERS * 32;
}

// 4 independent chains of {SHF, LOP3, IADD3}: the instruction mix of SHA-256 on the ALU pipe.
constexpr int INT_ITERS = 4096;
__global__ void __launch_bounds__(256) intpipe_kernel(uint32_t* out, uint32_t seed)
{
    uint32_t a0 = seed + threadIdx.x, b0 = a0 * 3u, c0 = a0 ^ 0x55u;
    uint32_t a1 = a0 + 1, b1 = b0 + 1, c1 = c0 + 1, a2 = a0 + 2, b2 = b0 + 2, c2 = c0 + 2, a3 = a0 + 3, b3 = b0 + 3, c3 = c0 + 3;
#pragma unroll 8
    for (int i = 0; i < INT_ITERS; i++) {
        a0 = __funnelshift_r(a0, b0, 7);  b0 = b0 ^ c0 ^ a0;  c0 = c0 + a0 + b0;
        a1 = __funnelshift_r(a1, b1, 11); b1 = b1 ^ c1 ^ a1;  c1 = c1 + a1 + b1;
        a2 = __funnelshift_r(a2, b2, 13); b2 = b2 ^ c2 ^ a2;  c2 = c2 + a2 + b2;
        a3 = __funnelshift_r(a3, b3, 17); b3 = b3 ^ c3 ^ a3;  c3 = c3 + a3 + b3;
    }
    uint32_t r = c0 ^ c1 ^ c2 ^ c3 ^ a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
    if (r == 0x12345678u) out[0] = r;  // practically never: keeps the chains alive
}

// FMA-pipe integer throughput (IMAD with an opaque multiplier) and a 1:1 ALU/IMAD mix: tells whether the
// two pipes really issue independently on this part (they do on paper: B300_MICROARCH "Pipe rates").
__global__ void __launch_bounds__(256) imad_kernel(uint32_t* out, uint32_t seed, uint32_t m)
{
    uint32_t a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 ^ 0x55u, a3 = a0 + 7u;
    uint32_t b0 = a0 + 1, b1 = a1 + 1, b2 = a2 + 1, b3 = a3 + 1, c0 = 5, c1 = 6, c2 = 7, c3 = 8;
#pragma unroll 8
    for (int i = 0; i < INT_ITERS; i++) {
        a0 = a0 * m + b0; b0 = b0 * m + c0; c0 = c0 * m + a0;
        a1 = a1 * m + b1; b1 = b1 * m + c1; c1 = c1 * m + a1;
        a2 = a2 * m + b2; b2 = b2 * m + c2; c2 = c2 * m + a2;
        a3 = a3 * m + b3; b3 = b3 * m + c3; c3 = c3 * m + a3;
    }
    uint32_t r = c0 ^ c1 ^ c2 ^ c3 ^ a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
    if (r == 0x12345678u) out[0] = r;
}
__global__ void __launch_bounds__(256) mixed_kernel(uint32_t* out, uint32_t seed, uint32_t m)
{
    uint32_t a0 = seed + threadIdx.x, b0 = a0 * 3u, c0 = a0 ^ 0x55u;
    uint32_t a1 = a0 + 1, b1 = b0 + 1, c1 = c0 + 1, a2 = a0 + 2, b2 = b0 + 2, c2 = c0 + 2, a3 = a0 + 3, b3 = b0 + 3, c3 = c0 + 3;
    uint32_t x0 = 1, x1 = 2, x2 = 3, x3 = 4, y0 = 5, y1 = 6, y2 = 7, y3 = 8, z0 = 9, z1 = 10, z2 = 11, z3 = 12;
#pragma unroll 8
    for (int i = 0; i < INT_ITERS; i++) {
        a0 = __funnelshift_r(a0, b0, 7);  x0 = x0 * m + y0; b0 = b0 ^ c0 ^ a0;  y0 = y0 * m + z0; c0 = c0 + a0 + b0; z0 = z0 * m + x0;
        a1 = __funnelshift_r(a1, b1, 11); x1 = x1 * m + y1; b1 = b1 ^ c1 ^ a1;  y1 = y1 * m + z1; c1 = c1 + a1 + b1; z1 = z1 * m + x1;
        a2 = __funnelshift_r(a2, b2, 13); x2 = x2 * m + y2; b2 = b2 ^ c2 ^ a2;  y2 = y2 * m + z2; c2 = c2 + a2 + b2; z2 = z2 * m + x2;
        a3 = __funnelshift_r(a3, b3, 17); x3 = x3 * m + y3; b3 = b3 ^ c3 ^ a3;  y3 = y3 * m + z3; c3 = c3 + a3 + b3; z3 = z3 * m + x3;
    }
    uint32_t r = c0 ^ c1 ^ c2 ^ c3 ^ a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3 ^ x0 ^ x1 ^ x2 ^ x3 ^ y0 ^ y1 ^ y2 ^ y3 ^ z0 ^ z1 ^ z2 ^ z3;
    if (r == 0x12345678u) out[0] = r;
}

cudaError_t bench_fill(cudaStream_t s, uint32_t* p, size_t n, uint32_t seed)
{
    fill_kernel<<<1184, 256, 0, s>>>(p, n, seed);
    return cudaGetLastError();
}
cudaError_t bench_flush(cudaStream_t s, void* p, size_t bytes)
{
    flush_kernel<<<1184, 256, 0, s>>>((uint4*)p, bytes / 16);
    return cudaGetLastError();
}
// returns the number of integer thread-instructions one launch executes.  mode 0: ALU pipe only
// (SHF/LOP3/IADD3), 1: IMAD only, 2: 1:1 mix
double bench_intpipe(cudaStream_t s, uint32_t* out, int n_sm, int mode)
{
    int blocks = n_sm * 8;
    if (mode == 1) { imad_kernel<<<blocks, 256, 0, s>>>(out, 12345u, 3u + (uint32_t)(n_sm & 1)); return (double)blocks * 256.0 * INT_ITERS * 12.0; }
    if (mode == 2) { mixed_kernel<<<blocks, 256, 0, s>>>(out, 12345u, 3u + (uint32_t)(n_sm & 1)); return (double)blocks * 256.0 * INT_ITERS * 24.0; }
    intpipe_kernel<<<blocks, 256, 0, s>>>(out, 12345u);
    return (double)blocks * 256.0 * INT_ITERS * 12.0;
}

}  // namespace zk
