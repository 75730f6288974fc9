// prover.rs:32-39 for a BATCH of proofs: a[0] = 1, a[1] = secret, a[i] = a[i-2]^2 + a[i-1]^2.
// The recurrence is one dependent squaring per step, so one thread runs one proof and a warp runs 32 of them; the warp
// stages 32 x 32 values in shared memory and writes them transposed, so global stores are full 128-byte rows.
// The loop carries (s, y) = (a[i-2]^2, a[i-1]) in Montgomery form: per step one squaring and one addition on the critical
// path (~35 cycles of dependent latency); the conversion of a[i] out of Montgomery form is off the path.
#include "common.cuh"
#include "hostair.hpp"

namespace zk {

// GEN: another program of the AIR family (hostair.hpp TraceAir) -- the loop carries a[i-2], a[i-1] and
// a[i-2]^2 and spends seven products per step instead of one.
template <bool GEN>
__global__ void __launch_bounds__(32) trace_batch_kernel(const uint32_t* secrets, uint32_t* traces, uint64_t stride,
                                                         uint32_t log_n, uint32_t batch, TraceAir air)
{
    __shared__ uint32_t tile[32][33];
    const uint32_t lane = threadIdx.x, b0 = blockIdx.x * 32, b = b0 + lane;
    const uint32_t n1 = (1u << log_n) - 1;  // values per trace
    const uint32_t secret = b < batch ? secrets[b] % P : 0u;
    // state before step i: s_m = a[i-2]^2, y_m = a[i-1] (Montgomery form); the first block starts at i = 2
    uint32_t y_m = to_mont(secret);
    uint32_t s_m = to_mont(1u);             // a[0]^2 = 1
    uint32_t x0_m = 0;                      // GEN: a[i-2]
    if (GEN) {
        x0_m = to_mont(air.a0);
        s_m = mont_mul(x0_m, x0_m);
    }
    for (uint32_t i0 = 0; i0 < n1; i0 += 32) {
#pragma unroll
        for (uint32_t t = 0; t < 32; t++) {
            uint32_t v;
            if (i0 == 0 && t == 0) v = GEN ? air.a0 : 1u;
            else if (i0 == 0 && t == 1) v = secret;
            else if (GEN) {
                const uint32_t yy = mont_mul(y_m, y_m);
                uint32_t z = fadd(mont_mul(yy, air.c20m), mont_mul(s_m, air.c02m));
                z = fadd(z, mont_mul(mont_mul(x0_m, y_m), air.c11m));
                z = fadd(z, fadd(mont_mul(y_m, air.c10m), mont_mul(x0_m, air.c01m)));
                z = fadd(z, air.c00m);
                x0_m = y_m;
                s_m = yy;
                y_m = z;
                v = from_mont(z);
            } else {
                const uint32_t yy = mont_mul(y_m, y_m);     // a[i-1]^2
                const uint32_t z = fadd(s_m, yy);           // a[i]
                s_m = yy;
                y_m = z;
                v = from_mont(z);
            }
            tile[lane][t] = v;
        }
        __syncwarp();
#pragma unroll 8
        for (uint32_t r = 0; r < 32; r++) {
            const uint32_t i = i0 + lane;
            if (b0 + r < batch && i < n1) traces[(size_t)(b0 + r) * stride + i] = tile[r][lane];
        }
        __syncwarp();
    }
}

cudaError_t launch_trace_batch(cudaStream_t s, const uint32_t* secrets_dev, uint32_t* traces, uint64_t stride,
                               uint32_t log_n, uint32_t batch, const TraceAir& air)
{
    if (air.generic) trace_batch_kernel<true><<<(batch + 31) / 32, 32, 0, s>>>(secrets_dev, traces, stride, log_n, batch, air);
    else trace_batch_kernel<false><<<(batch + 31) / 32, 32, 0, s>>>(secrets_dev, traces, stride, log_n, batch, air);
    return cudaGetLastError();
}

}  // namespace zk
