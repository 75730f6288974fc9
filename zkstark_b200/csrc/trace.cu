// prover.rs:32-39 for a BATCH of proofs: a[0] = 1, a[1] = secret, a[i] = a[i-2]^2 + a[i-1]^2.
// The recurrence is sequential per proof, so one thread runs one proof; a warp stages 32 x 32 values in shared
// memory and writes them transposed, so global stores are full 128-byte rows.  ~45 cycles per step: worth it for
// batches (1024 proofs of 2^16 in ~1.5 ms instead of 0.6 s of host time), not for one long trace.
#include "common.cuh"

namespace zk {

__global__ void __launch_bounds__(32) trace_batch_kernel(const uint32_t* secrets, uint32_t* traces, uint64_t stride,
                                                         uint32_t log_n, uint32_t batch)
{
    __shared__ uint32_t tile[32][33];
    const uint32_t lane = threadIdx.x, b0 = blockIdx.x * 32, b = b0 + lane;
    const uint32_t n1 = (1u << log_n) - 1;  // values per trace
    const uint32_t secret = b < batch ? secrets[b] % P : 0u;
    uint32_t x_m = to_mont(1u), y_m = to_mont(secret);  // a[i-2], a[i-1] in Montgomery form
    for (uint32_t i0 = 0; i0 < n1; i0 += 32) {
#pragma unroll 4
        for (uint32_t t = 0; t < 32; t++) {
            const uint32_t i = i0 + t;
            uint32_t v;
            if (i == 0) v = 1u;
            else if (i == 1) v = secret;
            else {
                const uint32_t z_m = fadd(mont_mul(x_m, x_m), mont_mul(y_m, y_m));
                x_m = y_m;
                y_m = z_m;
                v = from_mont(z_m);
            }
            tile[lane][t] = v;
        }
        __syncwarp();
        for (uint32_t r = 0; r < 32; r++) {
            const uint32_t i = i0 + lane;
            if (b0 + r < batch && i < n1) traces[(size_t)(b0 + r) * stride + i] = tile[r][lane];
        }
        __syncwarp();
    }
}

cudaError_t launch_trace_batch(cudaStream_t s, const uint32_t* secrets_dev, uint32_t* traces, uint64_t stride,
                               uint32_t log_n, uint32_t batch)
{
    trace_batch_kernel<<<(batch + 31) / 32, 32, 0, s>>>(secrets_dev, traces, stride, log_n, batch);
    return cudaGetLastError();
}

}  // namespace zk
