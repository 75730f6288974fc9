// The program being proven, host view: a one-column AIR over a three-row window,
//   a[0] = a0, a[1] = secret, a[i+2] = c20 a[i+1]^2 + c02 a[i]^2 + c11 a[i] a[i+1] + c10 a[i+1] + c01 a[i] + c00,
// with the reference's two boundary constraints (first value, value n - 2; prover.rs:101-113).  generic == false is
// the reference's own program (prover.rs:32-39, :134-138: a0 = 1, c20 = c02 = 1), which keeps its specialised code.
#pragma once
#include <stdint.h>

namespace zk {
struct HostAir {
    uint32_t a0 = 1, c20 = 1, c02 = 1, c11 = 0, c10 = 0, c01 = 0, c00 = 0;   // canonical residues
    bool generic = false;
};
// the same program for the device trace generator (trace.cu): a0 canonical, the coefficients in Montgomery form
struct TraceAir { uint32_t generic, a0, c20m, c02m, c11m, c10m, c01m, c00m; };
}  // namespace zk
