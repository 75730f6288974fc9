//! Level-1 drop-in for src/prover.rs: `generate_proof(channel) -> Proof` with the reference signature
//! (prover.rs:9), forwarded whole to the device-resident pipeline.  `Proof { state, data }` is byte-identical,
//! so the untouched `proof.verify()` and `proof.size()` (main.rs:28,35) work on it.
//! Use: in main.rs replace `mod prover;` by `#[path = "prover_gpu.rs"] mod prover;`.
use crate::channel::Channel;
use crate::gpu;
use crate::proof::Proof;

pub fn generate_proof(channel: Channel) -> Proof {
    // Channel::new() is the only constructor (channel.rs:12-17) and main.rs hands it over untouched (main.rs:19-23);
    // a channel somebody has already committed to cannot be continued on the device (patch: channel.rs `is_untouched`)
    assert!(channel.is_untouched(), "generate_proof expects a fresh Channel::new()");
    let log_n = 10u32; // the reference's literals: 1023-point trace, 1024 subgroup, 8192 coset (prover.rs:32,56-57)
    let secret = 3141592u32; // prover.rs:34
    let (state, data) = gpu::prove_fibsq(log_n, secret);
    Proof::new(state, data.into_boxed_slice())
}
