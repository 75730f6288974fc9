//! Drop-in for the reference's src/prover.rs: `generate_proof(channel) -> Proof`, forwarded whole to the
//! device-resident pipeline (zk_prove_fibsq).  Channel::new() starts from the zero state (channel.rs:12-17),
//! which is the only state the reference ever constructs.
use crate::channel::Channel;
use crate::ffi;
use crate::proof::Proof;

pub fn generate_proof(channel: Channel) -> Proof {
    assert!(channel.is_fresh(), "generate_proof expects Channel::new()");
    let log_n = 10u32; // the reference's literals: 1023-point trace, 1024 subgroup, 8192 coset
    let secret = 3141592u32; // prover.rs:34
    let cap = unsafe { ffi::zk_transcript_len(log_n) };
    let mut data = vec![0u8; cap];
    let mut state = [0u8; 32];
    let mut len = 0usize;
    ffi::with_ctx(|ctx| unsafe {
        let rc = ffi::zk_prove_fibsq(ctx, log_n, secret, data.as_mut_ptr(), cap, &mut len, state.as_mut_ptr());
        ffi::check(ctx, rc);
    });
    data.truncate(len);
    Proof::new(state, data.into_boxed_slice())
}
