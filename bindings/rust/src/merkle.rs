//! Drop-in for the reference's src/merkle.rs: same type, same constructor, same `trace`/`Index`.
use crate::ffi;
use std::ops::Index;

#[derive(Clone, Debug)]
pub struct Merkle(Box<[Hash]>);
pub type Hash = [u8; 32];
pub type AuthPath = Box<[Hash]>;

impl Merkle {
    /// merkle.rs:14 -- Merkle::new(size, data): forwarded to zk_merkle_build.
    pub fn new(size: usize, data: impl Iterator<Item = u32>) -> Self {
        assert!(size.is_power_of_two()); // merkle.rs:17-21
        let leaves: Vec<u32> = data.collect();
        assert_eq!(leaves.len(), size);
        let mut heap = vec![Hash::default(); 2 * size - 1];
        ffi::with_ctx(|ctx| unsafe {
            let rc = ffi::zk_merkle_build(ctx, leaves.as_ptr(), size as u64, heap.as_mut_ptr() as *mut u8);
            ffi::check(ctx, rc);
        });
        Self(heap.into_boxed_slice())
    }

    /// merkle.rs:54-71, unchanged (host side: log2(size) copies).
    pub fn trace(&self, mut i: usize) -> AuthPath {
        i += self.0.len() / 2;
        let mut v = vec![];
        while i != 0 {
            if i % 2 == 0 { v.push(self[i - 1]); i -= 2; } else { v.push(self[i + 1]); i -= 1; }
            i >>= 1;
        }
        v.into_boxed_slice()
    }
}

impl Index<usize> for Merkle {
    type Output = Hash;
    fn index(&self, i: usize) -> &Self::Output { &self.0[i] }
}
