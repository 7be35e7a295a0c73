//! Drop-in for the reference's src/commit.rs (UNCOMPILED, see ../README.md): same `Commitment` type and
//! `commit` signature, the body forwards to the CUDA library.
use crate::ffi;

pub type Commitment = [u8; 32];

pub fn commit(data: &[u8], log_blowup_factor: u32) -> Commitment {
    let mut root = [0u8; 32];
    let rc = unsafe { ffi::frieda_commit(crate::gpu::ctx(), data.as_ptr(), data.len(), log_blowup_factor, root.as_mut_ptr()) };
    // the reference panics exactly where the library returns FRIEDA_ERR_PANIC (-1)
    assert!(rc == 0, "frieda_b200: {}", crate::gpu::last_error());
    root
}
