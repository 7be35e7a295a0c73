//! Drop-in for the reference's src/commit.rs (UNCOMPILED, see ../README.md): same `Commitment` type and
//! `commit` signature, the body forwards to the CUDA library.
use crate::ffi;

pub type Commitment = [u8; 32];

pub fn commit(data: &[u8], log_blowup_factor: u32) -> Commitment {
    let mut root = [0u8; 32];
    let _guard = crate::gpu::lock();
    let rc = unsafe { ffi::frieda_commit(crate::gpu::ctx(), data.as_ptr(), data.len(), log_blowup_factor, root.as_mut_ptr()) };
    // the reference panics exactly where the library returns FRIEDA_ERR_PANIC (-1)
    assert!(rc == 0, "frieda_b200: {}", crate::gpu::last_error());
    root
}

/// n blobs of equal length in one device pass (no reference counterpart).
pub fn commit_batch(blobs: &[&[u8]], log_blowup_factor: u32) -> Vec<Commitment> {
    let Some(first) = blobs.first() else { return Vec::new() };
    let len = first.len();
    assert!(blobs.iter().all(|b| b.len() == len), "commit_batch: blobs must have equal length");
    let flat: Vec<u8> = blobs.concat();
    let mut roots = vec![[0u8; 32]; blobs.len()];
    let _guard = crate::gpu::lock();
    let rc = unsafe {
        ffi::frieda_commit_batch(crate::gpu::ctx(), flat.as_ptr(), len, len, blobs.len(), log_blowup_factor,
                                 roots.as_mut_ptr() as *mut u8)
    };
    assert!(rc == 0, "frieda_b200: {}", crate::gpu::last_error());
    roots
}

#[cfg(test)]
mod tests {
    use super::*;

    // the reference's src/commit.rs:28-38, unchanged
    #[test]
    fn test_commit() {
        let data = include_bytes!("../../tests/golden/blob");
        let commitment = commit(data, 4);
        assert_eq!(
            commitment,
            [209, 162, 213, 6, 157, 197, 135, 229, 93, 194, 156, 198, 37, 90, 249, 55, 255, 127, 237, 14, 228, 27,
             223, 90, 249, 135, 23, 249, 215, 79, 96, 232]
        );
    }
}
