//! FRIEDA with the commit path on a B200 (UNCOMPILED, see ../README.md).
//!
//! The public surface is the reference's: `frieda::api::{commit, generate_proof, verify}`, the `commit` and `proof`
//! modules and the `M31` re-export (reference src/lib.rs:14-44).  `utils` is gone: packing happens on the device.
//! `gpu` and `ffi` are the new, private plumbing.

/// Re-export of stwo-prover's M31 field for arithmetic operations
pub use stwo_prover::core::fields::m31::M31;

pub mod commit;
mod ffi;
mod gpu;
pub mod proof;

/// Core public API for FRIEDA
pub mod api {
    use stwo_prover::core::pcs::PcsConfig;

    use crate::{commit::Commitment, proof::Proof};

    use super::*;

    /// Commit to data using FRI protocol
    pub fn commit(data: &[u8], log_blowup_factor: u32) -> Commitment {
        commit::commit(data, log_blowup_factor)
    }

    /// Generate a FRI proof for committed data
    pub fn generate_proof(data: &[u8], seed: Option<u64>, pcs_config: PcsConfig) -> Proof {
        proof::generate_proof(data, seed, pcs_config)
    }

    /// Verify a FRI proof against a commitment
    pub fn verify(proof: Proof, seed: Option<u64>) -> bool {
        proof::verify_proof(proof, seed)
    }

    /// Batched forms (no reference counterpart): n blobs of equal length, one device pass.
    pub fn commit_batch(blobs: &[&[u8]], log_blowup_factor: u32) -> Vec<Commitment> {
        commit::commit_batch(blobs, log_blowup_factor)
    }
}

#[cfg(test)]
mod tests {
    use stwo_prover::core::{fri::FriConfig, pcs::PcsConfig};

    use super::*;

    // the reference's src/lib.rs:52-85, unchanged
    #[test]
    fn test_end_to_end() {
        let original_data = b"This is the original data that needs to be made available.";
        let commitment = api::commit(original_data, 4);
        let pcs_config = PcsConfig {
            fri_config: FriConfig { log_blowup_factor: 4, log_last_layer_degree_bound: 0, n_queries: 20 },
            pow_bits: 20,
        };
        let proof = api::generate_proof(original_data, None, pcs_config);
        assert_eq!(proof.proof.first_layer.commitment.0, commitment);
        assert!(api::verify(proof, None));
    }

    // wire format: what libfrieda_b200's frieda_proof_serialize_bincode writes is bincode's encoding of `Proof`
    #[test]
    fn test_bincode_layout_matches_the_library() {
        let data = b"This is the original data that needs to be made available.";
        let pcs_config = PcsConfig {
            fri_config: FriConfig { log_blowup_factor: 4, log_last_layer_degree_bound: 0, n_queries: 20 },
            pow_bits: 8,
        };
        let proof = api::generate_proof(data, Some(7), pcs_config);
        let ours = bincode::serialize(&proof).unwrap();
        let theirs = gpu::serialize_bincode_via_library(&proof);
        assert_eq!(ours, theirs);
        let back: proof::Proof = bincode::deserialize(&theirs).unwrap();
        assert!(api::verify(back, Some(7)));
    }
}
