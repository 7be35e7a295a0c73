//! Drop-in for the reference's src/proof.rs (UNCOMPILED, see ../README.md): `Proof`, `generate_proof`,
//! `commit_and_generate_proof`, `verify_proof` keep their signatures.
use serde::{Deserialize, Serialize};
use stwo_prover::core::{
    fields::qm31::QM31, fri::FriProof, pcs::PcsConfig, vcs::blake2_merkle::Blake2sMerkleHasher,
};

use crate::{commit::Commitment, ffi, gpu};

#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct Proof {
    pub proof: FriProof<Blake2sMerkleHasher>,
    pub proof_of_work: u64,
    pub pcs_config: PcsConfig,
    pub log_size_bound: u32,
    pub evaluations: Vec<QM31>,
}

pub fn generate_proof(data: &[u8], seed: Option<u64>, pcs_config: PcsConfig) -> Proof {
    commit_and_generate_proof(data, seed, pcs_config).1
}

pub fn commit_and_generate_proof(data: &[u8], seed: Option<u64>, pcs_config: PcsConfig) -> (Commitment, Proof) {
    let cfg = gpu::config_to_c(&pcs_config);
    let mut root = [0u8; 32];
    let mut p: *mut ffi::frieda_proof = std::ptr::null_mut();
    let seed_ptr = seed.as_ref().map_or(std::ptr::null(), |s| s as *const u64);
    let _guard = gpu::lock();
    let rc = unsafe {
        ffi::frieda_prove(gpu::ctx(), data.as_ptr(), data.len(), seed_ptr, &cfg, root.as_mut_ptr(), &mut p)
    };
    assert!(rc == 0, "frieda_b200: {}", gpu::last_error());
    // field-by-field copy into FriProof { first_layer, inner_layers, last_layer_poly } etc.
    let proof = unsafe { gpu::proof_from_c(&*p, pcs_config) };
    unsafe { ffi::frieda_proof_free(p) };
    (root, proof)
}

pub fn verify_proof(proof: Proof, seed: Option<u64>) -> bool {
    let c = gpu::proof_to_c(&proof); // borrows the Vecs of `proof`
    let seed_ptr = seed.as_ref().map_or(std::ptr::null(), |s| s as *const u64);
    match unsafe { ffi::frieda_verify(c.as_ptr(), seed_ptr) } {
        1 => true,
        0 => false,
        -1 => panic!("called `Option::unwrap()` on a `None` value"), // reference: src/proof.rs:166-173
        e => panic!("frieda_b200 error {e}"),
    }
}

// The reference's `mod tests` (src/proof.rs:103-194: generate / commit_and_generate / verify, the five tamper cases,
// the `#[should_panic]` short-evaluations case and the seed test) is kept verbatim by the maintainer; it compiles
// against this file unchanged because `Proof` and the three signatures are unchanged.  The same behaviours are
// exercised against the library in tests/test_gpu_parity.py and tests/cpp/test_frieda_api.cpp.
