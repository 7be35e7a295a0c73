//! extern "C" declarations of include/frieda_b200.h (UNCOMPILED, see ../README.md).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int};

#[repr(C)]
#[derive(Clone, Copy)]
pub struct frieda_qm31 {
    pub v: [u32; 4],
}
#[repr(C)]
#[derive(Clone, Copy)]
pub struct frieda_pcs_config {
    pub log_blowup_factor: u32,
    pub log_last_layer_degree_bound: u32,
    pub n_queries: u64,
    pub pow_bits: u32,
}
#[repr(C)]
pub struct frieda_layer_proof {
    pub commitment: [u8; 32],
    pub n_fri_witness: u32,
    pub fri_witness: *mut frieda_qm31,
    pub n_hash_witness: u32,
    pub hash_witness: *mut u8,
    pub n_column_witness: u32,
    pub column_witness: *mut u32,
}
#[repr(C)]
pub struct frieda_proof {
    pub first_layer: frieda_layer_proof,
    pub n_inner_layers: u32,
    pub inner_layers: *mut frieda_layer_proof,
    pub n_last_layer_poly: u32,
    pub last_layer_poly: *mut frieda_qm31,
    pub proof_of_work: u64,
    pub pcs_config: frieda_pcs_config,
    pub log_size_bound: u32,
    pub n_evaluations: u32,
    pub evaluations: *mut frieda_qm31,
}
pub enum frieda_ctx {}

extern "C" {
    pub fn frieda_ctx_create(device: c_int, out: *mut *mut frieda_ctx) -> c_int;
    pub fn frieda_ctx_destroy(ctx: *mut frieda_ctx);
    pub fn frieda_last_error(ctx: *const frieda_ctx) -> *const c_char;
    pub fn frieda_commit(ctx: *mut frieda_ctx, data: *const u8, len: usize, log_blowup: u32, root_out: *mut u8) -> c_int;
    pub fn frieda_commit_batch(ctx: *mut frieda_ctx, blobs: *const u8, blob_len: usize, blob_stride: usize, n: usize,
                               log_blowup: u32, roots_out: *mut u8) -> c_int;
    pub fn frieda_prove(ctx: *mut frieda_ctx, data: *const u8, len: usize, seed_or_null: *const u64,
                        cfg: *const frieda_pcs_config, root_out: *mut u8, proof_out: *mut *mut frieda_proof) -> c_int;
    pub fn frieda_prove_batch(ctx: *mut frieda_ctx, blobs: *const u8, blob_len: usize, blob_stride: usize, n: usize,
                              seeds: *const u64, cfg: *const frieda_pcs_config, roots_out: *mut u8,
                              proofs_out: *mut *mut frieda_proof) -> c_int;
    pub fn frieda_verify(proof: *const frieda_proof, seed_or_null: *const u64) -> c_int;
    pub fn frieda_verify_batch(ctx: *mut frieda_ctx, proofs: *const *const frieda_proof, n: usize,
                               seeds_or_null: *const u64, results: *mut c_int) -> c_int;
    pub fn frieda_proof_free(proof: *mut frieda_proof);
    // wire formats: the library's flat encoding and the bincode-1.x layout of the serde-derived `Proof`
    pub fn frieda_proof_serialize(proof: *const frieda_proof, out: *mut u8, cap: usize) -> usize;
    pub fn frieda_proof_deserialize(bytes: *const u8, len: usize, proof_out: *mut *mut frieda_proof) -> c_int;
    pub fn frieda_proof_serialize_bincode(proof: *const frieda_proof, out: *mut u8, cap: usize) -> usize;
    // beyond the reference API: the positions `proof.evaluations` belong to (count, or 0 / negative)
    pub fn frieda_proof_query_positions(proof: *const frieda_proof, seed_or_null: *const u64, positions_out: *mut u32,
                                        cap: usize) -> i64;
}
