//! Process-wide device context and the conversions between `frieda::proof::Proof` and the C struct
//! (UNCOMPILED, see ../README.md).  Field names follow stwo @ 19d12d7:
//!   FriProof { first_layer, inner_layers, last_layer_poly }            (core/fri.rs)
//!   FriLayerProof { fri_witness, decommitment, commitment }            (core/fri.rs)
//!   MerkleDecommitment { hash_witness, column_witness }                (core/vcs/prover.rs)
//!   LinePoly::new(coeffs) / LinePoly::into_ordered_coefficients        (core/poly/line.rs; coeffs in storage order)
//!   Blake2sHash(pub [u8; 32])                                          (core/vcs/blake2_hash.rs)
//!   PcsConfig { pow_bits, fri_config: FriConfig { log_blowup_factor, log_last_layer_degree_bound, n_queries } }
use std::{ffi::CStr, ptr, sync::OnceLock};

use stwo_prover::core::{
    fields::{m31::M31, qm31::QM31},
    fri::{FriLayerProof, FriProof},
    pcs::PcsConfig,
    poly::line::LinePoly,
    vcs::{blake2_hash::Blake2sHash, blake2_merkle::Blake2sMerkleHasher, prover::MerkleDecommitment},
};

use crate::{ffi, proof::Proof};

struct CtxPtr(*mut ffi::frieda_ctx);
// one context per process, used under the mutex below: a frieda_ctx is not internally locked
unsafe impl Send for CtxPtr {}
unsafe impl Sync for CtxPtr {}
static CTX: OnceLock<CtxPtr> = OnceLock::new();
static CALL: std::sync::Mutex<()> = std::sync::Mutex::new(());

/// The process-wide context on device `FRIEDA_DEVICE` (default 0).  Panics when no CUDA device is usable:
/// there is no CPU fallback.
pub(crate) fn ctx() -> *mut ffi::frieda_ctx {
    CTX.get_or_init(|| {
        let device = std::env::var("FRIEDA_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
        let mut p = ptr::null_mut();
        let rc = unsafe { ffi::frieda_ctx_create(device, &mut p) };
        if rc != 0 {
            let msg = unsafe { CStr::from_ptr(ffi::frieda_last_error(ptr::null())) }.to_string_lossy().into_owned();
            panic!("frieda_b200: cannot create a CUDA context ({rc}): {msg}");
        }
        CtxPtr(p)
    })
    .0
}

/// Serialises calls on the shared context (the reference functions are pure and re-entrant; callers that want
/// concurrency create one context per thread through `ffi` directly).
pub(crate) fn lock() -> std::sync::MutexGuard<'static, ()> {
    CALL.lock().unwrap_or_else(|e| e.into_inner())
}

pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(ffi::frieda_last_error(ctx())) }.to_string_lossy().into_owned()
}

pub(crate) fn config_to_c(c: &PcsConfig) -> ffi::frieda_pcs_config {
    ffi::frieda_pcs_config {
        log_blowup_factor: c.fri_config.log_blowup_factor,
        log_last_layer_degree_bound: c.fri_config.log_last_layer_degree_bound,
        n_queries: c.fri_config.n_queries as u64,
        pow_bits: c.pow_bits,
    }
}

fn qm31_from_c(q: &ffi::frieda_qm31) -> QM31 {
    QM31::from_u32_unchecked(q.v[0], q.v[1], q.v[2], q.v[3])
}
fn qm31_to_c(q: &QM31) -> ffi::frieda_qm31 {
    // QM31(CM31(a, b), CM31(c, d)), each M31(pub u32)
    ffi::frieda_qm31 { v: [q.0 .0 .0, q.0 .1 .0, q.1 .0 .0, q.1 .1 .0] }
}

unsafe fn slice<'a, T>(p: *const T, n: u32) -> &'a [T] {
    if n == 0 { &[] } else { std::slice::from_raw_parts(p, n as usize) }
}

unsafe fn layer_from_c(l: &ffi::frieda_layer_proof) -> FriLayerProof<Blake2sMerkleHasher> {
    let hashes = slice(l.hash_witness, l.n_hash_witness * 32);
    FriLayerProof {
        fri_witness: slice(l.fri_witness, l.n_fri_witness).iter().map(qm31_from_c).collect(),
        decommitment: MerkleDecommitment {
            hash_witness: hashes.chunks_exact(32).map(|h| Blake2sHash(h.try_into().unwrap())).collect(),
            column_witness: slice(l.column_witness, l.n_column_witness).iter().map(|&x| M31(x)).collect(),
        },
        commitment: Blake2sHash(l.commitment),
    }
}

/// Copies a library-owned proof into the crate's `Proof` (the caller frees the C object afterwards).
pub(crate) unsafe fn proof_from_c(p: &ffi::frieda_proof, pcs_config: PcsConfig) -> Proof {
    Proof {
        proof: FriProof {
            first_layer: layer_from_c(&p.first_layer),
            inner_layers: slice(p.inner_layers, p.n_inner_layers).iter().map(|l| layer_from_c(l)).collect(),
            // storage (bit-reversed) order, exactly as FriProver::commit leaves it
            last_layer_poly: LinePoly::new(slice(p.last_layer_poly, p.n_last_layer_poly).iter().map(qm31_from_c).collect()),
        },
        proof_of_work: p.proof_of_work,
        pcs_config,
        log_size_bound: p.log_size_bound,
        evaluations: slice(p.evaluations, p.n_evaluations).iter().map(qm31_from_c).collect(),
    }
}

/// A C view of a `Proof`: owns flat copies of the witness vectors, `as_ptr()` is valid while it lives.
pub(crate) struct CProof {
    c: ffi::frieda_proof,
    _fri: Vec<Vec<ffi::frieda_qm31>>,
    _hash: Vec<Vec<u8>>,
    _colw: Vec<Vec<u32>>,
    _inner: Vec<ffi::frieda_layer_proof>,
    _last: Vec<ffi::frieda_qm31>,
    _evals: Vec<ffi::frieda_qm31>,
}
impl CProof {
    pub(crate) fn as_ptr(&self) -> *const ffi::frieda_proof {
        &self.c
    }
}

pub(crate) fn proof_to_c(proof: &Proof) -> CProof {
    let layers: Vec<&FriLayerProof<Blake2sMerkleHasher>> =
        std::iter::once(&proof.proof.first_layer).chain(proof.proof.inner_layers.iter()).collect();
    let mut fri: Vec<Vec<ffi::frieda_qm31>> = layers.iter().map(|l| l.fri_witness.iter().map(qm31_to_c).collect()).collect();
    let mut hash: Vec<Vec<u8>> =
        layers.iter().map(|l| l.decommitment.hash_witness.iter().flat_map(|h| h.0).collect()).collect();
    let mut colw: Vec<Vec<u32>> = layers.iter().map(|l| l.decommitment.column_witness.iter().map(|m| m.0).collect()).collect();
    let mut c_layers: Vec<ffi::frieda_layer_proof> = layers
        .iter()
        .enumerate()
        .map(|(i, l)| ffi::frieda_layer_proof {
            commitment: l.commitment.0,
            n_fri_witness: fri[i].len() as u32,
            fri_witness: fri[i].as_mut_ptr(),
            n_hash_witness: (hash[i].len() / 32) as u32,
            hash_witness: hash[i].as_mut_ptr(),
            n_column_witness: colw[i].len() as u32,
            column_witness: colw[i].as_mut_ptr(),
        })
        .collect();
    // LinePoly keeps its coefficients private in storage order; Deref<Target = [SecureField]> exposes them
    let mut last: Vec<ffi::frieda_qm31> = proof.proof.last_layer_poly.iter().map(qm31_to_c).collect();
    let mut evals: Vec<ffi::frieda_qm31> = proof.evaluations.iter().map(qm31_to_c).collect();
    let first = c_layers.remove(0);
    let c = ffi::frieda_proof {
        first_layer: first,
        n_inner_layers: c_layers.len() as u32,
        inner_layers: c_layers.as_mut_ptr(),
        n_last_layer_poly: last.len() as u32,
        last_layer_poly: last.as_mut_ptr(),
        proof_of_work: proof.proof_of_work,
        pcs_config: config_to_c(&proof.pcs_config),
        log_size_bound: proof.log_size_bound,
        n_evaluations: evals.len() as u32,
        evaluations: evals.as_mut_ptr(),
    };
    CProof { c, _fri: fri, _hash: hash, _colw: colw, _inner: c_layers, _last: last, _evals: evals }
}

/// `frieda_proof_serialize_bincode` of a proof (test support for the wire-format check in lib.rs).
#[cfg(test)]
pub(crate) fn serialize_bincode_via_library(proof: &Proof) -> Vec<u8> {
    let c = proof_to_c(proof);
    let n = unsafe { ffi::frieda_proof_serialize_bincode(c.as_ptr(), ptr::null_mut(), 0) };
    let mut out = vec![0u8; n];
    unsafe { ffi::frieda_proof_serialize_bincode(c.as_ptr(), out.as_mut_ptr(), n) };
    out
}
