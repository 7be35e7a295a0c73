"""frieda_b200 -- FRIEDA's data-parallel commit path (pack -> circle-FFT LDE -> Merkle -> FRI
layers -> grind -> query decommitment) as hand-written sm_100a CUDA kernels behind a C ABI
(include/frieda_b200.h), with this Python host layer mirroring the reference crate's API
(src/lib.rs:31-43).  No CPU fallback."""
from .api import (Context, FriedaError, PcsConfig, Proof, ReferencePanic, commit, commit_and_generate_proof,
                  default_context, generate_proof, load_library, query_positions, verify, verify_proof)

__all__ = ["Context", "FriedaError", "PcsConfig", "Proof", "ReferencePanic", "commit", "commit_and_generate_proof",
           "default_context", "generate_proof", "load_library", "query_positions", "verify", "verify_proof"]
