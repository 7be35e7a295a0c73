"""CPU-side checks of the product library: it loads, exports every symbol include/frieda_b200.h
declares, refuses to compute without a GPU (no CPU fallback), and its host-only verifier and proof
codec agree with the oracle on oracle-generated proofs.  No device compute here."""
import os
import re

import pytest

import frieda_b200 as F
from frieda_b200 import api
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = (1 << 31) - 1


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "frieda_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(frieda_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations parsed"
    lib = F.load_library()
    for name in declared:
        assert hasattr(lib, name), f"libfrieda_b200.so does not export {name}"
    assert sorted(api.EXPORTS) == declared


def test_product_does_not_reference_the_oracle():
    # the oracle is test infrastructure: nothing under frieda_b200/ may import, link or execute it
    for dirpath, _, files in os.walk(os.path.join(ROOT, "frieda_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "frieda_oracle" not in text and "from oracle" not in text and "import oracle" not in text, f


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(F.FriedaError) as ei:
        F.Context(0)
    assert ei.value.code == api.ERR_CUDA
    with pytest.raises(F.FriedaError):
        F.commit(b"abc", 4)


def test_n_inner_layers_shape():
    lib = F.load_library()
    import ctypes as C
    cfg = F.PcsConfig(4, 0, 20, 20)
    assert lib.frieda_fri_n_inner_layers(131072, C.byref(cfg)) == 13
    assert lib.frieda_fri_n_inner_layers(262146, C.byref(cfg)) == 14      # C1, log_last 0
    cfg1 = F.PcsConfig(4, 1, 20, 20)
    assert lib.frieda_fri_n_inner_layers(262146, C.byref(cfg1)) == 13     # C1, log_last 1
    assert lib.frieda_fri_n_inner_layers(2, C.byref(cfg)) == api.ERR_PANIC  # reference panics


@pytest.fixture(scope="module")
def oracle_proof(blob_bytes):
    cfg = O.make_config(4, 1, 20, 20)
    _, pr = O.prove(blob_bytes, None, cfg)
    return pr


def test_proof_codec_roundtrip(oracle_proof):
    b = oracle_proof.serialize()
    p = F.Proof.deserialize(b)
    assert p.serialize() == b
    assert p.clone().serialize() == b
    assert p.n_inner_layers == 13 and p.log_size_bound == 15
    with pytest.raises(F.FriedaError):
        F.Proof.deserialize(b[:-1])
    with pytest.raises(F.FriedaError):
        F.Proof.deserialize(b"XXXX" + b[4:])


def test_host_verifier_matches_reference_test_suite(oracle_proof):
    p = F.Proof.deserialize(oracle_proof.serialize())
    assert F.verify_proof(p, None)                               # src/proof.rs:136-141
    q = p.clone()
    q.proof_of_work += 1
    assert not F.verify_proof(q, None)                           # :143-149
    ev = p.evaluations
    q = p.clone()
    q.set_evaluation(0, [(x + 1) % P for x in ev[0]])
    assert not F.verify_proof(q, None)                           # :151-157
    q = p.clone()
    for i, e in enumerate(reversed(ev)):
        q.set_evaluation(i, e)
    assert not F.verify_proof(q, None)                           # :158-164
    q = p.clone()
    q.set_evaluation(0, ev[1])
    q.set_evaluation(1, ev[0])
    assert not F.verify_proof(q, None)                           # :175-181
    q = p.clone()
    q.pop_evaluation()
    with pytest.raises(F.ReferencePanic):                        # :166-173
        F.verify_proof(q, None)
    q.c.n_evaluations += 1
    assert not F.verify_proof(p, 7)                              # wrong seed


def test_host_verifier_seed_binding(blob_bytes):
    cfg = O.make_config(4, 1, 20, 20)
    _, p1 = O.prove(blob_bytes, 1, cfg)
    _, p2 = O.prove(blob_bytes, 2, cfg)
    f1, f2 = F.Proof.deserialize(p1.serialize()), F.Proof.deserialize(p2.serialize())
    assert F.verify_proof(f1, 1) and F.verify_proof(f2, 2)       # src/proof.rs:183-193
    assert not F.verify_proof(f1, 2) and not F.verify_proof(f2, 1)


def test_host_verifier_rejects_witness_tampering(oracle_proof):
    p = F.Proof.deserialize(oracle_proof.serialize())
    q = p.clone()
    q.c.first_layer.hash_witness[5] ^= 1
    assert not F.verify_proof(q, None)
    q = p.clone()
    q.c.inner_layers[3].fri_witness[0].v[2] ^= 1
    assert not F.verify_proof(q, None)
    q = p.clone()
    q.c.inner_layers[0].commitment[0] ^= 1
    assert not F.verify_proof(q, None)
    q = p.clone()
    q.c.last_layer_poly[0].v[0] ^= 1
    assert not F.verify_proof(q, None)
    q = p.clone()
    q.c.n_inner_layers -= 1
    assert not F.verify_proof(q, None)
    q.c.n_inner_layers += 1


@pytest.mark.parametrize("case", ["pattern_1024_seedlen", "e2e_string", "pattern_3000_b2_l2"])
def test_host_verifier_on_other_shapes(case, golden):
    g = next(x for x in golden["oracle_generated"]["prove"] if x["name"] == case)
    if case == "e2e_string":
        data = b"This is the original data that needs to be made available."
    else:
        data = bytes(i % 256 for i in range(g["len"]))
    cfg = O.make_config(*g["cfg"])
    _, pr = O.prove(data, g["seed"], cfg)
    p = F.Proof.deserialize(pr.serialize())
    assert F.verify_proof(p, g["seed"])
    assert not F.verify_proof(p, (g["seed"] or 0) + 1)


def test_bincode_layout_shape(oracle_proof):
    # structure check of the serde/bincode-1.x encoding (it cannot be validated against Rust here):
    # total length = sum of the field encodings, first u64 = first layer's fri_witness count
    import struct
    p = F.Proof.deserialize(oracle_proof.serialize())
    b = p.serialize_bincode()
    c = p.c
    def layer_len(l):
        return 8 + 16 * l.n_fri_witness + 8 + 32 * l.n_hash_witness + 8 + 4 * l.n_column_witness + 32
    want = layer_len(c.first_layer) + 8 + sum(layer_len(c.inner_layers[i]) for i in range(c.n_inner_layers))
    want += 8 + 16 * c.n_last_layer_poly + 4 + 8 + 4 + 4 + 4 + 8 + 4 + 8 + 16 * c.n_evaluations
    assert len(b) == want
    assert struct.unpack_from("<Q", b, 0)[0] == c.first_layer.n_fri_witness
    assert struct.unpack_from("<Q", b, len(b) - 8 - 16 * c.n_evaluations)[0] == c.n_evaluations


# ---- the GPU batch verifier's core, executed on the host (same __host__ __device__ code as the kernels) ----
def _core(p, seed):
    return api.verify_core_host(p, seed)


def test_batch_verifier_core_matches_reference_behaviours(oracle_proof):
    p = F.Proof.deserialize(oracle_proof.serialize())
    assert _core(p, None) == 1
    assert _core(p, 7) == 0
    q = p.clone()
    q.proof_of_work += 1
    assert _core(q, None) == 0
    ev = p.evaluations
    q = p.clone()
    q.set_evaluation(0, [(x + 1) % P for x in ev[0]])
    assert _core(q, None) == 0
    q = p.clone()
    for i, e in enumerate(reversed(ev)):
        q.set_evaluation(i, e)
    assert _core(q, None) == 0
    q = p.clone()
    q.set_evaluation(0, ev[1])
    q.set_evaluation(1, ev[0])
    assert _core(q, None) == 0
    q = p.clone()
    q.pop_evaluation()
    assert _core(q, None) == api.ERR_PANIC                       # src/proof.rs:166-173
    q.c.n_evaluations += 1


def test_batch_verifier_core_agrees_with_host_verifier_on_tampering(oracle_proof):
    p = F.Proof.deserialize(oracle_proof.serialize())

    def both(q, seed=None):
        try:
            want = int(F.verify_proof(q, seed))
        except F.ReferencePanic:
            want = api.ERR_PANIC
        assert _core(q, seed) == want
        return want

    assert both(p) == 1
    q = p.clone(); q.c.first_layer.hash_witness[5] ^= 1; assert both(q) == 0
    q = p.clone(); q.c.inner_layers[3].fri_witness[0].v[2] ^= 1; assert both(q) == 0
    q = p.clone(); q.c.inner_layers[0].commitment[0] ^= 1; assert both(q) == 0
    q = p.clone(); q.c.last_layer_poly[0].v[0] ^= 1; assert both(q) == 0
    q = p.clone(); q.c.inner_layers[12].hash_witness[0] ^= 1; assert both(q) == 0
    q = p.clone(); q.c.n_inner_layers -= 1; assert both(q) == 0; q.c.n_inner_layers += 1
    q = p.clone(); q.c.first_layer.n_hash_witness -= 1; assert both(q) == 0; q.c.first_layer.n_hash_witness += 1
    q = p.clone(); q.c.inner_layers[2].n_fri_witness -= 1; assert both(q) == 0; q.c.inner_layers[2].n_fri_witness += 1
    # Merkle failure in layer 0 together with short evaluations: the rebuild panics first
    q = p.clone(); q.c.first_layer.hash_witness[0] ^= 1; q.pop_evaluation(); assert both(q) == api.ERR_PANIC
    q.c.n_evaluations += 1


@pytest.mark.parametrize("case", ["pattern_1024_seedlen", "e2e_string", "pattern_3000_b2_l2"])
def test_batch_verifier_core_other_shapes(case, golden):
    g = next(x for x in golden["oracle_generated"]["prove"] if x["name"] == case)
    data = b"This is the original data that needs to be made available." if case == "e2e_string" else bytes(
        i % 256 for i in range(g["len"]))
    _, pr = O.prove(data, g["seed"], O.make_config(*g["cfg"]))
    p = F.Proof.deserialize(pr.serialize())
    assert _core(p, g["seed"]) == 1
    assert _core(p, (g["seed"] or 0) + 1) == 0


def test_peer_entry_points_reject_a_null_context():
    # the split-commit entry points over peer memory (include/frieda_b200.h) validate before touching a device
    import ctypes as C
    L = F.load_library()
    ptrs = (C.c_void_p * 2)(None, None)
    out = (C.c_uint8 * 32)()
    assert L.frieda_commit_split_local_peers(None, ptrs, 2, 16, 32, 2, 0, None) == api.ERR_ARG
    assert L.frieda_merkle_combine_peers(None, ptrs, 2, out) == api.ERR_ARG
    assert L.frieda_commit_split_peers(None, None, 0, 2, 0, 2, ptrs, 16, ptrs, ptrs, 1, out) == api.ERR_ARG


def test_query_positions_match_the_oracle_transcript(blob_bytes):
    # frieda_proof_query_positions replays the verifier's transcript: same positions as the prover drew (oracle trace)
    for data, seed, cfg in ((blob_bytes, None, (4, 1, 20, 20)), (bytes(i % 256 for i in range(5000)), 9, (3, 0, 33, 6))):
        t = O.trace(data, seed, O.make_config(*cfg), with_trees=False)
        p = F.Proof.deserialize(t.proof_bytes)
        pos = F.query_positions(p, seed)
        assert pos == sorted(set(int(x) for x in t.queries)) and len(pos) == len(p.evaluations)
        assert F.query_positions(p, 12345) != pos                  # another seed, another transcript
        q = p.clone()
        q.proof_of_work += 1
        assert F.query_positions(q, seed) == []                    # proof of work rejected before the queries
