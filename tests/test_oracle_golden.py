"""Pins the CPU oracle against every vector the reference's own tests hold for this path
 plus the survey's independent probe values.  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import oracle as O

P = (1 << 31) - 1


def pattern(n):
    return bytes(i % 256 for i in range(n))


def test_commit_blob_golden_root(blob_bytes, golden):
    # src/commit.rs:28-38 -- the reference's only known-answer test.
    assert O.commit(blob_bytes, 4).hex() == golden["reference_certified"]["commit_blob_blowup4"]


def test_bytes_to_one_felt():
    # src/utils.rs:40-49
    for i in range(256):
        f = O.bytes_to_felts(bytes([i]))
        assert len(f) == 1 and int(f[0]) == i


def test_bytes_to_two_felt():
    # src/utils.rs:51-66: 60 bits -> 8 bytes -> 3 felts [i, i, 0]
    for i in range(513):
        bits = i | (i << 30)
        f = O.bytes_to_felts(bits.to_bytes(8, "little"))
        assert [int(x) for x in f] == [i, i, 0]


def test_packing_matches_bigint_slicing():
    rng = np.random.default_rng(1)
    for n in (0, 1, 3, 4, 15, 16, 29, 30, 31, 100, 1000, 4097):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        big = int.from_bytes(data, "little")
        cnt = (8 * n + 29) // 30
        want = [(big >> (30 * k)) & ((1 << 30) - 1) for k in range(cnt)]
        assert [int(x) for x in O.bytes_to_felts(data)] == want


def test_poly_log_matches_float_formula():
    import math
    for n in list(range(0, 300)) + [1024, 4096, 131072, 262146, 1 << 20, (1 << 20) + 1]:
        cnt = (8 * n + 29) // 30
        lg = max(math.ceil(math.log2(cnt)) if cnt else 0, 2)  # src/utils.rs:23
        assert O.poly_log(n) == lg - 2


def test_survey_probe_commit_vectors(golden):
    sp = golden["survey_probe"]
    assert O.commit(pattern(1024), 4).hex() == sp["commit_pattern_1024"]
    assert O.commit(pattern(65536), 4).hex() == sp["commit_pattern_65536"]
    assert O.commit(pattern(131072), 4).hex() == sp["commit_pattern_131072"]
    assert O.commit(bytes(131072), 4).hex() == sp["commit_zeros_131072"]
    assert O.commit(b"This is the original data that needs to be made available.", 4).hex() == sp[
        "commit_e2e_string"]
    h = O.blake2s_compress([0] * 8, [0] * 16)
    assert b"".join(x.to_bytes(4, "little") for x in h).hex() == sp["compress_zero"]


def test_survey_probe_blob_intermediates(blob_bytes, golden):
    sp = golden["survey_probe"]
    f = O.bytes_to_felts(blob_bytes)
    assert len(f) == sp["blob_felts"]["count"]
    assert [int(x) for x in f[:4]] == sp["blob_felts"]["first4"] and int(f[-1]) == sp["blob_felts"]["last"]
    tw, itw = O.precompute_twiddles(18)
    assert [int(x) for x in tw[:4]] == sp["twiddles_k18"]["first4"]
    assert int(tw[-2]) == sp["twiddles_k18"]["penultimate"] and int(tw[-1]) == 1
    assert all(O.lib().fo_m31_mul(int(a), int(b)) == 1 for a, b in zip(tw[:64], itw[:64]))
    t = O.trace(blob_bytes, None, O.make_config(4, 1, 20, 20))
    assert [int(x) for x in t.layer_columns[0][0][:4]] == sp["blob_col0_first4"]
    assert not t.layer_columns[0][3].any()
    assert t.tree_levels[0][19][0].tobytes().hex() == sp["blob_leaf0"]
    assert t.tree_levels[0][18][0].tobytes().hex() == sp["blob_layer18_node0"]
    pp = sp["blob_proof_cfg_4_1_20_20"]
    for i, (pre, suf) in enumerate(pp["inner_root_prefix_suffix"]):
        r = t.tree_levels[i + 1][0].tobytes().hex()
        assert r.startswith(pre) and r.endswith(suf)
    assert [list(q) for q in t.last_layer_poly] == pp["last_layer_poly"]
    assert t.nonce == pp["nonce"]
    assert [int(q) for q in t.queries[:5]] == pp["first_queries"]
    assert len(t.layer_logs) - 1 == 13  # inner layers for C1 with log_last 1


def test_blake2s_256_matches_hashlib():
    for n in (0, 1, 31, 32, 48, 63, 64, 65, 127, 128, 129, 1000):
        d = pattern(n)
        assert O.blake2s_256(d) == hashlib.blake2s(d).digest()


def test_circle_point_generator_order():
    L = O.lib()
    import ctypes as C
    x, y = C.c_uint32(), C.c_uint32()
    L.fo_circle_point(1, C.byref(x), C.byref(y))
    assert (x.value, y.value) == (2, 1268011823)
    assert (x.value * x.value + y.value * y.value) % P == 1
    L.fo_circle_point(1 << 30, C.byref(x), C.byref(y))
    assert (x.value, y.value) == (P - 1, 0)


@pytest.mark.parametrize("log_size", [1, 2, 3, 4, 5, 7])
def test_fft_matches_semantic_definition(log_size):
    # out[brev(i)] = f(domain.at(i)) with the phi basis; covers stwo's D=1,2 special cases.
    rng = np.random.default_rng(log_size)
    for n_log in range(0, log_size + 1):
        coeffs = rng.integers(0, P, 1 << n_log, dtype=np.uint32)
        assert np.array_equal(O.circle_fft(coeffs, log_size), O.circle_eval_naive(coeffs, log_size))


def test_oracle_generated_vectors_stable(golden, blob_bytes):
    for c in golden["oracle_generated"]["commit"][:9]:
        if c["name"] == "empty":
            data = b""
        elif c["name"] == "one_byte":
            data = b"\x07"
        else:
            data = pattern(c["len"])
        assert O.commit(data, c["log_blowup"]).hex() == c["root"], c["name"]


PCS = O.make_config(4, 1, 20, 20)  # src/proof.rs:109-116


@pytest.fixture(scope="module")
def blob_proof(blob_bytes):
    return O.prove(blob_bytes, None, PCS)


def test_generate_proof_and_commitment(blob_bytes, blob_proof):
    root, pr = blob_proof
    assert pr.c.n_inner_layers != 0                      # src/proof.rs:119-124
    assert root == O.commit(blob_bytes, 4)                # src/proof.rs:126-135
    assert bytes(pr.c.first_layer.commitment) == root


def test_verify_and_tamper(blob_proof):
    _, pr = blob_proof
    assert O.verify(pr, None)                             # src/proof.rs:136-141
    p2 = pr.clone()
    p2.c.proof_of_work += 1
    assert not O.verify(p2, None)                         # :143-149
    p2 = pr.clone()
    for j in range(4):
        p2.c.evaluations[0].v[j] = (p2.c.evaluations[0].v[j] + 1) % P
    assert not O.verify(p2, None)                         # :151-157
    p2 = pr.clone()
    ev = pr.evaluations
    for i in range(len(ev)):
        for j in range(4):
            p2.c.evaluations[i].v[j] = ev[len(ev) - 1 - i][j]
    assert not O.verify(p2, None)                         # :158-164
    p2 = pr.clone()
    for j in range(4):
        p2.c.evaluations[0].v[j], p2.c.evaluations[1].v[j] = ev[1][j], ev[0][j]
    assert not O.verify(p2, None)                         # :175-181
    p2 = pr.clone()
    p2.c.n_evaluations -= 1
    with pytest.raises(O.OraclePanic):                    # :166-173 (#[should_panic])
        O.verify(p2, None)
    p2.c.n_evaluations += 1


def test_verify_with_seed(blob_bytes):
    # src/proof.rs:183-193
    _, p1 = O.prove(blob_bytes, 1, PCS)
    _, p2 = O.prove(blob_bytes, 2, PCS)
    assert p1.evaluations != p2.evaluations
    assert O.verify(p1, 1) and O.verify(p2, 2)
    assert not O.verify(p1, 2) and not O.verify(p2, 1)


def test_end_to_end_small():
    # src/lib.rs:52-85
    data = b"This is the original data that needs to be made available."
    _, pr = O.prove(data, None, O.make_config(4, 0, 20, 20))
    assert O.verify(pr, None)


def test_prover_panics_like_reference_on_tiny_input():
    # poly_log 0 with log_last 0: commit_last_layer's assert_eq! fires in the reference.
    with pytest.raises(O.OraclePanic):
        O.prove(b"\x01\x02", None, O.make_config(4, 0, 20, 4))
    with pytest.raises(O.OraclePanic):
        O.commit(b"\x01", 0)  # half_odds(0 + 0 - 1) underflows
