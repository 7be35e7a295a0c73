"""ctypes loader for the CPU oracle (oracle/frieda_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under frieda_b200/ does.  See oracle/frieda_oracle.h for the
parity status ("parity unpinned" for the FRI transcript conventions).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libfrieda_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "frieda_oracle.c")
    hdr = os.path.join(_HERE, "frieda_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_LIB_PATH) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class QM31(C.Structure):
    _fields_ = [("v", C.c_uint32 * 4)]

    def tuple(self):
        return tuple(int(x) for x in self.v)


class PcsConfig(C.Structure):
    _fields_ = [
        ("log_blowup_factor", C.c_uint32),
        ("log_last_layer_degree_bound", C.c_uint32),
        ("n_queries", C.c_uint64),
        ("pow_bits", C.c_uint32),
    ]


class LayerProof(C.Structure):
    _fields_ = [
        ("commitment", C.c_uint8 * 32),
        ("n_fri_witness", C.c_uint32),
        ("fri_witness", C.POINTER(QM31)),
        ("n_hash_witness", C.c_uint32),
        ("hash_witness", C.POINTER(C.c_uint8)),
        ("n_column_witness", C.c_uint32),
        ("column_witness", C.POINTER(C.c_uint32)),
    ]


class Proof(C.Structure):
    _fields_ = [
        ("first_layer", LayerProof),
        ("n_inner_layers", C.c_uint32),
        ("inner_layers", C.POINTER(LayerProof)),
        ("n_last_layer_poly", C.c_uint32),
        ("last_layer_poly", C.POINTER(QM31)),
        ("proof_of_work", C.c_uint64),
        ("pcs_config", PcsConfig),
        ("log_size_bound", C.c_uint32),
        ("n_evaluations", C.c_uint32),
        ("evaluations", C.POINTER(QM31)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_LIB_PATH)
    u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
    L.fo_m31_mul.restype = C.c_uint32
    L.fo_m31_mul.argtypes = [C.c_uint32, C.c_uint32]
    L.fo_m31_inv.restype = C.c_uint32
    L.fo_m31_inv.argtypes = [C.c_uint32]
    L.fo_qm31_mul.restype = QM31
    L.fo_qm31_mul.argtypes = [QM31, QM31]
    L.fo_circle_point.argtypes = [C.c_uint32, u32p, u32p]
    L.fo_blake2s_compress.argtypes = [u32p, u32p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
    L.fo_blake2s_256.argtypes = [C.c_char_p, C.c_size_t, u8p]
    L.fo_bytes_to_felts.restype = C.c_size_t
    L.fo_bytes_to_felts.argtypes = [C.c_char_p, C.c_size_t, u32p, C.c_size_t]
    L.fo_poly_log.restype = C.c_uint32
    L.fo_poly_log.argtypes = [C.c_size_t]
    L.fo_precompute_twiddles.argtypes = [C.c_uint32, u32p, u32p]
    L.fo_circle_fft.argtypes = [u32p, C.c_uint32, u32p]
    L.fo_circle_eval_naive.argtypes = [u32p, C.c_uint32, C.c_uint32, u32p]
    L.fo_commit.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, u8p]
    L.fo_prove.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(PcsConfig), u8p,
                           C.POINTER(C.POINTER(Proof))]
    L.fo_verify.argtypes = [C.POINTER(Proof), C.POINTER(C.c_uint64)]
    L.fo_proof_free.argtypes = [C.POINTER(Proof)]
    L.fo_proof_clone.restype = C.POINTER(Proof)
    L.fo_proof_clone.argtypes = [C.POINTER(Proof)]
    L.fo_proof_serialize.restype = C.c_size_t
    L.fo_proof_serialize.argtypes = [C.POINTER(Proof), u8p, C.c_size_t]
    L.fo_trace_run.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(PcsConfig), C.c_int,
                               C.POINTER(C.c_void_p)]
    L.fo_trace_free.argtypes = [C.c_void_p]
    for name in ("fo_trace_poly_log", "fo_trace_n_felts", "fo_trace_n_layers", "fo_trace_n_queries"):
        getattr(L, name).restype = C.c_uint32
        getattr(L, name).argtypes = [C.c_void_p]
    L.fo_trace_coeffs.restype = u32p
    L.fo_trace_coeffs.argtypes = [C.c_void_p]
    L.fo_trace_twiddles.restype = u32p
    L.fo_trace_twiddles.argtypes = [C.c_void_p, C.c_int]
    L.fo_trace_layer_log.restype = C.c_uint32
    L.fo_trace_layer_log.argtypes = [C.c_void_p, C.c_uint32]
    L.fo_trace_layer_column.restype = u32p
    L.fo_trace_layer_column.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.fo_trace_tree_level.restype = u8p
    L.fo_trace_tree_level.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
    L.fo_trace_alpha.restype = QM31
    L.fo_trace_alpha.argtypes = [C.c_void_p, C.c_uint32]
    L.fo_trace_last_eval.restype = u32p
    L.fo_trace_last_eval.argtypes = [C.c_void_p, u32p]
    L.fo_trace_last_eval_col.restype = u32p
    L.fo_trace_last_eval_col.argtypes = [C.c_void_p, C.c_uint32]
    L.fo_trace_digest_after_fri.restype = u8p
    L.fo_trace_digest_after_fri.argtypes = [C.c_void_p]
    L.fo_trace_nonce.restype = C.c_uint64
    L.fo_trace_nonce.argtypes = [C.c_void_p]
    L.fo_trace_queries.restype = u32p
    L.fo_trace_queries.argtypes = [C.c_void_p]
    L.fo_trace_proof.restype = C.POINTER(Proof)
    L.fo_trace_proof.argtypes = [C.c_void_p]
    L.fo_trace_root.restype = u8p
    L.fo_trace_root.argtypes = [C.c_void_p]
    L.fo_fri_commit.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(PcsConfig), u8p,
                                C.c_uint32, u32p, C.POINTER(QM31), C.c_uint32]
    _lib = L
    return L


class OraclePanic(RuntimeError):
    """The reference would panic on this input (assert / unwrap / overflow)."""


def _check(rc: int):
    if rc == -1:
        raise OraclePanic("reference panics on this input")
    if rc < 0:
        raise MemoryError(f"oracle error {rc}")


def _seed_ptr(seed: Optional[int]):
    return C.byref(C.c_uint64(seed)) if seed is not None else None


def make_config(log_blowup=4, log_last=0, n_queries=20, pow_bits=20) -> PcsConfig:
    return PcsConfig(log_blowup, log_last, n_queries, pow_bits)


def bytes_to_felts(data: bytes) -> np.ndarray:
    n = (len(data) * 8 + 29) // 30
    out = np.zeros(max(n, 1), dtype=np.uint32)
    lib().fo_bytes_to_felts(data, len(data), out.ctypes.data_as(C.POINTER(C.c_uint32)), n)
    return out[:n]


def poly_log(length: int) -> int:
    return int(lib().fo_poly_log(length))


def precompute_twiddles(k: int):
    tw = np.zeros(1 << k, dtype=np.uint32)
    itw = np.zeros(1 << k, dtype=np.uint32)
    p = C.POINTER(C.c_uint32)
    _check(lib().fo_precompute_twiddles(k, tw.ctypes.data_as(p), itw.ctypes.data_as(p)))
    return tw, itw


def circle_fft(coeffs: np.ndarray, log_size: int) -> np.ndarray:
    v = np.zeros(1 << log_size, dtype=np.uint32)
    v[: len(coeffs)] = coeffs
    tw, _ = precompute_twiddles(log_size - 1)
    p = C.POINTER(C.c_uint32)
    _check(lib().fo_circle_fft(v.ctypes.data_as(p), log_size, tw.ctypes.data_as(p)))
    return v


def circle_eval_naive(coeffs: np.ndarray, log_size: int) -> np.ndarray:
    n_log = int(len(coeffs)).bit_length() - 1
    assert 1 << n_log == len(coeffs)
    c = np.ascontiguousarray(coeffs, dtype=np.uint32)
    out = np.zeros(1 << log_size, dtype=np.uint32)
    p = C.POINTER(C.c_uint32)
    _check(lib().fo_circle_eval_naive(c.ctypes.data_as(p), n_log, log_size, out.ctypes.data_as(p)))
    return out


def blake2s_compress(h, m, t0=0, t1=0, f0=0, f1=0):
    hh = (C.c_uint32 * 8)(*h)
    mm = (C.c_uint32 * 16)(*m)
    lib().fo_blake2s_compress(hh, mm, t0, t1, f0, f1)
    return [int(x) for x in hh]


def blake2s_256(data: bytes) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().fo_blake2s_256(data, len(data), out)
    return bytes(out)


def commit(data: bytes, log_blowup: int) -> bytes:
    out = (C.c_uint8 * 32)()
    _check(lib().fo_commit(data, len(data), log_blowup, out))
    return bytes(out)


class ProofHandle:
    """Owns an fo_proof*; mirrors frieda::proof::Proof (src/proof.rs:19-26)."""

    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        if getattr(self, "ptr", None):
            lib().fo_proof_free(self.ptr)
            self.ptr = None

    @property
    def c(self) -> Proof:
        return self.ptr.contents

    def clone(self) -> "ProofHandle":
        return ProofHandle(lib().fo_proof_clone(self.ptr))

    def serialize(self) -> bytes:
        n = lib().fo_proof_serialize(self.ptr, None, 0)
        buf = (C.c_uint8 * n)()
        lib().fo_proof_serialize(self.ptr, buf, n)
        return bytes(buf)

    @property
    def evaluations(self):
        return [self.c.evaluations[i].tuple() for i in range(self.c.n_evaluations)]

    @property
    def last_layer_poly(self):
        return [self.c.last_layer_poly[i].tuple() for i in range(self.c.n_last_layer_poly)]


def prove(data: bytes, seed: Optional[int], cfg: PcsConfig):
    root = (C.c_uint8 * 32)()
    pp = C.POINTER(Proof)()
    _check(lib().fo_prove(data, len(data), _seed_ptr(seed), C.byref(cfg), root, C.byref(pp)))
    return bytes(root), ProofHandle(pp)


def verify(proof: ProofHandle, seed: Optional[int]) -> bool:
    rc = lib().fo_verify(proof.ptr, _seed_ptr(seed))
    _check(rc)
    return bool(rc)


def fri_commit(data: bytes, seed: Optional[int], cfg: PcsConfig):
    roots = (C.c_uint8 * (32 * 40))()
    n_layers = C.c_uint32()
    n_last = 1 << cfg.log_last_layer_degree_bound
    last = (QM31 * n_last)()
    _check(lib().fo_fri_commit(data, len(data), _seed_ptr(seed), C.byref(cfg), roots, 40, C.byref(n_layers), last,
                               n_last))
    rb = bytes(roots)
    return [rb[32 * i: 32 * i + 32] for i in range(n_layers.value)], [q.tuple() for q in last]


@dataclass
class Trace:
    """All intermediates of one proof run, copied out as numpy arrays."""
    poly_log: int = 0
    n_felts: int = 0
    coeffs: np.ndarray = None
    twiddles: np.ndarray = None
    itwiddles: np.ndarray = None
    layer_logs: List[int] = field(default_factory=list)
    layer_columns: List[np.ndarray] = field(default_factory=list)   # [layer] -> (4, 2^log) u32
    tree_levels: List[List[np.ndarray]] = field(default_factory=list)  # [layer][level] -> (2^level, 32) u8
    alphas: List[tuple] = field(default_factory=list)
    last_eval: np.ndarray = None     # (4, 2^last_log)
    digest_after_fri: bytes = b""
    nonce: int = 0
    queries: np.ndarray = None
    root: bytes = b""
    last_layer_poly: List[tuple] = field(default_factory=list)
    proof_bytes: bytes = b""


def trace(data: bytes, seed: Optional[int], cfg: PcsConfig, stop_after_fri: bool = False,
          with_trees: bool = True) -> Trace:
    L = lib()
    h = C.c_void_p()
    _check(L.fo_trace_run(data, len(data), _seed_ptr(seed), C.byref(cfg), int(stop_after_fri), C.byref(h)))
    try:
        t = Trace()
        t.poly_log = L.fo_trace_poly_log(h)
        t.n_felts = L.fo_trace_n_felts(h)
        t.coeffs = np.ctypeslib.as_array(L.fo_trace_coeffs(h), shape=(4 << t.poly_log,)).copy()
        D = L.fo_trace_layer_log(h, 0)
        t.twiddles = np.ctypeslib.as_array(L.fo_trace_twiddles(h, 0), shape=(1 << (D - 1),)).copy()
        t.itwiddles = np.ctypeslib.as_array(L.fo_trace_twiddles(h, 1), shape=(1 << (D - 1),)).copy()
        for layer in range(L.fo_trace_n_layers(h)):
            lg = L.fo_trace_layer_log(h, layer)
            t.layer_logs.append(lg)
            cols = np.stack([
                np.ctypeslib.as_array(L.fo_trace_layer_column(h, layer, c), shape=(1 << lg,)) for c in range(4)
            ]).copy()
            t.layer_columns.append(cols)
            if with_trees:
                t.tree_levels.append([
                    np.ctypeslib.as_array(L.fo_trace_tree_level(h, layer, k), shape=(1 << k, 32)).copy()
                    for k in range(lg + 1)
                ])
            t.alphas.append(L.fo_trace_alpha(h, layer).tuple())
        ll = C.c_uint32()
        L.fo_trace_last_eval(h, C.byref(ll))
        t.last_eval = np.stack([
            np.ctypeslib.as_array(L.fo_trace_last_eval_col(h, c), shape=(1 << ll.value,)) for c in range(4)
        ]).copy()
        t.digest_after_fri = bytes(np.ctypeslib.as_array(L.fo_trace_digest_after_fri(h), shape=(32,)))
        t.root = bytes(np.ctypeslib.as_array(L.fo_trace_root(h), shape=(32,)))
        pr = L.fo_trace_proof(h)
        t.last_layer_poly = [pr.contents.last_layer_poly[i].tuple() for i in range(pr.contents.n_last_layer_poly)]
        if not stop_after_fri:
            t.nonce = L.fo_trace_nonce(h)
            nq = L.fo_trace_n_queries(h)
            t.queries = np.ctypeslib.as_array(L.fo_trace_queries(h), shape=(nq,)).copy()
            n = L.fo_proof_serialize(pr, None, 0)
            buf = (C.c_uint8 * n)()
            L.fo_proof_serialize(pr, buf, n)
            t.proof_bytes = bytes(buf)
        return t
    finally:
        L.fo_trace_free(h)


def splitmix64_bytes(state0: int, n: int) -> bytes:
    """synthetic blob generator (SplitMix64 stream, 8 bytes LE per draw)."""
    M = (1 << 64) - 1
    n_words = (n + 7) // 8
    with np.errstate(over="ignore"):
        idx = np.arange(1, n_words + 1, dtype=np.uint64)
        s = np.uint64(state0 & M) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = s
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z.astype("<u8").tobytes()[:n]
