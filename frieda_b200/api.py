"""Host-side mirror of the reference crate's API over libfrieda_b200.so (ctypes).

Same names and argument meaning as the reference (no Rust toolchain exists in this image, so
the Rust shim of INTEGRATION.md cannot be compiled here; this module is the tested host side):

    frieda::api::commit(data, log_blowup_factor) -> Commitment            src/lib.rs:31
    frieda::api::generate_proof(data, seed, pcs_config) -> Proof          src/lib.rs:36
    frieda::api::verify(proof, seed) -> bool                              src/lib.rs:41
    frieda::proof::commit_and_generate_proof(...) -> (Commitment, Proof)  src/proof.rs:32

There is no CPU fallback: every compute call needs the CUDA library and a CUDA device and raises
FriedaError otherwise.  `verify` is host-only by design.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

ERR_PANIC, ERR_ALLOC, ERR_CUDA, ERR_ARG = -1, -2, -3, -4


class FriedaError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"frieda_b200 error {code}: {msg}")
        self.code = code


class ReferencePanic(FriedaError):
    """The reference implementation panics on this input (assert / unwrap)."""


class QM31(C.Structure):
    _fields_ = [("v", C.c_uint32 * 4)]

    def tuple(self):
        return tuple(int(x) for x in self.v)


class PcsConfig(C.Structure):
    """stwo PcsConfig { pow_bits, fri_config { log_blowup_factor, log_last_layer_degree_bound, n_queries } }."""
    _fields_ = [
        ("log_blowup_factor", C.c_uint32),
        ("log_last_layer_degree_bound", C.c_uint32),
        ("n_queries", C.c_uint64),
        ("pow_bits", C.c_uint32),
    ]

    def __init__(self, log_blowup_factor=4, log_last_layer_degree_bound=0, n_queries=20, pow_bits=20):
        super().__init__(log_blowup_factor, log_last_layer_degree_bound, n_queries, pow_bits)


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


class ProofStruct(C.Structure):
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


# every symbol include/frieda_b200.h declares
EXPORTS = [
    "frieda_ctx_create", "frieda_ctx_destroy", "frieda_last_error", "frieda_ctx_set_workspace_limit",
    "frieda_ctx_launch_count", "frieda_ctx_stream", "frieda_ctx_set_profiling", "frieda_ctx_profile_read", "frieda_commit", "frieda_commit_batch",
    "frieda_commit_batch_device", "frieda_fri_n_inner_layers", "frieda_fri_commit_batch",
    "frieda_fri_commit_batch_device", "frieda_prove", "frieda_prove_batch", "frieda_verify", "frieda_verify_batch",
    "frieda_verify_batch_bytes",
    "frieda_verify_core_host", "frieda_verify_core_host_bytes", "frieda_ctx_take_error", "frieda_proof_query_positions", "frieda_proof_free",
    "frieda_proof_clone", "frieda_proof_serialize", "frieda_proof_deserialize", "frieda_proof_serialize_bincode", "frieda_commit_split_local",
    "frieda_commit_split_local_device", "frieda_commit_split_local_peers", "frieda_merkle_combine_peers",
    "frieda_commit_split_peers",
    "frieda_merkle_combine", "frieda_decode_block", "frieda_decode_blocks",
    "frieda_fri_split_begin", "frieda_fri_split_begin_device", "frieda_fri_split_begin_peers", "frieda_fri_split_layer", "frieda_fri_split_combine", "frieda_fri_split_layers_peers",
    "frieda_fri_split_handoff", "frieda_fri_split_finish", "frieda_fri_split_decommit", "frieda_fri_split_assemble",
    "frieda_buffer_free", "frieda_pass_pack", "frieda_pass_lde", "frieda_pass_merkle", "frieda_pass_fold",
    "frieda_twiddles", "frieda_debug_fetch", "frieda_ctx_set_debug_keep",
]

_lib = None


def load_library(build_if_missing: bool = True):
    """Loads libfrieda_b200.so (building it with nvcc when stale).  Fails loudly if impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB
    if build_if_missing and _build.is_stale():
        _build.build()
    if not os.path.exists(path):
        raise FriedaError(ERR_CUDA, f"{path} is missing: build it with `python -m frieda_b200.build` "
                                    "(there is no CPU fallback)")
    L = C.CDLL(path)
    u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
    vp, sz = C.c_void_p, C.c_size_t
    cfgp, qp = C.POINTER(PcsConfig), C.POINTER(QM31)
    pp = C.POINTER(ProofStruct)
    sig = {
        "frieda_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "frieda_ctx_destroy": (None, [vp]),
        "frieda_last_error": (C.c_char_p, [vp]),
        "frieda_ctx_set_workspace_limit": (C.c_int, [vp, sz]),
        "frieda_ctx_launch_count": (C.c_uint64, [vp]),
        "frieda_ctx_stream": (vp, [vp]),
        "frieda_ctx_set_profiling": (C.c_int, [vp, C.c_int]),
        "frieda_ctx_profile_read": (C.c_long, [vp, C.c_char_p, sz, C.c_int]),
        "frieda_commit": (C.c_int, [vp, vp, sz, C.c_uint32, u8p]),
        "frieda_commit_batch": (C.c_int, [vp, vp, sz, sz, sz, C.c_uint32, vp]),
        "frieda_commit_batch_device": (C.c_int, [vp, vp, sz, sz, sz, C.c_uint32, vp]),
        "frieda_fri_n_inner_layers": (C.c_int, [sz, cfgp]),
        "frieda_fri_commit_batch": (C.c_int, [vp, vp, sz, sz, sz, vp, cfgp, vp, vp]),
        "frieda_fri_commit_batch_device": (C.c_int, [vp, vp, sz, sz, sz, vp, cfgp, vp, vp]),
        "frieda_prove": (C.c_int, [vp, vp, sz, u64p, cfgp, u8p, C.POINTER(pp)]),
        "frieda_prove_batch": (C.c_int, [vp, vp, sz, sz, sz, vp, cfgp, vp, C.POINTER(pp)]),
        "frieda_verify": (C.c_int, [pp, u64p]),
        "frieda_verify_batch": (C.c_int, [vp, C.POINTER(pp), sz, vp, C.POINTER(C.c_int)]),
        "frieda_verify_batch_bytes": (C.c_int, [vp, vp, sz, vp, sz, vp, C.POINTER(C.c_int)]),
        "frieda_verify_core_host": (C.c_int, [pp, u64p]),
        "frieda_verify_core_host_bytes": (C.c_int, [vp, sz, u64p]),
        "frieda_ctx_take_error": (C.c_int, [vp]),
        "frieda_proof_query_positions": (C.c_longlong, [pp, u64p, C.POINTER(C.c_uint32), sz]),
        "frieda_proof_free": (None, [pp]),
        "frieda_proof_clone": (pp, [pp]),
        "frieda_proof_serialize": (sz, [pp, vp, sz]),
        "frieda_proof_deserialize": (C.c_int, [C.c_char_p, sz, C.POINTER(pp)]),
        "frieda_proof_serialize_bincode": (sz, [pp, vp, sz]),
        "frieda_commit_split_local": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
        "frieda_commit_split_local_device": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint32, C.c_uint32, vp]),
        "frieda_merkle_combine": (C.c_int, [vp, vp, C.c_uint32, u8p]),
        "frieda_commit_split_local_peers": (C.c_int, [vp, C.POINTER(C.c_void_p), C.c_uint32, sz, sz, C.c_uint32,
                                                      C.c_uint32, vp]),
        "frieda_merkle_combine_peers": (C.c_int, [vp, C.POINTER(C.c_void_p), C.c_uint32, u8p]),
        "frieda_commit_split_peers": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), sz,
                                                C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32, u8p]),
        "frieda_fri_split_begin": (C.c_int, [vp, vp, sz, u64p, cfgp, C.c_uint32, C.c_uint32, C.c_int, u32p, u32p, u32p]),
        "frieda_fri_split_begin_device": (C.c_int, [vp, vp, sz, u64p, cfgp, C.c_uint32, C.c_uint32, C.c_int, u32p, u32p,
                                                    u32p]),
        "frieda_fri_split_decommit": (C.c_int, [vp, C.POINTER(vp), C.POINTER(sz)]),
        "frieda_fri_split_assemble": (C.c_int, [C.POINTER(vp), C.POINTER(sz), C.c_uint32, C.POINTER(pp)]),
        "frieda_buffer_free": (None, [vp]),
        "frieda_fri_split_layer": (C.c_int, [vp, C.c_uint32, vp]),
        "frieda_fri_split_combine": (C.c_int, [vp, C.c_uint32, vp]),
        "frieda_fri_split_begin_peers": (C.c_int, [vp, vp, C.c_size_t, vp, vp, C.c_uint32, C.c_uint32, C.c_int, vp,
                                                   C.c_size_t, vp, C.c_uint32, vp, vp, vp]),
        "frieda_fri_split_layers_peers": (C.c_int, [vp, vp, vp, C.c_uint32]),
        "frieda_fri_split_handoff": (C.c_int, [vp, vp]),
        "frieda_fri_split_finish": (C.c_int, [vp, vp, vp, vp]),
        "frieda_decode_block": (C.c_int, [vp, vp, sz, C.c_uint32, C.c_uint32, vp]),
        "frieda_decode_blocks": (C.c_int, [vp, vp, vp, sz, sz, C.c_uint32, vp, C.POINTER(C.c_uint32)]),
        "frieda_pass_pack": (C.c_int, [vp, vp, sz, sz, sz, vp]),
        "frieda_pass_lde": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, sz, C.c_uint32, vp]),
        "frieda_pass_merkle": (C.c_int, [vp, vp, C.c_uint32, sz, vp, vp]),
        "frieda_pass_fold": (C.c_int, [vp, vp, C.c_uint32, C.c_int, sz, vp, vp]),
        "frieda_twiddles": (C.c_int, [vp, C.c_uint32, vp, vp]),
        "frieda_debug_fetch": (C.c_long, [vp, C.c_int, sz, C.c_uint32, C.c_uint32, vp, sz]),
        "frieda_ctx_set_debug_keep": (C.c_int, [vp, C.c_int]),
    }
    for name in EXPORTS:
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = sig[name]
    _lib = L
    return L


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        a = data
        if a.dtype != np.uint8:
            a = a.view(np.uint8)
        return np.ascontiguousarray(a).reshape(-1)
    return np.frombuffer(bytes(data), dtype=np.uint8)


class Proof:
    """Owns a frieda_proof*: fields mirror frieda::proof::Proof (src/proof.rs:19-26)."""

    def __init__(self, ptr):
        self._ptr = ptr

    def __del__(self):
        ptr = getattr(self, "_ptr", None)
        if ptr and _lib is not None:
            _lib.frieda_proof_free(ptr)
            self._ptr = None

    @property
    def c(self) -> ProofStruct:
        return self._ptr.contents

    @property
    def ptr(self):
        return self._ptr

    def clone(self) -> "Proof":
        p = load_library().frieda_proof_clone(self._ptr)
        if not p:
            raise FriedaError(ERR_ALLOC, "clone failed")
        return Proof(p)

    def serialize(self) -> bytes:
        L = load_library()
        n = L.frieda_proof_serialize(self._ptr, None, 0)
        buf = (C.c_uint8 * n)()
        L.frieda_proof_serialize(self._ptr, buf, n)
        return bytes(buf)

    def serialize_bincode(self) -> bytes:
        """bincode-1.x layout of the reference's serde `Proof` (unvalidated against Rust)."""
        L = load_library()
        n = L.frieda_proof_serialize_bincode(self._ptr, None, 0)
        buf = (C.c_uint8 * n)()
        L.frieda_proof_serialize_bincode(self._ptr, buf, n)
        return bytes(buf)

    @staticmethod
    def deserialize(data: bytes) -> "Proof":
        L = load_library()
        p = C.POINTER(ProofStruct)()
        rc = L.frieda_proof_deserialize(data, len(data), C.byref(p))
        if rc:
            raise FriedaError(rc, "malformed proof bytes")
        return Proof(p)

    # -- field views ---------------------------------------------------------------
    @property
    def proof_of_work(self) -> int:
        return int(self.c.proof_of_work)

    @proof_of_work.setter
    def proof_of_work(self, v: int):
        self.c.proof_of_work = v

    @property
    def log_size_bound(self) -> int:
        return int(self.c.log_size_bound)

    @property
    def pcs_config(self) -> PcsConfig:
        return self.c.pcs_config

    @property
    def evaluations(self) -> List[tuple]:
        return [self.c.evaluations[i].tuple() for i in range(self.c.n_evaluations)]

    def set_evaluation(self, i: int, value: Sequence[int]):
        for j in range(4):
            self.c.evaluations[i].v[j] = int(value[j])

    def pop_evaluation(self):
        self.c.n_evaluations -= 1

    @property
    def last_layer_poly(self) -> List[tuple]:
        return [self.c.last_layer_poly[i].tuple() for i in range(self.c.n_last_layer_poly)]

    @property
    def first_layer_commitment(self) -> bytes:
        return bytes(self.c.first_layer.commitment)

    @property
    def inner_layer_commitments(self) -> List[bytes]:
        return [bytes(self.c.inner_layers[i].commitment) for i in range(self.c.n_inner_layers)]

    @property
    def n_inner_layers(self) -> int:
        return int(self.c.n_inner_layers)


class Context:
    """One CUDA device + stream + twiddle cache + workspace (frieda_ctx)."""

    def __init__(self, device: int = 0):
        self._L = load_library()
        h = C.c_void_p()
        rc = self._L.frieda_ctx_create(device, C.byref(h))
        if rc:
            msg = self._L.frieda_last_error(None)
            raise FriedaError(rc, (msg or b"").decode() or "context creation failed (no CPU fallback exists)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            import sys
            par = sys.modules.get("frieda_b200.parallel")
            if par is not None:  # peer-mapped buffers were used on this context's stream: free them first
                par.release_peer_memory(self)
            self._L.frieda_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc == 0:
            return
        msg = (self._L.frieda_last_error(self._h) or b"").decode()
        if rc == ERR_PANIC:
            raise ReferencePanic(rc, msg)
        raise FriedaError(rc, msg)

    @property
    def handle(self):
        return self._h

    @property
    def launch_count(self) -> int:
        return int(self._L.frieda_ctx_launch_count(self._h))

    @property
    def stream_ptr(self) -> int:
        return int(self._L.frieda_ctx_stream(self._h) or 0)

    def set_workspace_limit(self, nbytes: int):
        self._check(self._L.frieda_ctx_set_workspace_limit(self._h, nbytes))

    def set_profiling(self, on: bool):
        self._check(self._L.frieda_ctx_set_profiling(self._h, int(on)))

    def profile_read(self, reset: bool = True) -> dict:
        """{kernel name: (launches, total_ms)} measured with CUDA events on the context's stream."""
        buf = C.create_string_buffer(1 << 16)
        rc = self._L.frieda_ctx_profile_read(self._h, buf, len(buf), int(reset))
        if rc < 0:
            self._check(int(rc))
        out = {}
        for line in buf.value.decode().splitlines():
            name, n, ms = line.rsplit(" ", 2)
            out[name] = (int(n), float(ms))
        return out

    def take_error(self):
        """Synchronises the stream and raises what the asynchronous *_device entry points could not report when they
        returned (ReferencePanic for the reference's "invalid degree" assert); clears the pending error."""
        self._check(self._L.frieda_ctx_take_error(self._h))

    def set_debug_keep(self, on: bool):
        self._check(self._L.frieda_ctx_set_debug_keep(self._h, int(on)))

    # -- commit --------------------------------------------------------------------
    def commit(self, data, log_blowup_factor: int) -> bytes:
        a = _as_u8(data)
        out = (C.c_uint8 * 32)()
        self._check(self._L.frieda_commit(self._h, a.ctypes.data, a.size, log_blowup_factor, out))
        return bytes(out)

    def commit_batch(self, blobs: np.ndarray, log_blowup_factor: int) -> np.ndarray:
        """blobs: (n, blob_len) uint8, row-contiguous.  Returns (n, 32) uint8 roots."""
        assert blobs.ndim == 2 and blobs.dtype == np.uint8 and blobs.strides[1] == 1
        n, blob_len = blobs.shape
        roots = np.zeros((n, 32), dtype=np.uint8)
        self._check(self._L.frieda_commit_batch(self._h, blobs.ctypes.data, blob_len, blobs.strides[0], n,
                                                log_blowup_factor, roots.ctypes.data))
        return roots

    def commit_batch_ptr(self, blobs_ptr: int, blob_len: int, stride: int, n: int, log_blowup_factor: int,
                         roots_ptr: int, device: bool):
        fn = self._L.frieda_commit_batch_device if device else self._L.frieda_commit_batch
        self._check(fn(self._h, blobs_ptr, blob_len, stride, n, log_blowup_factor, roots_ptr))

    # -- FRI commit phase ----------------------------------------------------------
    def n_inner_layers(self, blob_len: int, cfg: PcsConfig) -> int:
        rc = self._L.frieda_fri_n_inner_layers(blob_len, C.byref(cfg))
        if rc < 0:
            raise (ReferencePanic if rc == ERR_PANIC else FriedaError)(rc, "invalid shape for this config")
        return rc

    def fri_commit_batch(self, blobs: np.ndarray, seeds: Optional[Sequence[int]], cfg: PcsConfig):
        """Returns (layer_roots (n, 1 + n_inner, 32) uint8, last_layer_poly (n, 2^log_last, 4) uint32)."""
        assert blobs.ndim == 2 and blobs.dtype == np.uint8 and blobs.strides[1] == 1
        n, blob_len = blobs.shape
        L = 1 + self.n_inner_layers(blob_len, cfg)
        roots = np.zeros((n, L, 32), dtype=np.uint8)
        last = np.zeros((n, 1 << cfg.log_last_layer_degree_bound, 4), dtype=np.uint32)
        sd = None
        if seeds is not None:
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
            assert sd.shape == (n,)
        self._check(self._L.frieda_fri_commit_batch(self._h, blobs.ctypes.data, blob_len, blobs.strides[0], n,
                                                    sd.ctypes.data if sd is not None else None, C.byref(cfg),
                                                    roots.ctypes.data, last.ctypes.data))
        return roots, last

    def fri_commit_batch_ptr(self, blobs_ptr: int, blob_len: int, stride: int, n: int, seeds_ptr: Optional[int],
                             cfg: PcsConfig, roots_ptr: int, last_ptr: int, device: bool):
        fn = self._L.frieda_fri_commit_batch_device if device else self._L.frieda_fri_commit_batch
        self._check(fn(self._h, blobs_ptr, blob_len, stride, n, seeds_ptr, C.byref(cfg), roots_ptr, last_ptr))

    # -- proofs --------------------------------------------------------------------
    def commit_and_generate_proof(self, data, seed: Optional[int], cfg: PcsConfig) -> Tuple[bytes, Proof]:
        a = _as_u8(data)
        root = (C.c_uint8 * 32)()
        p = C.POINTER(ProofStruct)()
        sp = C.byref(C.c_uint64(seed)) if seed is not None else None
        self._check(self._L.frieda_prove(self._h, a.ctypes.data, a.size, sp, C.byref(cfg), root, C.byref(p)))
        return bytes(root), Proof(p)

    def generate_proof(self, data, seed: Optional[int], cfg: PcsConfig) -> Proof:
        return self.commit_and_generate_proof(data, seed, cfg)[1]

    def prove_batch(self, blobs: np.ndarray, seeds: Optional[Sequence[int]], cfg: PcsConfig):
        assert blobs.ndim == 2 and blobs.dtype == np.uint8 and blobs.strides[1] == 1
        n, blob_len = blobs.shape
        roots = np.zeros((n, 32), dtype=np.uint8)
        arr = (C.POINTER(ProofStruct) * n)()
        sd = None
        if seeds is not None:
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
        self._check(self._L.frieda_prove_batch(self._h, blobs.ctypes.data, blob_len, blobs.strides[0], n,
                                               sd.ctypes.data if sd is not None else None, C.byref(cfg),
                                               roots.ctypes.data, arr))
        return roots, [Proof(arr[i]) for i in range(n)]

    def verify_batch(self, proofs: Sequence["Proof"], seeds: Optional[Sequence[int]]) -> List[int]:
        """GPU batch verification: 1 valid / 0 invalid / -1 where the reference panics, per proof."""
        n = len(proofs)
        arr = (C.POINTER(ProofStruct) * n)(*[p.ptr for p in proofs])
        res = (C.c_int * n)()
        sd = None
        if seeds is not None:
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
            assert sd.shape == (n,)
        self._check(self._L.frieda_verify_batch(self._h, arr, n, sd.ctypes.data if sd is not None else None, res))
        return [int(x) for x in res]

    def verify_batch_bytes(self, blob: np.ndarray, byte_offsets: np.ndarray, seeds: Optional[Sequence[int]]) -> List[int]:
        """GPU batch verification of serialised proofs: blob = concatenated Proof.serialize() outputs (uint8,
        4-byte aligned pieces), byte_offsets = n + 1 uint64 offsets."""
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        offs = np.ascontiguousarray(byte_offsets, dtype=np.uint64)
        n = len(offs) - 1
        if n < 0:
            raise FriedaError(ERR_ARG, "byte_offsets needs n + 1 entries")
        res = (C.c_int * max(n, 1))()
        sd = None
        if seeds is not None:
            sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64))
            if sd.shape != (n,):
                raise FriedaError(ERR_ARG, f"{n} proofs but {sd.shape} seeds")
        # the library checks every offset against blob.size (FRIEDA_ERR_ARG), so a bad offsets array cannot make
        # the upload read past the host buffer
        self._check(self._L.frieda_verify_batch_bytes(self._h, blob.ctypes.data, blob.size, offs.ctypes.data, n,
                                                      sd.ctypes.data if sd is not None else None, res))
        res = res[:n]
        return [int(x) for x in res]

    # -- split blob ----------------------------------------------------------------
    def commit_split_local(self, data, log_blowup_factor: int, rank: int, world: int, subroot_dev_ptr: int):
        a = _as_u8(data)
        self._check(self._L.frieda_commit_split_local(self._h, a.ctypes.data, a.size, log_blowup_factor, rank, world,
                                                      subroot_dev_ptr))

    def commit_split_local_device(self, data_dev_ptr: int, length: int, log_blowup_factor: int, rank: int, world: int,
                                  subroot_dev_ptr: int):
        self._check(self._L.frieda_commit_split_local_device(self._h, data_dev_ptr, length, log_blowup_factor, rank,
                                                             world, subroot_dev_ptr))

    def commit_split_local_peers(self, peer_slice_ptrs, slice_len: int, length: int, log_blowup_factor: int,
                                 rank: int, subroot_dev_ptr: int):
        """Slice r of the blob is read in place from peer_slice_ptrs[r] (peer-mapped device memory).
        Asynchronous on the context's stream."""
        world = len(peer_slice_ptrs)
        arr = (C.c_void_p * world)(*[int(p) for p in peer_slice_ptrs])
        self._check(self._L.frieda_commit_split_local_peers(self._h, arr, world, slice_len, length, log_blowup_factor,
                                                            rank, subroot_dev_ptr))

    def commit_split_peers(self, data, log_blowup_factor: int, rank: int, peer_slice_ptrs, slice_len: int,
                           peer_root_ptrs, peer_flag_ptrs, epoch: int, resident_len: Optional[int] = None) -> bytes:
        """The whole split commit of this rank in one call (frieda_commit_split_peers): exchanges over peer memory.
        resident_len: the blob's length when every rank's slice already lies in its symmetric buffer (data ignored)."""
        world = len(peer_slice_ptrs)
        mk = lambda ps: (C.c_void_p * world)(*[int(p) for p in ps])  # noqa: E731
        out = (C.c_uint8 * 32)()
        if resident_len is not None:
            self._check(self._L.frieda_commit_split_peers(self._h, None, resident_len, log_blowup_factor, rank, world,
                                                          mk(peer_slice_ptrs), slice_len, mk(peer_root_ptrs),
                                                          mk(peer_flag_ptrs), epoch, out))
            return bytes(out)
        a = _as_u8(data)
        self._check(self._L.frieda_commit_split_peers(self._h, a.ctypes.data, a.size, log_blowup_factor, rank, world,
                                                      mk(peer_slice_ptrs), slice_len, mk(peer_root_ptrs),
                                                      mk(peer_flag_ptrs), epoch, out))
        return bytes(out)

    def merkle_combine_peers(self, peer_root_ptrs) -> bytes:
        world = len(peer_root_ptrs)
        arr = (C.c_void_p * world)(*[int(p) for p in peer_root_ptrs])
        out = (C.c_uint8 * 32)()
        self._check(self._L.frieda_merkle_combine_peers(self._h, arr, world, out))
        return bytes(out)

    def merkle_combine(self, subroots_dev_ptr: int, world: int) -> bytes:
        out = (C.c_uint8 * 32)()
        self._check(self._L.frieda_merkle_combine(self._h, subroots_dev_ptr, world, out))
        return bytes(out)

    # -- FRI commit phase of one blob split over ranks (frieda_fri_split_*) ------------
    def fri_split_begin(self, data, seed: Optional[int], cfg: PcsConfig, rank: int, world: int,
                        device_ptr: Optional[int] = None, length: Optional[int] = None,
                        keep_trees: bool = False) -> Tuple[int, int, int]:
        """Returns (n_split_layers, n_layers, handoff_log).  data: host bytes / array, or device_ptr + length.
        keep_trees: a proof follows (fri_split_decommit)."""
        ns, nl, hl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        sp = C.byref(C.c_uint64(seed)) if seed is not None else None
        if device_ptr is not None:
            rc = self._L.frieda_fri_split_begin_device(self._h, device_ptr, length, sp, C.byref(cfg), rank, world,
                                                       int(keep_trees), C.byref(ns), C.byref(nl), C.byref(hl))
        else:
            a = _as_u8(data)
            rc = self._L.frieda_fri_split_begin(self._h, a.ctypes.data, a.size, sp, C.byref(cfg), rank, world,
                                                int(keep_trees), C.byref(ns), C.byref(nl), C.byref(hl))
        self._check(rc)
        return int(ns.value), int(nl.value), int(hl.value)

    def fri_split_begin_peers(self, data, seed: Optional[int], cfg: PcsConfig, rank: int, peer_slice_ptrs, slice_len: int,
                              peer_flag_ptrs, epoch: int, keep_trees: bool = False) -> Tuple[int, int, int]:
        """fri_split_begin with the blob in peer-mapped slices: this rank uploads only slice `rank` of `data`."""
        world = len(peer_slice_ptrs)
        mk = lambda ps: (C.c_void_p * world)(*[int(p) for p in ps])  # noqa: E731
        ns, nl, hl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        sp = C.byref(C.c_uint64(seed)) if seed is not None else None
        a = _as_u8(data)
        self._check(self._L.frieda_fri_split_begin_peers(self._h, a.ctypes.data, a.size, sp, C.byref(cfg), rank, world,
                                                         int(keep_trees), mk(peer_slice_ptrs), slice_len,
                                                         mk(peer_flag_ptrs), epoch, C.byref(ns), C.byref(nl), C.byref(hl)))
        return int(ns.value), int(nl.value), int(hl.value)

    def fri_split_layer(self, layer: int, subroot_dev_ptr: int):
        self._check(self._L.frieda_fri_split_layer(self._h, layer, subroot_dev_ptr))

    def fri_split_combine(self, layer: int, subroots_dev_ptr: int):
        self._check(self._L.frieda_fri_split_combine(self._h, layer, subroots_dev_ptr))

    def fri_split_layers_peers(self, peer_root_ptrs, peer_flag_ptrs, epoch: int):
        """All split layers in one call; the subtree roots travel through peer-mapped memory (frieda_fri_split_layers_peers).
        peer_root_ptrs[r] = base of rank r's symmetric roots area (64 x 32 bytes)."""
        world = len(peer_root_ptrs)
        mk = lambda ps: (C.c_void_p * world)(*[int(p) for p in ps])  # noqa: E731
        self._check(self._L.frieda_fri_split_layers_peers(self._h, mk(peer_root_ptrs), mk(peer_flag_ptrs), epoch))

    def fri_split_handoff(self, cols_local_dev_ptr: int):
        """This rank's 4 x 2^handoff_log u32 share of the first unsplit layer -> the caller's device buffer."""
        self._check(self._L.frieda_fri_split_handoff(self._h, cols_local_dev_ptr))

    def fri_split_finish(self, cols_all_dev_ptr: int, n_layers: int, log_last: int):
        roots = np.zeros((n_layers, 32), dtype=np.uint8)
        last = np.zeros((1 << log_last, 4), dtype=np.uint32)
        self._check(self._L.frieda_fri_split_finish(self._h, cols_all_dev_ptr, roots.ctypes.data, last.ctypes.data))
        return roots, last

    def fri_split_decommit(self) -> bytes:
        """Proof of work, queries and THIS rank's share of the decommitment (after fri_split_finish of a commit begun
        with keep_trees).  The shares of all ranks, in rank order, go to `split_assemble`."""
        p, n = C.c_void_p(0), C.c_size_t(0)
        self._check(self._L.frieda_fri_split_decommit(self._h, C.byref(p), C.byref(n)))
        try:
            return C.string_at(p.value, n.value)
        finally:
            self._L.frieda_buffer_free(p)

    # -- erasure recovery ------------------------------------------------------------
    def decode_block(self, block_evals: np.ndarray, length: int, log_blowup_factor: int, block: int) -> bytes:
        """Recovers the `length` original bytes from ONE coset block of the evaluation:
        block_evals = (4, 2^poly_log) uint32, the entries [block 2^p, (block+1) 2^p) of each column."""
        ev = np.ascontiguousarray(block_evals, dtype=np.uint32)
        out = np.zeros(max(length, 1), dtype=np.uint8)
        self._check(self._L.frieda_decode_block(self._h, ev.ctypes.data, length, log_blowup_factor, block,
                                                out.ctypes.data))
        return out[:length].tobytes()

    def decode_blocks(self, blocks: Sequence[np.ndarray], block_ids: Sequence[int], length: int,
                      log_blowup_factor: int) -> Tuple[bytes, int]:
        """Same from several whole coset blocks: returns (data, position in the list of the block it was decoded
        from); corrupted blocks are skipped."""
        ev = np.ascontiguousarray(np.stack([np.asarray(b, dtype=np.uint32) for b in blocks]))
        ids = np.ascontiguousarray(np.asarray(block_ids, dtype=np.uint32))
        out = np.zeros(max(length, 1), dtype=np.uint8)
        used = C.c_uint32(0)
        self._check(self._L.frieda_decode_blocks(self._h, ev.ctypes.data, ids.ctypes.data, len(ids), length,
                                                 log_blowup_factor, out.ctypes.data, C.byref(used)))
        return out[:length].tobytes(), int(used.value)

    # -- standalone passes / introspection ------------------------------------------
    def twiddles(self, k: int):
        tw = np.zeros(1 << k, dtype=np.uint32)
        itw = np.zeros(1 << k, dtype=np.uint32)
        self._check(self._L.frieda_twiddles(self._h, k, tw.ctypes.data, itw.ctypes.data))
        return tw, itw

    def pass_pack(self, d_blobs: int, blob_len: int, stride: int, n: int, d_coeffs: int):
        self._check(self._L.frieda_pass_pack(self._h, d_blobs, blob_len, stride, n, d_coeffs))

    def pass_lde(self, d_coeffs: int, poly_log: int, log_blowup: int, n: int, n_felts: int, d_evals: int):
        self._check(self._L.frieda_pass_lde(self._h, d_coeffs, poly_log, log_blowup, n, n_felts, d_evals))

    def pass_merkle(self, d_cols: int, log: int, n: int, d_tree: Optional[int], d_roots: int):
        self._check(self._L.frieda_pass_merkle(self._h, d_cols, log, n, d_tree, d_roots))

    def pass_fold(self, d_src: int, log: int, is_circle: bool, n: int, d_alpha: int, d_dst: int):
        self._check(self._L.frieda_pass_fold(self._h, d_src, log, int(is_circle), n, d_alpha, d_dst))

    def debug_fetch(self, what: int, blob: int, layer: int, level: int, nbytes: int) -> np.ndarray:
        out = np.zeros(nbytes, dtype=np.uint8)
        rc = self._L.frieda_debug_fetch(self._h, what, blob, layer, level, out.ctypes.data, nbytes)
        if rc < 0:
            self._check(int(rc))
        return out[:rc]


def verify_proof(proof: Proof, seed: Optional[int]) -> bool:
    """frieda::proof::verify_proof (src/proof.rs:79-101).  Raises ReferencePanic where the reference
    panics (too few evaluations, src/proof.rs:166-173)."""
    L = load_library()
    sp = C.byref(C.c_uint64(seed)) if seed is not None else None
    rc = L.frieda_verify(proof.ptr, sp)
    if rc == ERR_PANIC:
        raise ReferencePanic(rc, "reference panics on this proof")
    if rc < 0:
        raise FriedaError(rc, "verify failed")
    return bool(rc)


def query_positions(proof: Proof, seed: Optional[int]):
    """Sorted distinct positions (bit-reversed evaluation-domain indices) that `proof.evaluations` belong to; an empty
    list when the transcript is rejected before the queries are drawn."""
    L = load_library()
    sp = C.byref(C.c_uint64(seed)) if seed is not None else None
    n = int(L.frieda_proof_query_positions(proof.ptr, sp, None, 0))
    if n == ERR_PANIC:
        raise ReferencePanic(ERR_PANIC, "reference panics on this proof")
    if n < 0:
        raise FriedaError(n, "frieda_proof_query_positions failed")
    buf = (C.c_uint32 * max(n, 1))()
    L.frieda_proof_query_positions(proof.ptr, sp, buf, n)
    return [int(buf[i]) for i in range(n)]


def split_assemble(shares: Sequence[bytes]) -> Proof:
    """Merges the per-rank shares of a split blob's decommitment (rank order) into the Proof.  Host only."""
    L = load_library()
    world = len(shares)
    bufs = [np.frombuffer(bytes(s) + b"\0" * (-len(s) % 4), dtype=np.uint8).copy() for s in shares]
    ptrs = (C.c_void_p * world)(*[b.ctypes.data for b in bufs])
    lens = (C.c_size_t * world)(*[len(s) for s in shares])
    p = C.POINTER(ProofStruct)()
    rc = L.frieda_fri_split_assemble(ptrs, lens, world, C.byref(p))
    if rc:
        raise FriedaError(rc, "the shares do not form a proof (different blobs, ranks out of order, or corrupted)")
    return Proof(p)


def verify_core_host(proof: Proof, seed: Optional[int]) -> int:
    """The GPU batch verifier's core executed on the CPU for one proof: 1 / 0 / -1 (reference panics)."""
    L = load_library()
    sp = C.byref(C.c_uint64(seed)) if seed is not None else None
    return int(L.frieda_verify_core_host(proof.ptr, sp))


def verify_core_host_bytes(data: bytes, seed: Optional[int]) -> int:
    """Same over one serialised (untrusted) proof: the kernels' parser and both phases on the CPU, raw words in."""
    L = load_library()
    a = np.frombuffer(bytes(data) + b"\0" * (-len(data) % 4), dtype=np.uint8).copy()
    sp = C.byref(C.c_uint64(seed)) if seed is not None else None
    return int(L.frieda_verify_core_host_bytes(a.ctypes.data, len(data), sp))


# ---- module-level functions with the reference's names (default context on device 0) ----------
_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def commit(data, log_blowup_factor: int) -> bytes:
    return default_context().commit(data, log_blowup_factor)


def generate_proof(data, seed: Optional[int], pcs_config: PcsConfig) -> Proof:
    return default_context().generate_proof(data, seed, pcs_config)


def commit_and_generate_proof(data, seed: Optional[int], pcs_config: PcsConfig) -> Tuple[bytes, Proof]:
    return default_context().commit_and_generate_proof(data, seed, pcs_config)


def verify(proof: Proof, seed: Optional[int]) -> bool:
    return verify_proof(proof, seed)
