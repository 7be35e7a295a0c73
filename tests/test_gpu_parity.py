"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs,
bit-exact, on every intermediate (coefficients, twiddles, evaluations, every Merkle level, every
FRI layer, alphas, nonce, queries, witnesses), plus the reference's own test behaviours
(src/commit.rs:28-38, src/proof.rs:119-193, src/lib.rs:52-85) through the product API."""
import hashlib
import os

import numpy as np
import pytest

import frieda_b200 as F
from oracle import oracle as O

pytestmark = pytest.mark.gpu
P = (1 << 31) - 1
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def pattern(n):
    return bytes(i % 256 for i in range(n))


@pytest.fixture(scope="module")
def ctx():
    c = F.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    assert torch.cuda.is_available()
    return torch


def first_diff(a, b):
    a, b = np.asarray(a).reshape(-1), np.asarray(b).reshape(-1)
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    d = np.nonzero(a != b)[0]
    return None if d.size == 0 else f"{d.size} mismatches, first at {int(d[0])}: {a[d[0]]} vs {b[d[0]]}"


# ------------------------------------------------------------------ twiddles
@pytest.mark.parametrize("k", [0, 1, 2, 3, 7, 12, 17])
def test_twiddles(ctx, k):
    tw, itw = ctx.twiddles(k)
    otw, oitw = O.precompute_twiddles(k)
    assert first_diff(tw, otw) is None
    assert first_diff(itw, oitw) is None


def test_twiddles_small_after_large(ctx):
    ctx.twiddles(18)
    tw, itw = ctx.twiddles(9)  # served from the tail of the cached larger tree
    otw, oitw = O.precompute_twiddles(9)
    assert first_diff(tw, otw) is None and first_diff(itw, oitw) is None


# ------------------------------------------------------------------ commit (KAT + vectors)
def test_commit_blob_golden_root(ctx, blob_bytes, golden):
    # the reference's only known-answer test: src/commit.rs:28-38
    assert ctx.commit(blob_bytes, 4).hex() == golden["reference_certified"]["commit_blob_blowup4"]


def test_commit_survey_vectors(ctx, golden):
    sp = golden["survey_probe"]
    assert ctx.commit(pattern(1024), 4).hex() == sp["commit_pattern_1024"]
    assert ctx.commit(pattern(65536), 4).hex() == sp["commit_pattern_65536"]
    assert ctx.commit(pattern(131072), 4).hex() == sp["commit_pattern_131072"]
    assert ctx.commit(bytes(131072), 4).hex() == sp["commit_zeros_131072"]
    assert ctx.commit(b"This is the original data that needs to be made available.", 4).hex() == sp[
        "commit_e2e_string"]


def test_commit_golden_vector_file(ctx, golden):
    for c in golden["oracle_generated"]["commit"]:
        name = c["name"]
        if name == "empty":
            data = b""
        elif name == "one_byte":
            data = b"\x07"
        elif name.startswith("splitmix"):
            state = {"splitmix_c2": 0x4652494544410000, "splitmix_c2_b1": 0x4652494544410001,
                     "splitmix_1MiB_b2": 0x4652494544414236}[name]
            data = O.splitmix64_bytes(state, c["len"])
        else:
            data = pattern(c["len"])
        assert ctx.commit(data, c["log_blowup"]).hex() == c["root"], name


@pytest.mark.parametrize("n,blow", [(0, 1), (1, 1), (1, 2), (3, 1), (4, 3), (15, 1), (16, 1), (17, 4), (29, 2), (30, 2),
                                    (31, 5), (100, 1), (257, 4), (1000, 2), (4097, 3), (16385, 4), (70000, 1)])
def test_commit_ragged_sizes_vs_oracle(ctx, n, blow):
    rng = np.random.default_rng(n * 7 + blow)
    data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    assert ctx.commit(data, blow) == O.commit(data, blow)


def test_commit_panics_like_reference(ctx):
    with pytest.raises(F.ReferencePanic):
        ctx.commit(b"\x01", 0)  # Coset::half_odds(0 + 0 - 1)


def test_commit_batch_vs_oracle(ctx):
    rng = np.random.default_rng(5)
    for blob_len, n, blow in [(1000, 7, 3), (4096, 33, 4), (131072, 5, 4)]:
        blobs = rng.integers(0, 256, (n, blob_len), dtype=np.uint8)
        roots = ctx.commit_batch(blobs, blow)
        for i in range(n):
            assert roots[i].tobytes() == O.commit(blobs[i].tobytes(), blow), (blob_len, i)


def test_commit_batch_strided_and_waves(ctx):
    rng = np.random.default_rng(6)
    n, blob_len = 20, 5000
    big = rng.integers(0, 256, (n, blob_len + 24), dtype=np.uint8)
    view = big[:, :blob_len]  # stride > len
    ctx.set_workspace_limit(3 * (1 << 20))  # forces several waves
    try:
        roots = ctx.commit_batch(view, 4)
    finally:
        ctx.set_workspace_limit(0)
    for i in range(n):
        assert roots[i].tobytes() == O.commit(view[i].tobytes(), 4), i


def test_commit_batch_device_pointers(ctx, torch_mod):
    torch = torch_mod
    rng = np.random.default_rng(8)
    n, blob_len = 16, 131072
    blobs = rng.integers(0, 256, (n, blob_len), dtype=np.uint8)
    d_in = torch.from_numpy(blobs).cuda()
    d_out = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.commit_batch_ptr(d_in.data_ptr(), blob_len, blob_len, n, 4, d_out.data_ptr(), device=True)
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    got = d_out.cpu().numpy()
    want = ctx.commit_batch(blobs, 4)
    assert first_diff(got, want) is None
    assert got[3].tobytes() == O.commit(blobs[3].tobytes(), 4)


# ------------------------------------------------------------------ standalone passes
@pytest.mark.parametrize("p,beta,nz_frac", [(0, 3, 1.0), (1, 2, 1.0), (2, 1, 1.0), (1, 1, 1.0), (0, 1, 1.0), (3, 2, 1.0),
                                            (4, 2, 1.0), (5, 4, 0.6), (6, 1, 1.0), (7, 3, 1.0), (8, 1, 0.4),
                                            (9, 2, 1.0), (10, 1, 1.0), (11, 3, 0.3), (12, 1, 1.0), (13, 2, 0.7),
                                            (14, 4, 0.53), (15, 2, 1.0), (16, 1, 0.55), (17, 2, 0.9),
                                            # zero-padded columns: the top layers of the column holding the last felt only replicate
                                            (14, 2, 0.2501), (14, 1, 0.26), (12, 2, 0.2503), (15, 1, 0.30), (13, 1, 0.5003),
                                            (11, 2, 0.76), (10, 2, 0.27), (15, 1, 0.2500001),
                                            # radix-256 strided sweep (poly_log >= 20), alone and followed by radix-16
                                            (20, 1, 0.8), (21, 2, 0.3), (22, 1, 1.0), (23, 1, 0.55), (24, 1, 0.26)])
def test_lde_pass_vs_oracle_fft(ctx, torch_mod, p, beta, nz_frac):
    torch = torch_mod
    rng = np.random.default_rng(p * 31 + beta)
    n_blobs = 2 if p < 22 else 1
    n4 = 1 << p
    n_felts = max(1, int(4 * n4 * nz_frac))
    coef = np.zeros((n_blobs, 4 * n4), dtype=np.uint32)
    coef[:, :n_felts] = rng.integers(0, P, (n_blobs, n_felts), dtype=np.uint32)
    D = p + beta
    d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
    d_eval = torch.zeros((n_blobs, 4 << D), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.pass_lde(d_coef.data_ptr(), p, beta, n_blobs, n_felts, d_eval.data_ptr())
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    got = d_eval.cpu().numpy().view(np.uint32).reshape(n_blobs, 4, 1 << D)
    for b in range(n_blobs):
        for c in range(4):
            want = O.circle_fft(coef[b, c * n4:(c + 1) * n4], D)
            assert first_diff(got[b, c], want) is None, (b, c, first_diff(got[b, c], want))


def test_fold_pass_vs_oracle_c2_shape(ctx, torch_mod):
    # frieda_pass_fold (the kernel behind bench.py's fold roofline figures) against the oracle's FRI layers at
    # C2's shape: circle fold 2^18 -> 2^17 with alpha_0, then every line fold down to 2^5, two blobs per launch
    torch = torch_mod
    cfg = O.make_config(4, 0, 20, 20)
    traces = [O.trace(O.splitmix64_bytes(0x4652494544410000 + b, 131072), b, cfg, stop_after_fri=True,
                      with_trees=False) for b in range(2)]
    n_layers = len(traces[0].layer_logs)
    assert traces[0].layer_logs[0] == 18
    stream = torch.cuda.ExternalStream(ctx.stream_ptr)
    for layer in range(n_layers):
        lg = traces[0].layer_logs[layer]
        src = np.stack([t.layer_columns[layer] for t in traces])                        # (2, 4, 2^lg)
        want = np.stack([t.layer_columns[layer + 1] if layer + 1 < n_layers else t.last_eval for t in traces])
        alpha = np.array([t.alphas[layer] for t in traces], dtype=np.uint32)           # (2, 4)
        d_src = torch.from_numpy(src.view(np.int32)).cuda()
        d_alpha = torch.from_numpy(alpha.view(np.int32)).cuda()
        d_dst = torch.zeros((2, 4, 1 << (lg - 1)), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        ctx.pass_fold(d_src.data_ptr(), lg, layer == 0, 2, d_alpha.data_ptr(), d_dst.data_ptr())
        stream.synchronize()
        got = d_dst.cpu().numpy().view(np.uint32)
        assert first_diff(got, want) is None, (layer, first_diff(got, want))


@pytest.mark.parametrize("blob_len,stride,n,offset", [
    (131072, 131072, 3, 0),      # C2 shape: the staged 128-bit kernel
    (131072, 131088, 2, 0),      # padded stride, still 16-byte aligned
    (131071, 131075, 3, 1),      # odd length, odd stride, misaligned base: the byte path
    (1 << 20, 1 << 20, 1, 0),    # poly_log 17
    (15361, 15376, 2, 0),        # one full staging chunk + 1 byte
    (1000, 1000, 5, 0), (30, 30, 2, 0), (1, 1, 3, 0), (15, 19, 2, 3),
])
def test_pack_pass_vs_oracle(ctx, torch_mod, blob_len, stride, n, offset):
    # frieda_pass_pack against bytes_to_felt_le / polynomial_from_bytes (src/utils.rs:10-33)
    torch = torch_mod
    rng = np.random.default_rng(blob_len + stride)
    buf = rng.integers(0, 256, offset + n * stride + 16, dtype=np.uint8)
    buf[offset + blob_len - 1::stride][:n] |= 0x80  # top bits of the last byte set: the last limb is partial
    p = O.poly_log(blob_len)
    d_buf = torch.from_numpy(buf).cuda()
    d_coef = torch.full((n, 4 << p), -1, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.pass_pack(d_buf.data_ptr() + offset, blob_len, stride, n, d_coef.data_ptr())
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    got = d_coef.cpu().numpy().view(np.uint32)
    for b in range(n):
        felts = O.bytes_to_felts(buf[offset + b * stride: offset + b * stride + blob_len].tobytes())
        want = np.zeros(4 << p, dtype=np.uint32)
        want[:len(felts)] = felts
        assert first_diff(got[b], want) is None, (b, first_diff(got[b], want))


def test_lde_linearity_full_size(ctx, torch_mod):
    # size-independent property at C2's full size: LDE(a + b) == LDE(a) + LDE(b) (mod P)
    torch = torch_mod
    rng = np.random.default_rng(77)
    p, beta = 14, 4
    a = rng.integers(0, P, 4 << p, dtype=np.uint32)
    b = rng.integers(0, P, 4 << p, dtype=np.uint32)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    coef = np.stack([a, b, s])
    d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
    d_eval = torch.zeros((3, 4 << (p + beta)), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.pass_lde(d_coef.data_ptr(), p, beta, 3, 4 << p, d_eval.data_ptr())
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    ev = d_eval.cpu().numpy().view(np.uint32).astype(np.uint64)
    assert np.array_equal((ev[0] + ev[1]) % P, ev[2])
    assert ev.max() < P


def oracle_merkle_root(cols_b):
    """Root of the tree over one blob's (4, n) columns with the oracle's compression (hash_node)."""
    n = cols_b.shape[1]
    level = [O.blake2s_compress([0] * 8, [int(cols_b[c, i]) for c in range(4)] + [0] * 12) for i in range(n)]
    while len(level) > 1:
        level = [O.blake2s_compress([0] * 8, list(level[2 * i]) + list(level[2 * i + 1])) for i in range(len(level) // 2)]
    return np.array(level[0], dtype="<u4").tobytes()


@pytest.mark.parametrize("log,n_blobs", [(12, 3), (14, 1), (15, 2), (16, 1)])
def test_merkle_pass_latency_form_shapes(ctx, torch_mod, log, n_blobs):
    # a few blobs = the latency form of the Merkle passes (merkle.cu): chunks of 128 / 256 leaves in 128-thread CTAs
    # (in-place reduction by rounds), 512-leaf chunks in 256-thread CTAs; truncated and kept trees must give the
    # oracle's root
    torch = torch_mod
    rng = np.random.default_rng(1000 + log)
    cols = rng.integers(0, P, (n_blobs, 4, 1 << log), dtype=np.uint32)
    d_cols = torch.from_numpy(cols.view(np.int32)).cuda()
    want = [oracle_merkle_root(cols[b]) for b in range(n_blobs)]
    for keep in (False, True):
        d_roots = torch.zeros((n_blobs, 32), dtype=torch.uint8, device="cuda")
        d_tree = torch.zeros((n_blobs, 2 << log, 32), dtype=torch.uint8, device="cuda") if keep else None
        torch.cuda.synchronize()
        ctx.pass_merkle(d_cols.data_ptr(), log, n_blobs, d_tree.data_ptr() if keep else None, d_roots.data_ptr())
        torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
        roots = d_roots.cpu().numpy()
        for b in range(n_blobs):
            assert roots[b].tobytes() == want[b], (log, b, keep)
            if keep:
                assert d_tree[b, 1].cpu().numpy().tobytes() == want[b]


@pytest.mark.parametrize("log", [0, 1, 2, 5, 9, 10, 11, 13, 21])
def test_merkle_pass_vs_oracle(ctx, torch_mod, log):
    torch = torch_mod
    rng = np.random.default_rng(log)
    n_blobs = 2 if log < 20 else 1
    cols = rng.integers(0, P, (n_blobs, 4, 1 << log), dtype=np.uint32)
    d_cols = torch.from_numpy(cols.view(np.int32)).cuda()
    d_roots = torch.zeros((n_blobs, 32), dtype=torch.uint8, device="cuda")
    keep = log <= 13
    d_tree = torch.zeros((n_blobs, 2 << log, 32), dtype=torch.uint8, device="cuda") if keep else None
    torch.cuda.synchronize()
    ctx.pass_merkle(d_cols.data_ptr(), log, n_blobs, d_tree.data_ptr() if keep else None, d_roots.data_ptr())
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    roots = d_roots.cpu().numpy()
    # oracle tree through hashlib-free restatement: reuse oracle's compress
    for b in range(n_blobs):
        level = [None] * (log + 1)
        leaves = np.zeros((1 << log, 8), dtype=np.uint32)
        if log <= 13:
            for i in range(1 << log):
                m = [int(cols[b, c, i]) for c in range(4)] + [0] * 12
                leaves[i] = O.blake2s_compress([0] * 8, m)
            level[log] = leaves
            for k in range(log - 1, -1, -1):
                cur = np.zeros((1 << k, 8), dtype=np.uint32)
                for i in range(1 << k):
                    m = [int(x) for x in level[k + 1][2 * i]] + [int(x) for x in level[k + 1][2 * i + 1]]
                    cur[i] = O.blake2s_compress([0] * 8, m)
                level[k] = cur
            assert roots[b].tobytes() == level[0].astype("<u4").tobytes()
            tree = d_tree[b].cpu().numpy()
            for k in range(log + 1):
                got = tree[(1 << k):(2 << k)].reshape(-1)
                want = np.frombuffer(level[k].astype("<u4").tobytes(), dtype=np.uint8)
                assert first_diff(got, want) is None, (k, first_diff(got, want))
    if log == 21:
        # large tree: root only, against the oracle's own tree over the same columns via a commit-shaped check
        # (idempotence: two runs agree, and the kept-tree variant agrees with the truncated one)
        d_tree2 = torch.zeros((1, 2 << log, 32), dtype=torch.uint8, device="cuda")
        d_roots2 = torch.zeros((1, 32), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ctx.pass_merkle(d_cols.data_ptr(), log, 1, d_tree2.data_ptr(), d_roots2.data_ptr())
        torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
        assert d_roots2.cpu().numpy().tobytes() == roots.tobytes()
        t = d_tree2[0].cpu().numpy()
        assert t[1].tobytes() == roots[0].tobytes()
        # spot-check one path from a leaf to the root with the oracle's compression
        i = 123457
        m = [int(cols[0, c, i]) for c in range(4)] + [0] * 12
        h = np.array(O.blake2s_compress([0] * 8, m), dtype="<u4").tobytes()
        assert t[(1 << log) + i].tobytes() == h
        for k in range(log - 1, -1, -1):
            j = i >> (log - k)
            l, r = t[(2 << k) + 2 * j], t[(2 << k) + 2 * j + 1]
            m = list(np.frombuffer(l.tobytes() + r.tobytes(), dtype="<u4"))
            h = np.array(O.blake2s_compress([0] * 8, [int(x) for x in m]), dtype="<u4").tobytes()
            assert t[(1 << k) + j].tobytes() == h, k


# ------------------------------------------------------------------ FRI commit phase, every intermediate
CASES = [
    ("pattern", 1024, 1024, (4, 0, 20, 20)),
    ("pattern", 3000, 5, (2, 2, 33, 8)),
    ("pattern", 65536, 65536, (4, 0, 20, 20)),
    ("e2e", 58, None, (4, 0, 20, 20)),
    ("pattern", 200, 3, (1, 0, 5, 3)),
    ("pattern", 40000, None, (3, 1, 17, 10)),
    # above one shared-memory LDE block: poly_log 17 (one strided radix-16 pass) and poly_log 20 (two), D = 19 / 22
    # (Merkle middle passes, kept-tree strides), and poly_log 17 at blowup 2^4 (D = 21) with an 8-coefficient last layer
    ("splitmix", 1 << 20, 11, (2, 0, 20, 12)),
    ("splitmix", 8 << 20, None, (2, 0, 20, 12)),
    ("splitmix7", 1 << 20, 3, (4, 3, 64, 10)),
]


def case_data(kind, n):
    if kind == "e2e":
        return b"This is the original data that needs to be made available."
    if kind == "splitmix":
        return O.splitmix64_bytes(0x4652494544414236, n)
    if kind == "splitmix7":
        return O.splitmix64_bytes(0x4652494544414237, n)
    return pattern(n)


@pytest.mark.parametrize("kind,n,seed,cfg", CASES)
def test_fri_commit_every_intermediate(ctx, kind, n, seed, cfg):
    data = case_data(kind, n)
    ocfg = O.make_config(*cfg)
    t = O.trace(data, seed, ocfg, stop_after_fri=True, with_trees=True)
    blobs = np.frombuffer(data, dtype=np.uint8).reshape(1, -1).copy()
    ctx.set_debug_keep(True)
    try:
        roots, last = ctx.fri_commit_batch(blobs, None if seed is None else [seed], F.PcsConfig(*cfg))
        n_layers = len(t.layer_logs)
        assert roots.shape[1] == n_layers
        coef = ctx.debug_fetch(0, 0, 0, 0, 16 << t.poly_log).view(np.uint32)
        assert first_diff(coef, t.coeffs) is None, "coefficients: " + str(first_diff(coef, t.coeffs))
        for layer in range(n_layers):
            lg = t.layer_logs[layer]
            cols = ctx.debug_fetch(1, 0, layer, 0, 16 << lg).view(np.uint32).reshape(4, -1)
            d = first_diff(cols, t.layer_columns[layer])
            assert d is None, f"layer {layer} columns: {d}"
            for level in range(lg, -1, -1):
                got = ctx.debug_fetch(2, 0, layer, level, 32 << level)
                d = first_diff(got, t.tree_levels[layer][level])
                assert d is None, f"layer {layer} tree level {level}: {d}"
            assert roots[0, layer].tobytes() == t.tree_levels[layer][0].tobytes(), f"root of layer {layer}"
            alpha = tuple(int(x) for x in ctx.debug_fetch(3, 0, layer, 0, 16).view(np.uint32))
            assert alpha == t.alphas[layer], f"alpha of layer {layer}"
        last_cols = ctx.debug_fetch(1, 0, n_layers, 0, 16 << (cfg[0] + cfg[1])).view(np.uint32).reshape(4, -1)
        assert first_diff(last_cols, t.last_eval) is None, "last evaluation"
        assert [tuple(int(x) for x in q) for q in last[0]] == t.last_layer_poly
        assert ctx.debug_fetch(4, 0, 0, 0, 32).tobytes() == t.digest_after_fri
    finally:
        ctx.set_debug_keep(False)


def test_fri_commit_blob_layers(ctx, blob_bytes, golden):
    g = next(x for x in golden["oracle_generated"]["prove"] if x["name"] == "blob_4_1_20_20")
    blobs = np.frombuffer(blob_bytes, dtype=np.uint8).reshape(1, -1).copy()
    roots, last = ctx.fri_commit_batch(blobs, None, F.PcsConfig(4, 1, 20, 20))
    assert [r.tobytes().hex() for r in roots[0]] == g["layer_roots"]
    assert [[int(x) for x in q] for q in last[0]] == g["last_layer_poly"]
    assert roots[0, 0].tobytes().hex() == golden["reference_certified"]["commit_blob_blowup4"]


def test_fri_commit_batch_with_seeds_vs_oracle(ctx):
    cfg = (4, 0, 20, 20)
    n, blob_len = 6, 131072
    blobs = np.stack([np.frombuffer(O.splitmix64_bytes(0x4652494544410000 + b, blob_len), dtype=np.uint8)
                      for b in range(n)])
    seeds = list(range(n))
    roots, last = ctx.fri_commit_batch(blobs, seeds, F.PcsConfig(*cfg))
    for b in range(n):
        oroots, olast = O.fri_commit(blobs[b].tobytes(), seeds[b], O.make_config(*cfg))
        assert [r.tobytes() for r in roots[b]] == oroots, b
        assert [tuple(int(x) for x in q) for q in last[b]] == olast, b


def test_fri_commit_device_pointers(ctx, torch_mod):
    torch = torch_mod
    cfg = F.PcsConfig(4, 0, 20, 20)
    n, blob_len = 4, 131072
    blobs = np.stack([np.frombuffer(O.splitmix64_bytes(0x4652494544410000 + b, blob_len), dtype=np.uint8)
                      for b in range(n)])
    L = 1 + ctx.n_inner_layers(blob_len, cfg)
    d_in = torch.from_numpy(blobs).cuda()
    d_roots = torch.zeros((n, L, 32), dtype=torch.uint8, device="cuda")
    d_last = torch.zeros((n, 1, 4), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.fri_commit_batch_ptr(d_in.data_ptr(), blob_len, blob_len, n, None, cfg, d_roots.data_ptr(),
                             d_last.data_ptr(), device=True)
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    roots, last = ctx.fri_commit_batch(blobs, None, cfg)
    assert first_diff(d_roots.cpu().numpy(), roots) is None
    assert first_diff(d_last.cpu().numpy().view(np.uint32), last) is None


def test_device_entry_point_reports_invalid_degree_through_take_error(ctx, torch_mod):
    # stwo asserts "invalid degree" inside FriProver::commit; the asynchronous device entry point cannot report it when
    # it returns, frieda_ctx_take_error does.  No input of the public API can trigger the assert (the evaluation IS the
    # LDE of a polynomial), so this checks the plumbing: take_error synchronises the stream, returns OK after clean
    # device calls, and can be called repeatedly.
    torch = torch_mod
    cfg = F.PcsConfig(4, 0, 20, 20)
    blobs = np.stack([np.frombuffer(O.splitmix64_bytes(0x4652494544410000 + b, 4096), dtype=np.uint8) for b in range(3)])
    L = 1 + ctx.n_inner_layers(4096, cfg)
    d_in = torch.from_numpy(blobs).cuda()
    d_roots = torch.zeros((3, L, 32), dtype=torch.uint8, device="cuda")
    d_last = torch.zeros((3, 1, 4), dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.fri_commit_batch_ptr(d_in.data_ptr(), 4096, 4096, 3, None, cfg, d_roots.data_ptr(), d_last.data_ptr(),
                             device=True)
    ctx.take_error()  # synchronises: results are complete without any other sync
    oroots, _ = O.fri_commit(blobs[2].tobytes(), None, O.make_config(4, 0, 20, 20))
    assert [r.tobytes() for r in d_roots[2].cpu().numpy()] == oroots
    ctx.take_error()


def test_fri_reference_panic_shapes(ctx):
    with pytest.raises(F.ReferencePanic):
        ctx.fri_commit_batch(np.zeros((1, 2), dtype=np.uint8), None, F.PcsConfig(4, 0, 20, 4))


# ------------------------------------------------------------------ proofs
def check_proof_vs_oracle(ctx, data, seed, cfg):
    ocfg = O.make_config(*cfg)
    t = O.trace(data, seed, ocfg, with_trees=False)
    root, pr = ctx.commit_and_generate_proof(data, seed, F.PcsConfig(*cfg))
    assert root == t.root
    assert pr.proof_of_work == t.nonce, f"nonce {pr.proof_of_work} vs {t.nonce}"
    nq = len(t.queries)
    got_q = ctx.debug_fetch(6, 0, 0, 0, 4 * cfg[2]).view(np.uint32)
    assert first_diff(got_q, t.queries) is None, "queries"
    assert len(pr.evaluations) == nq
    assert pr.last_layer_poly == t.last_layer_poly
    assert pr.serialize() == t.proof_bytes, "serialized proof differs from the oracle's"
    return pr


@pytest.mark.parametrize("kind,n,seed,cfg", CASES)
def test_proof_bit_exact_vs_oracle(ctx, kind, n, seed, cfg):
    data = case_data(kind, n)
    pr = check_proof_vs_oracle(ctx, data, seed, cfg)
    assert F.verify_proof(pr, seed)
    assert not F.verify_proof(pr, (seed or 0) + 1)


def test_proof_golden_vector_file(ctx, blob_bytes, golden):
    for g in golden["oracle_generated"]["prove"]:
        name = g["name"]
        if name.startswith("blob"):
            data = blob_bytes
        elif name == "e2e_string":
            data = b"This is the original data that needs to be made available."
        elif name.startswith("splitmix_c4_b"):
            data = O.splitmix64_bytes(0x4652494544410000 + int(name.rsplit("b", 1)[1]), g["len"])
        elif name.startswith("splitmix_big"):
            data = O.splitmix64_bytes(0x4652494544414237 if name.endswith("l3") else 0x4652494544414236, g["len"])
        else:
            data = pattern(g["len"])
        root, pr = ctx.commit_and_generate_proof(data, g["seed"], F.PcsConfig(*g["cfg"]))
        assert root.hex() == g["root"], name
        assert pr.proof_of_work == g["nonce"], name
        assert [list(q) for q in pr.last_layer_poly] == g["last_layer_poly"], name
        assert [pr.first_layer_commitment.hex()] + [c.hex() for c in pr.inner_layer_commitments] == g[
            "layer_roots"], name
        b = pr.serialize()
        assert len(b) == g["proof_len"] and hashlib.sha256(b).hexdigest() == g["proof_sha256"], name
        assert F.verify_proof(pr, g["seed"]), name


PCS = (4, 1, 20, 20)  # src/proof.rs:109-116


@pytest.fixture(scope="module")
def blob_proof(ctx, blob_bytes):
    return ctx.commit_and_generate_proof(blob_bytes, None, F.PcsConfig(*PCS))


def test_generate_proof(blob_proof):
    assert blob_proof[1].n_inner_layers != 0                      # src/proof.rs:119-124


def test_commit_and_generate_proof(ctx, blob_bytes, blob_proof):
    root, pr = blob_proof                                         # src/proof.rs:126-135
    assert root == ctx.commit(blob_bytes, PCS[0])
    assert pr.first_layer_commitment == root


def test_verify_proof(blob_proof):
    assert F.verify_proof(blob_proof[1], None)                    # src/proof.rs:136-141


def test_verify_proof_with_invalid_pow(blob_proof):
    p = blob_proof[1].clone()
    p.proof_of_work += 1
    assert not F.verify_proof(p, None)                            # :143-149


def test_verify_proof_with_invalid_evaluations(blob_proof):
    p = blob_proof[1].clone()
    p.set_evaluation(0, [(x + 1) % P for x in p.evaluations[0]])
    assert not F.verify_proof(p, None)                            # :151-157


def test_verify_proof_with_invalid_evaluations_order(blob_proof):
    p = blob_proof[1].clone()
    ev = p.evaluations
    for i, e in enumerate(reversed(ev)):
        p.set_evaluation(i, e)
    assert not F.verify_proof(p, None)                            # :158-164


def test_verify_proof_with_invalid_evaluations_length(blob_proof):
    p = blob_proof[1].clone()
    p.pop_evaluation()
    with pytest.raises(F.ReferencePanic):                         # :166-173 (#[should_panic])
        F.verify_proof(p, None)
    p.c.n_evaluations += 1


def test_verify_proof_with_invalid_1_evaluation_unordered(blob_proof):
    p = blob_proof[1].clone()
    ev = p.evaluations
    p.set_evaluation(0, ev[1])
    p.set_evaluation(1, ev[0])
    assert not F.verify_proof(p, None)                            # :175-181


def test_verify_proof_with_seed(ctx, blob_bytes):
    cfg = F.PcsConfig(*PCS)                                       # src/proof.rs:183-193
    p1 = ctx.generate_proof(blob_bytes, 1, cfg)
    p2 = ctx.generate_proof(blob_bytes, 2, cfg)
    assert p1.evaluations != p2.evaluations
    assert F.verify_proof(p1, 1) and F.verify_proof(p2, 2)
    assert not F.verify_proof(p1, 2) and not F.verify_proof(p2, 1)


def test_end_to_end(ctx):
    data = b"This is the original data that needs to be made available."   # src/lib.rs:52-85
    cfg = F.PcsConfig(4, 0, 20, 20)
    commitment = ctx.commit(data, 4)
    proof = ctx.generate_proof(data, None, cfg)
    assert proof.first_layer_commitment == commitment
    assert F.verify_proof(proof, None)


def test_oracle_verifier_accepts_gpu_proof(ctx, blob_bytes):
    cfg = (4, 1, 20, 20)
    _, pr = ctx.commit_and_generate_proof(blob_bytes, 3, F.PcsConfig(*cfg))
    _, opr = O.prove(blob_bytes, 3, O.make_config(*cfg))
    assert pr.serialize() == opr.serialize()
    assert O.verify(opr, 3)


def test_prove_batch_c4_vs_oracle(ctx):
    # BASELINE config 4: 64 queries per blob, seed = blob index, batched
    cfg = (4, 0, 64, 20)
    n, blob_len = 8, 131072
    blobs = np.stack([np.frombuffer(O.splitmix64_bytes(0x4652494544410000 + b, blob_len), dtype=np.uint8)
                      for b in range(n)])
    seeds = list(range(n))
    roots, proofs = ctx.prove_batch(blobs, seeds, F.PcsConfig(*cfg))
    for b in (0, 3, 7):
        oroot, opr = O.prove(blobs[b].tobytes(), seeds[b], O.make_config(*cfg))
        assert roots[b].tobytes() == oroot
        assert proofs[b].serialize() == opr.serialize(), b
    for b in range(n):
        assert F.verify_proof(proofs[b], seeds[b])
        assert not F.verify_proof(proofs[b], seeds[b] + 1)


def test_prove_batch_large_blobs_vs_oracle(ctx):
    # 4 blobs of 1 MiB (poly_log 17, blowup 2^2): the strided-LDE -> FRI -> decommit combination, batched;
    # every proof byte-exact against the oracle, then the same batch cut into waves of one blob
    cfg = (2, 0, 24, 10)
    n, blob_len = 4, 1 << 20
    blobs = np.stack([np.frombuffer(O.splitmix64_bytes(0x4652494544414236 + b, blob_len), dtype=np.uint8)
                      for b in range(n)])
    seeds = [40 + b for b in range(n)]
    want = [O.prove(blobs[b].tobytes(), seeds[b], O.make_config(*cfg)) for b in range(n)]
    roots, proofs = ctx.prove_batch(blobs, seeds, F.PcsConfig(*cfg))
    for b in range(n):
        assert roots[b].tobytes() == want[b][0], b
        assert proofs[b].serialize() == want[b][1].serialize(), b
        assert F.verify_proof(proofs[b], seeds[b]) and not F.verify_proof(proofs[b], seeds[b] + 1)
    layer_roots, last = ctx.fri_commit_batch(blobs, seeds, F.PcsConfig(*cfg))
    for b in range(n):
        oroots, olast = O.fri_commit(blobs[b].tobytes(), seeds[b], O.make_config(*cfg))
        assert [r.tobytes() for r in layer_roots[b]] == oroots, b
        assert [tuple(int(x) for x in q) for q in last[b]] == olast, b
    ctx.set_workspace_limit(400 << 20)  # one blob per wave
    try:
        roots2, proofs2 = ctx.prove_batch(blobs, seeds, F.PcsConfig(*cfg))
    finally:
        ctx.set_workspace_limit(0)
    assert np.array_equal(roots, roots2)
    assert [p.serialize() for p in proofs2] == [p.serialize() for p in proofs]


def test_prove_batch_in_waves(ctx):
    cfg = (3, 0, 9, 6)
    rng = np.random.default_rng(11)
    n, blob_len = 9, 3000
    blobs = rng.integers(0, 256, (n, blob_len), dtype=np.uint8)
    seeds = [100 + i for i in range(n)]
    ctx.set_workspace_limit(3 * (1 << 20))
    try:
        roots, proofs = ctx.prove_batch(blobs, seeds, F.PcsConfig(*cfg))
    finally:
        ctx.set_workspace_limit(0)
    for b in range(n):
        oroot, opr = O.prove(blobs[b].tobytes(), seeds[b], O.make_config(*cfg))
        assert roots[b].tobytes() == oroot and proofs[b].serialize() == opr.serialize(), b


# ------------------------------------------------------------------ split blob (BASELINE config 5)
@pytest.mark.parametrize("n_bytes,blow,worlds", [
    (1 << 20, 2, (1, 2, 4, 8)),      # poly_log 17: strided LDE passes; world 8 > 2^blowup -> partial blocks
    (131072, 1, (2, 4)),             # shared-memory LDE path with a partial block
    (131072, 4, (1, 2, 8, 16)),      # whole blocks per rank
    (3000, 3, (2, 16)),
])
def test_commit_split_virtual_ranks(ctx, torch_mod, n_bytes, blow, worlds):
    # sharding logic on ONE GPU: G virtual ranks run sequentially, subtree roots are combined,
    # and the result must equal the unsplit commit (and the oracle)
    torch = torch_mod
    data = O.splitmix64_bytes(0x4652494544414236, n_bytes)
    want = O.commit(data, blow)
    assert ctx.commit(data, blow) == want
    for world in worlds:
        subs = torch.zeros((world, 32), dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        for r in range(world):
            ctx.commit_split_local(data, blow, r, world, subs[r].data_ptr())
        assert ctx.merkle_combine(subs.data_ptr(), world) == want, world


@pytest.mark.parametrize("n_bytes,blow,worlds", [
    (1 << 20, 2, (2, 8)), (131072, 4, (1, 4, 16)), (3000, 3, (2, 16)), (100003, 2, (4, 8)), (17, 3, (2,)),
])
def test_commit_split_peer_slices_virtual_ranks(ctx, torch_mod, n_bytes, blow, worlds):
    # the peer-memory path on ONE GPU: every "peer" slice is a separate device buffer, the packing kernel reads the
    # slices in place and the combine kernel reads each root where its rank left it (frieda_b200/parallel.py)
    torch = torch_mod
    from frieda_b200.parallel import slice_bounds
    data = O.splitmix64_bytes(0x4652494544414236, n_bytes)
    want = O.commit(data, blow)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr)
    for world in worlds:
        per = slice_bounds(n_bytes, 0, world)[1]
        slices = []
        for r in range(world):
            lo, hi = slice_bounds(n_bytes, r, world)
            buf = torch.full((per,), 0xA5, dtype=torch.uint8, device="cuda")  # bytes past `len` must not matter
            if hi > lo:
                buf[: hi - lo] = torch.frombuffer(bytearray(data[lo:hi]), dtype=torch.uint8).cuda()
            slices.append(buf)
        roots = [torch.zeros(32, dtype=torch.uint8, device="cuda") for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            ctx.commit_split_local_peers([b.data_ptr() for b in slices], per, n_bytes, blow, r, roots[r].data_ptr())
        stream.synchronize()
        assert ctx.merkle_combine_peers([t.data_ptr() for t in roots]) == want, world


def test_commit_split_peers_single_rank(ctx, torch_mod):
    # frieda_commit_split_peers with world = 1: upload, barrier kernels (trivially satisfied), pack from the
    # slice, subtree, combine -- the one-call path the multi-GPU host logic uses
    torch = torch_mod
    data = O.splitmix64_bytes(0x4652494544414236, 300007)
    per = (len(data) + 15) // 16 * 16
    sl = torch.empty(per, dtype=torch.uint8, device="cuda")
    roots = torch.zeros(64 * 32, dtype=torch.uint8, device="cuda")
    flags = torch.zeros(512, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for epoch in (1, 2, 3):
        got = ctx.commit_split_peers(data, 3, 0, [sl.data_ptr()], per, [roots.data_ptr()], [flags.data_ptr()], epoch)
        assert got == O.commit(data, 3)
    assert flags.cpu().tolist()[0] == 3 and flags.cpu().tolist()[64] == 3
    # a slice of 1 MiB and more is uploaded in four parts on the copy stream, the packing kernel waiting per part
    big = O.splitmix64_bytes(0x4652494544414236, (3 << 20) + 77)
    per = (len(big) + 15) // 16 * 16
    sl = torch.empty(per, dtype=torch.uint8, device="cuda")
    for epoch in (4, 5):
        assert ctx.commit_split_peers(big, 2, 0, [sl.data_ptr()], per, [roots.data_ptr()], [flags.data_ptr()],
                                      epoch) == O.commit(big, 2)
    # data = NULL: the slice is already in the buffer (inputs resident in HBM), nothing is uploaded
    assert ctx.commit_split_peers(None, 2, 0, [sl.data_ptr()], per, [roots.data_ptr()], [flags.data_ptr()], 6,
                                  resident_len=len(big)) == O.commit(big, 2)
    f = flags.cpu().tolist()
    assert f[0] == 3 and f[64] == 6 and [f[64 * c] for c in (2, 3, 4, 5)] == [6, 6, 6, 6]


def test_commit_split_single_process_api(ctx):
    from frieda_b200.parallel import commit_split
    data = pattern(50000)
    assert commit_split(ctx, data, 3) == O.commit(data, 3)


def test_commit_split_subroots_are_tree_nodes(ctx, torch_mod):
    torch = torch_mod
    data = pattern(20000)
    t = O.trace(data, None, O.make_config(2, 0, 4, 1), stop_after_fri=True, with_trees=True)
    world = 4
    subs = torch.zeros((world, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    for r in range(world):
        ctx.commit_split_local(data, 2, r, world, subs[r].data_ptr())
    assert subs.cpu().numpy().tobytes() == t.tree_levels[0][2].tobytes()


def test_commit_config5_64MiB_split_and_unsplit(ctx, torch_mod, golden):
    # BASELINE config 5 at full size: one 64 MiB blob, blowup 2^2 (poly_log 23: strided LDE passes);
    # root from the oracle (tests/golden/make_vectors.py), unsplit and split over 8 virtual ranks
    torch = torch_mod
    for g in golden["oracle_generated"]["commit_large"]:
        data = O.splitmix64_bytes(int(g["state0"], 16), g["len"])
        assert ctx.commit(data, g["log_blowup"]).hex() == g["root"], g["name"]
        for world in (2, 8):
            subs = torch.zeros((world, 32), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            for r in range(world):
                ctx.commit_split_local(data, g["log_blowup"], r, world, subs[r].data_ptr())
            assert ctx.merkle_combine(subs.data_ptr(), world).hex() == g["root"], (g["name"], world)


# ------------------------------------------------------------------ split-blob FRI commit phase
@pytest.mark.parametrize("kind,n_bytes,seed,cfg,worlds", [
    ("splitmix", 1 << 20, 11, (2, 0, 20, 12), (1, 2, 4, 8)),   # poly_log 17: strided LDE on a partial range
    ("splitmix", 8 << 20, None, (2, 0, 20, 12), (2, 8)),       # poly_log 20
    ("splitmix", 131072, 3, (4, 0, 20, 20), (2, 4, 16)),       # C2 shape
    ("blob", 262146, None, (4, 1, 20, 20), (4,)),              # the reference's blob and config (src/proof.rs:109-116)
    ("splitmix", 131072, 9, (4, 5, 20, 20), (1, 2)),           # world 1: every committed layer is split, the tail
                                                               # starts from the gathered last evaluation
    ("pattern", 40000, 5, (3, 1, 17, 10), (2, 8)),
])
def test_fri_commit_split_virtual_ranks(torch_mod, blob_bytes, kind, n_bytes, seed, cfg, worlds):
    # G contexts on ONE GPU play the ranks in lockstep; the exchanges are plain device copies.  Layer roots and the
    # last-layer polynomial of EVERY rank must equal the oracle's FriProver::commit of the whole blob.
    torch = torch_mod
    data = blob_bytes if kind == "blob" else case_data(kind, n_bytes)
    oroots, olast = O.fri_commit(data, seed, O.make_config(*cfg))
    pcs = F.PcsConfig(*cfg)
    for world in worlds:
        ctxs = [F.Context(0) for _ in range(world)]
        try:
            shapes = [c.fri_split_begin(data, seed, pcs, r, world) for r, c in enumerate(ctxs)]
            assert len(set(shapes)) == 1
            n_split, n_layers, handoff_log = shapes[0]
            assert n_layers == len(oroots) and 1 <= n_split <= n_layers
            for layer in range(n_split):
                subs = torch.zeros((world, 32), dtype=torch.uint8, device="cuda")
                torch.cuda.synchronize()  # the fill runs on torch's stream, the library on its own non-blocking ones
                for r, c in enumerate(ctxs):
                    c.fri_split_layer(layer, subs[r].data_ptr())
                torch.cuda.synchronize()
                for c in ctxs:
                    c.fri_split_combine(layer, subs.data_ptr())
                torch.cuda.synchronize()
            cols = torch.zeros((world, 4 << handoff_log), dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            for r, c in enumerate(ctxs):
                c.fri_split_handoff(cols[r].data_ptr())
            torch.cuda.synchronize()
            for r, c in enumerate(ctxs):
                roots, last = c.fri_split_finish(cols.data_ptr(), n_layers, cfg[1])
                assert [x.tobytes() for x in roots] == oroots, (world, r)
                assert [tuple(int(x) for x in q) for q in last] == olast, (world, r)
        finally:
            for c in ctxs:
                c.close()


PEER_LAYERS_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch
import frieda_b200 as F
from frieda_b200 import api
from oracle import oracle as O
# `world` contexts on ONE GPU play the ranks; plain device buffers stand in for the peer-mapped roots areas and flag
# arrays.  Every rank's whole layer loop is enqueued (asynchronously) before anyone synchronises, and the ranks'
# barrier kernels then meet on the GPU.
for n_bytes, cfg, seed, worlds in ((131072, (4, 0, 20, 12), 3, (2, 4)), (1 << 20, (2, 1, 20, 10), None, (4,)),
                                   (40000, (3, 1, 17, 10), 5, (2,))):
    data = O.splitmix64_bytes(0x4652494544414236 + n_bytes, n_bytes)
    oroots, olast = O.fri_commit(data, seed, O.make_config(*cfg))
    oroot, oproof = O.prove(data, seed, O.make_config(*cfg))
    pcs = F.PcsConfig(*cfg)
    for world in worlds:
        ctxs = [F.Context(0) for _ in range(world)]
        areas = torch.zeros((world, 64 * 32), dtype=torch.uint8, device="cuda")
        flags = torch.zeros((world, 512), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        root_ptrs = [areas[r].data_ptr() for r in range(world)]
        flag_ptrs = [flags[r].data_ptr() for r in range(world)]
        # (frieda_fri_split_begin_peers is left to the two-rank test: its upload and the LDE's memsets use the copy
        # engine on both sides of a peer barrier, and on ONE GPU the virtual ranks share that engine's queue)
        for epoch, keep in ((1, False), (2, False), (3, True)):  # the slots are reused under an advancing epoch
            shapes = [c.fri_split_begin(data, seed, pcs, r, world, keep_trees=keep) for r, c in enumerate(ctxs)]
            n_split, n_layers, handoff_log = shapes[0]
            for c in ctxs:
                c.fri_split_layers_peers(root_ptrs, flag_ptrs, epoch)
            cols = torch.zeros((world, 4 << handoff_log), dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            for r, c in enumerate(ctxs):
                c.fri_split_handoff(cols[r].data_ptr())
            torch.cuda.synchronize()
            for r, c in enumerate(ctxs):
                roots, last = c.fri_split_finish(cols.data_ptr(), n_layers, cfg[1])
                assert [x.tobytes() for x in roots] == oroots, (n_bytes, world, r, epoch)
                assert [tuple(int(x) for x in q) for q in last] == olast, (n_bytes, world, r, epoch)
            if keep:
                proof = api.split_assemble([c.fri_split_decommit() for c in ctxs])
                assert proof.serialize() == oproof.serialize(), (n_bytes, world)
        for c in ctxs:
            c.close()
print("PEER_LAYERS_OK")
"""


def test_fri_split_layers_peers_virtual_ranks(tmp_path):
    # frieda_fri_split_layers_peers on ONE GPU: separate process so that CUDA_DEVICE_MAX_CONNECTIONS can give every
    # virtual rank's stream its own hardware queue (ranks that shared one would wait on each other's barrier kernels)
    import subprocess
    import sys
    if os.environ.get("CUDA_LAUNCH_BLOCKING") == "1":
        pytest.skip("needs kernels of different streams to run concurrently")
    script = tmp_path / "peer_layers_worker.py"
    script.write_text(PEER_LAYERS_WORKER % {"root": ROOT})
    # ... and CUDA_MODULE_LOADING=EAGER: with lazy loading the FIRST launch of a kernel may synchronise the device,
    # which would block the host behind rank 0's spinning barrier before rank 1 is even enqueued (real ranks are
    # separate processes and only serialise on that first load)
    env = dict(os.environ, CUDA_DEVICE_MAX_CONNECTIONS="32", CUDA_MODULE_LOADING="EAGER")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and "PEER_LAYERS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def split_prove_virtual(torch, data, seed, cfg, world):
    """commit_and_generate_proof of one blob over `world` contexts on one GPU (lockstep, device copies as the
    exchange): returns the per-rank shares of the decommitment."""
    pcs = F.PcsConfig(*cfg)
    ctxs = [F.Context(0) for _ in range(world)]
    try:
        shapes = [c.fri_split_begin(data, seed, pcs, r, world, keep_trees=True) for r, c in enumerate(ctxs)]
        n_split, n_layers, handoff_log = shapes[0]
        for layer in range(n_split):
            subs = torch.zeros((world, 32), dtype=torch.uint8, device="cuda")
            torch.cuda.synchronize()
            for r, c in enumerate(ctxs):
                c.fri_split_layer(layer, subs[r].data_ptr())
            torch.cuda.synchronize()
            for c in ctxs:
                c.fri_split_combine(layer, subs.data_ptr())
            torch.cuda.synchronize()
        cols = torch.zeros((world, 4 << handoff_log), dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        for r, c in enumerate(ctxs):
            c.fri_split_handoff(cols[r].data_ptr())
        torch.cuda.synchronize()
        for c in ctxs:
            c.fri_split_finish(cols.data_ptr(), n_layers, cfg[1])
        return [c.fri_split_decommit() for c in ctxs]
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("kind,n_bytes,seed,cfg,worlds", [
    ("splitmix", 1 << 20, 11, (2, 0, 20, 12), (1, 2, 8)),      # poly_log 17
    ("splitmix", 131072, 3, (4, 0, 64, 16), (2, 4, 16)),       # C4 shape: 64 queries
    ("blob", 262146, None, (4, 1, 20, 20), (4,)),              # the reference's blob and config
    ("splitmix", 131072, 9, (4, 5, 20, 8), (1, 2)),            # every committed layer split (world 1)
    ("pattern", 40000, 5, (3, 1, 200, 6), (2, 8)),             # many queries: most tree paths merge early
    ("splitmix", 8 << 20, None, (2, 0, 20, 12), (8,)),         # poly_log 20
])
def test_prove_split_virtual_ranks_byte_exact(torch_mod, blob_bytes, kind, n_bytes, seed, cfg, worlds):
    # owner-serves-path decommitment: the merged proof must be the oracle's proof, byte for byte
    from frieda_b200 import api
    data = blob_bytes if kind == "blob" else case_data(kind, n_bytes)
    oroot, opr = O.prove(data, seed, O.make_config(*cfg))
    want = opr.serialize()
    for world in worlds:
        shares = split_prove_virtual(torch_mod, data, seed, cfg, world)
        proof = api.split_assemble(shares)
        assert proof.first_layer_commitment == oroot, world
        assert proof.serialize() == want, world
        assert F.verify_proof(proof, seed) and not F.verify_proof(proof, (seed or 0) + 1)
        if world > 1:
            # shares that do not belong together are refused, not merged into a wrong proof
            with pytest.raises(F.FriedaError):
                api.split_assemble(shares[::-1])
            with pytest.raises(F.FriedaError):
                api.split_assemble(shares[:-1] + [shares[-1][:-32]])
            with pytest.raises(F.FriedaError):
                api.split_assemble(shares[: world // 2])


def test_prove_split_driver_single_rank(ctx):
    from frieda_b200.parallel import prove_split
    data = O.splitmix64_bytes(0x4652494544414236, 300000)
    cfg = (3, 0, 20, 8)
    root, proof = prove_split(ctx, data, 21, F.PcsConfig(*cfg))
    oroot, opr = O.prove(data, 21, O.make_config(*cfg))
    assert root == oroot and proof.serialize() == opr.serialize()
    with pytest.raises(F.FriedaError):   # no kept trees: nothing to decommit from
        ctx.fri_split_begin(data, None, F.PcsConfig(*cfg), 0, 1)
        ctx.fri_split_decommit()


def test_fri_commit_split_driver_single_rank_and_errors(ctx):
    from frieda_b200.parallel import fri_commit_split
    data = O.splitmix64_bytes(0x4652494544414236, 300000)
    cfg = (3, 0, 20, 8)
    roots, last = fri_commit_split(ctx, data, 21, F.PcsConfig(*cfg))
    oroots, olast = O.fri_commit(data, 21, O.make_config(*cfg))
    assert [x.tobytes() for x in roots] == oroots and [tuple(int(x) for x in q) for q in last] == olast
    with pytest.raises(F.FriedaError):          # out of order
        ctx.fri_split_begin(data, None, F.PcsConfig(*cfg), 0, 1)
        ctx.fri_split_combine(0, 0)
    with pytest.raises(F.FriedaError):          # too small to split over 64 ranks
        ctx.fri_split_begin(pattern(2000), None, F.PcsConfig(2, 0, 8, 2), 0, 64)
    with pytest.raises(F.FriedaError):
        ctx.fri_split_layer(0, 1)               # no split in progress after the failed begin
    assert ctx.commit(b"abc", 3) == O.commit(b"abc", 3)


# ------------------------------------------------------------------ GPU batch verification
def test_verify_batch_matches_host_verifier(ctx):
    cfg = (4, 0, 20, 12)
    rng = np.random.default_rng(21)
    n, blob_len = 24, 20000
    blobs = rng.integers(0, 256, (n, blob_len), dtype=np.uint8)
    seeds = [1000 + i for i in range(n)]
    _, proofs = ctx.prove_batch(blobs, seeds, F.PcsConfig(*cfg))
    cases, case_seeds = [], []
    for i, p in enumerate(proofs):
        q = p.clone()
        kind = i % 8
        if kind == 1:
            q.proof_of_work += 1
        elif kind == 2:
            q.set_evaluation(0, [(x + 1) % P for x in q.evaluations[0]])
        elif kind == 3:
            q.c.inner_layers[1].hash_witness[3] ^= 0x40
        elif kind == 4:
            q.c.first_layer.fri_witness[0].v[1] ^= 1
        elif kind == 5:
            q.pop_evaluation()
        elif kind == 6:
            q.c.last_layer_poly[0].v[3] ^= 2
        cases.append(q)
        case_seeds.append(seeds[i] + (1 if kind == 7 else 0))
    got = ctx.verify_batch(cases, case_seeds)
    want = []
    for q, s in zip(cases, case_seeds):
        try:
            want.append(int(F.verify_proof(q, s)))
        except F.ReferencePanic:
            want.append(-1)
    assert got == want
    assert got[0] == 1 and got[8] == 1 and got[16] == 1 and got[1] == 0 and got[5] == -1 and got[7] == 0
    for q in cases:
        if q.c.n_evaluations < len(q.evaluations) + 0:
            pass
    for i, q in enumerate(cases):
        if i % 8 == 5:
            q.c.n_evaluations += 1


def test_verify_batch_unseeded_blob_proof(ctx, blob_bytes):
    _, pr = ctx.commit_and_generate_proof(blob_bytes, None, F.PcsConfig(4, 1, 20, 20))
    bad = pr.clone()
    bad.c.inner_layers[7].commitment[31] ^= 0x80
    assert ctx.verify_batch([pr, bad, pr], None) == [1, 0, 1]


# ------------------------------------------------------------------ BASELINE config 3 at scale: properties
def test_c3_batch_properties_full_size_blobs(ctx):
    # 1024 independent 128 KiB blobs (config 3's blob size; the bench runs 4096 per GPU): size-independent
    # properties + a sample against the oracle
    from bench import synth_blobs
    n = 1024
    blobs = synth_blobs(n)
    cfg = F.PcsConfig(4, 0, 20, 20)
    roots = ctx.commit_batch(blobs, 4)
    layer_roots, last = ctx.fri_commit_batch(blobs, None, cfg)
    # (1) first FRI layer commitment == commit() root for every blob (src/proof.rs:126-135)
    assert np.array_equal(layer_roots[:, 0, :], roots)
    # (2) idempotence: a second run gives identical bytes
    layer_roots2, last2 = ctx.fri_commit_batch(blobs, None, cfg)
    assert np.array_equal(layer_roots, layer_roots2) and np.array_equal(last, last2)
    # (3) distinct inputs -> distinct commitments on every layer
    for layer in range(layer_roots.shape[1]):
        assert len({r.tobytes() for r in layer_roots[:, layer, :]}) == n
    # (4) a blob's result does not depend on its batch position
    perm = np.random.default_rng(3).permutation(n)
    roots_p = ctx.commit_batch(np.ascontiguousarray(blobs[perm]), 4)
    assert np.array_equal(roots_p, roots[perm])
    # (5) every value is a canonical M31 element
    assert int(last.max()) < P
    # (6) sample against the oracle
    for b in (0, 511, 1023):
        oroots, olast = O.fri_commit(blobs[b].tobytes(), None, O.make_config(4, 0, 20, 20))
        assert [r.tobytes() for r in layer_roots[b]] == oroots
        assert [tuple(int(x) for x in q) for q in last[b]] == olast


# ------------------------------------------------------------------ rarely exercised paths
def test_commit_device_pointers_unaligned_stride_and_waves(ctx, torch_mod):
    # odd blob stride (pack kernel's unaligned byte path) and several waves through the device entry point
    torch = torch_mod
    rng = np.random.default_rng(31)
    n, blob_len, stride = 37, 4099, 4101
    buf = rng.integers(0, 256, n * stride + 7, dtype=np.uint8)
    d_buf = torch.from_numpy(buf).cuda()
    d_out = torch.zeros((n, 32), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.set_workspace_limit(2 * (1 << 20))
    try:
        ctx.commit_batch_ptr(d_buf.data_ptr() + 3, blob_len, stride, n, 3, d_out.data_ptr(), device=True)
        torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    finally:
        ctx.set_workspace_limit(0)
    got = d_out.cpu().numpy()
    for i in (0, 1, 17, 36):
        data = buf[3 + i * stride: 3 + i * stride + blob_len].tobytes()
        assert got[i].tobytes() == O.commit(data, 3), i


@pytest.mark.parametrize("n_bytes,seed,cfg", [
    (100, 9, (1, 0, 64, 0)),      # more queries than domain points, pow_bits 0 -> nonce 0
    (200, None, (2, 1, 40, 2)),
    (6000, 77, (1, 3, 7, 5)),     # large last layer (log_last 3)
    (131072, 5, (4, 0, 64, 8)),
])
def test_proof_edge_configs_vs_oracle(ctx, n_bytes, seed, cfg):
    data = pattern(n_bytes)
    pr = check_proof_vs_oracle(ctx, data, seed, cfg)
    assert F.verify_proof(pr, seed)
    assert ctx.verify_batch([pr], None if seed is None else [seed]) == [1]
    if cfg[3] == 0:
        assert pr.proof_of_work == 0


def test_verify_batch_bytes(ctx):
    cfg = (3, 0, 16, 8)
    rng = np.random.default_rng(23)
    n, blob_len = 12, 9000
    blobs = rng.integers(0, 256, (n, blob_len), dtype=np.uint8)
    seeds = [5 + i for i in range(n)]
    _, proofs = ctx.prove_batch(blobs, seeds, F.PcsConfig(*cfg))
    proofs[4].c.inner_layers[0].hash_witness[1] ^= 1
    proofs[9].proof_of_work ^= 1
    pieces = [p.serialize() for p in proofs]
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in pieces])
    blob = np.frombuffer(b"".join(pieces), dtype=np.uint8)
    got = ctx.verify_batch_bytes(blob, offs, seeds)
    assert got == [0 if i in (4, 9) else 1 for i in range(n)]
    assert got == ctx.verify_batch(proofs, seeds)
    # truncated / garbage proofs are rejected, not crashed on
    bad = blob.copy()
    bad[int(offs[2]): int(offs[2]) + 4] = 0
    got2 = ctx.verify_batch_bytes(bad, offs, seeds)
    assert got2[2] == 0 and got2[0] == 1


# ------------------------------------------------------------------ error paths through the C ABI
def test_error_paths_fail_loudly(ctx):
    from frieda_b200 import api
    blobs = np.zeros((2, 1000), dtype=np.uint8)
    cases = [
        (lambda: ctx.prove_batch(blobs, None, F.PcsConfig(4, 0, 0, 4)), api.ERR_PANIC),       # n_queries 0 never returns
        (lambda: ctx.prove_batch(blobs, None, F.PcsConfig(4, 0, 5000, 4)), api.ERR_ARG),      # n_queries > 4096
        (lambda: ctx.prove_batch(blobs, None, F.PcsConfig(4, 0, 8, 41)), api.ERR_ARG),        # pow_bits > 40
        (lambda: ctx.fri_commit_batch(blobs, None, F.PcsConfig(8, 2, 8, 4)), api.ERR_ARG),    # last layer > 2^9
        (lambda: ctx.fri_commit_batch(blobs, None, F.PcsConfig(4, 9, 8, 4)), api.ERR_PANIC),  # poly too small
        (lambda: ctx.commit_batch_ptr(blobs.ctypes.data, 1000, 10, 2, 4, blobs.ctypes.data, device=False),
         api.ERR_ARG),                                                                          # stride < len
        (lambda: ctx.commit(bytes(10), 40), api.ERR_ARG),                                      # domain > 2^28
    ]
    for fn, code in cases:
        with pytest.raises(F.FriedaError) as ei:
            fn()
        assert ei.value.code == code, (ei.value.code, code, str(ei.value))
    # the context stays usable after errors
    assert ctx.commit(b"abc", 3) == O.commit(b"abc", 3)


# ------------------------------------------------------------------ encode -> erase -> decode
@pytest.mark.parametrize("n_bytes,blow", [(58, 4), (1000, 2), (4097, 3), (65536, 1), (131072, 4), (262146, 4),
                                          # above one shared-memory chunk: poly_log 16, 17, 20 and config 5's 23
                                          (300000, 3), (1 << 20, 2), (8 << 20, 1), (64 << 20, 2)])
def test_encode_erase_decode_round_trip(ctx, torch_mod, blob_bytes, n_bytes, blow):
    # commit-path evaluations -> keep ONE of the 2^blowup coset blocks (erase the rest) -> recover the bytes
    torch = torch_mod
    data = blob_bytes if n_bytes == 262146 else O.splitmix64_bytes(0x4652494544410000 + n_bytes, n_bytes)
    p = O.poly_log(len(data))
    D = p + blow
    n_felts = (len(data) * 8 + 29) // 30
    coef = np.zeros(4 << p, dtype=np.uint32)
    coef[:n_felts] = O.bytes_to_felts(data)
    d_coef = torch.from_numpy(coef.view(np.int32)).cuda()
    d_eval = torch.zeros(4 << D, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    ctx.pass_lde(d_coef.data_ptr(), p, blow, 1, n_felts, d_eval.data_ptr())
    torch.cuda.ExternalStream(ctx.stream_ptr).synchronize()
    ev = d_eval.cpu().numpy().view(np.uint32).reshape(4, 1 << D)
    blocks = sorted({0, 1, (1 << blow) - 1, (1 << blow) // 2}) if p < 22 else [(1 << blow) - 1]
    for block in blocks:
        piece = ev[:, block << p: (block + 1) << p]
        assert ctx.decode_block(piece, len(data), blow, block) == data, (n_bytes, blow, block)
    # a corrupted block is not an encoding of `len` bytes (or decodes to different data)
    bad = ev[:, : 1 << p].copy()
    bad[0, 0] ^= 1
    try:
        assert ctx.decode_block(bad, len(data), blow, 0) != data
    except F.FriedaError as e:
        assert e.code == -4
    if p < 22:
        # several collected blocks, the first of them corrupted: decoded from the next one
        last = (1 << blow) - 1
        got, used = ctx.decode_blocks([bad, ev[:, last << p: (last + 1) << p]], [0, last], len(data), blow)
        assert got == data and used == 1
        with pytest.raises(F.FriedaError):
            ctx.decode_blocks([bad], [0], len(data), blow)
