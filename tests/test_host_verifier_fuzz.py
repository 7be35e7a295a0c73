"""The host verifier (csrc/verify.cpp: lockstep Merkle walks, 8-/16-way hashing, batched field work) and the GPU batch
verifier's core (csrc/verify_core.cuh, executed on the host) against the oracle's verifier on randomly mutated proofs, at
every host ISA level.  The outcome (accept / reject /
reference panic) must agree mutation by mutation: the reference goes layer by layer and stops at the first
error or panic (src/proof.rs:79-101), which the lockstep schedule has to reproduce."""
import os
import random
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import random, sys
sys.path.insert(0, %(root)r)
import frieda_b200 as F
from frieda_b200 import api
from oracle import oracle as O

def outcome(fn):
    try:
        return "accept" if fn() else "reject"
    except (F.ReferencePanic, O.OraclePanic):
        return "panic"

def layers(c):
    return [c.first_layer] + [c.inner_layers[i] for i in range(c.n_inner_layers)]

def mutate(c, rng):
    # the same mutation is applied to both structs: rng state is identical on both calls
    kind = rng.randrange(10)
    ls = layers(c)
    l = ls[rng.randrange(len(ls))]
    if kind == 0 and c.n_evaluations:
        c.evaluations[rng.randrange(c.n_evaluations)].v[rng.randrange(4)] ^= 1 << rng.randrange(31)
    elif kind == 1 and l.n_fri_witness:
        l.fri_witness[rng.randrange(l.n_fri_witness)].v[rng.randrange(4)] ^= 1 << rng.randrange(31)
    elif kind == 2 and l.n_hash_witness:
        l.hash_witness[rng.randrange(32 * l.n_hash_witness)] ^= 1 << rng.randrange(8)
    elif kind == 3:
        l.commitment[rng.randrange(32)] ^= 1 << rng.randrange(8)
    elif kind == 4 and l.n_fri_witness:
        l.n_fri_witness -= 1
    elif kind == 5 and l.n_hash_witness:
        l.n_hash_witness -= rng.randrange(1, min(3, l.n_hash_witness) + 1)
    elif kind == 6:
        c.last_layer_poly[0].v[rng.randrange(4)] ^= 1 << rng.randrange(31)
    elif kind == 7:
        c.proof_of_work += rng.randrange(1, 5)
    elif kind == 8 and c.n_evaluations:
        c.n_evaluations -= rng.randrange(1, min(4, c.n_evaluations) + 1)
    elif kind == 9 and c.n_inner_layers:
        c.n_inner_layers -= 1
    return kind

n_mut, seed0 = int(sys.argv[1]), int(sys.argv[2])
counts = {"accept": 0, "reject": 0, "panic": 0}
for case, (n, cfg, seed) in enumerate([(3000, (2, 1, 9, 4), 5), (20000, (3, 0, 14, 6), None), (262146, (4, 0, 20, 8), 77)]):
    data = bytes((i * 7 + case) %% 256 for i in range(n))
    _, opr = O.prove(data, seed, O.make_config(*cfg))
    raw = opr.serialize()
    assert outcome(lambda: F.verify_proof(F.Proof.deserialize(raw), seed)) == "accept"
    for k in range(n_mut):
        fp, op = F.Proof.deserialize(raw), opr.clone()
        kf = mutate(fp.c, random.Random(seed0 * 100003 + case * 1009 + k))
        ko = mutate(op.c, random.Random(seed0 * 100003 + case * 1009 + k))
        assert kf == ko
        a, b = outcome(lambda: F.verify_proof(fp, seed)), outcome(lambda: O.verify(op, seed))
        assert a == b, (case, k, kf, a, b)
        # the GPU batch verifier's core (same __host__ __device__ code as the kernels), run on the host
        c = {1: "accept", 0: "reject", api.ERR_PANIC: "panic"}.get(api.verify_core_host(fp, seed), "error")
        assert c == a, (case, k, kf, a, c)
        counts[a] += 1
        # restore shrunk counters so the owners free what they allocated
        rf, ro = F.Proof.deserialize(raw).c, opr.c
        for dst, src in ((fp.c, rf), (op.c, ro)):
            dst.n_evaluations, dst.n_inner_layers = src.n_evaluations, src.n_inner_layers
            for ld, ls in zip(layers(dst), layers(src)):
                ld.n_fri_witness, ld.n_hash_witness = ls.n_fri_witness, ls.n_hash_witness
print("FUZZ_OK", counts)
"""


@pytest.mark.parametrize("isa", ["scalar", "avx2", "native"])
def test_host_verifier_agrees_with_oracle_on_mutations(isa, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % {"root": ROOT})
    env = dict(os.environ)
    env.pop("FRIEDA_HOST_ISA", None)
    if isa != "native":
        env["FRIEDA_HOST_ISA"] = isa
    r = subprocess.run([sys.executable, str(script), "120", str(random.Random(isa).randrange(1 << 20))], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "FUZZ_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    # the mutations must exercise both failing outcomes, not only "reject"
    counts = eval(r.stdout.split("FUZZ_OK", 1)[1])
    assert counts["reject"] > 100 and counts["panic"] > 0 and counts["accept"] == 0, counts


def test_deserializer_and_verifier_survive_corrupted_bytes(blob_bytes):
    """Memory safety of the proof codec and the host verifier: corrupted serialisations either fail to parse
    (FriedaError) or verify to accept / reject / reference panic -- never anything else."""
    import frieda_b200 as F
    from oracle import oracle as O
    _, opr = O.prove(blob_bytes[:20000], 3, O.make_config(3, 0, 12, 4))
    raw = opr.serialize()
    assert F.verify_proof(F.Proof.deserialize(raw), 3)
    rng = random.Random(1234)
    parsed = rejected = 0
    for k in range(600):
        b = bytearray(raw)
        kind = rng.randrange(5)
        if kind == 0:      # flip bits anywhere (headers and counts included)
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        elif kind == 1:    # truncate
            del b[rng.randrange(len(b)):]
        elif kind == 2:    # overwrite a 4-byte word with an extreme value
            i = rng.randrange(0, len(b) - 4)
            b[i:i + 4] = rng.choice([b"\xff\xff\xff\xff", b"\x00\x00\x00\x00", b"\x00\x00\x00\x80", b"\xff\xff\xff\x7f"])
        elif kind == 3:    # extend with junk
            b += bytes(rng.randrange(256) for _ in range(rng.randrange(1, 64)))
        else:              # splice a window from elsewhere
            i, j, n = rng.randrange(len(b) - 64), rng.randrange(len(b) - 64), rng.randrange(1, 64)
            b[i:i + n] = b[j:j + n]
        try:
            p = F.Proof.deserialize(bytes(b))
        except F.FriedaError:
            rejected += 1
            continue
        parsed += 1
        try:
            assert F.verify_proof(p, 3) in (True, False)
        except F.ReferencePanic:
            pass
    assert parsed > 50 and rejected > 50, (parsed, rejected)


# ---------------------------------------------------------------------------------------------------------------
# The batch verifier's own parser (verify_core.cuh: VReader) sees RAW untrusted words: frieda_verify_batch_bytes does not
# go through proof.cpp's deserializer.  Counts are untrusted u32 words, so `4 * n` style arithmetic must not wrap.
def count_word_positions(raw: bytes):
    """Word indices of every count field of a serialised proof (n_evals, n_last, n_layers, and per layer n_fri,
    n_hash, n_colw), from the FRDA layout of csrc/proof.cpp."""
    import struct
    w = struct.unpack("<%dI" % (len(raw) // 4), raw)
    pos = [9]
    p = 10 + 4 * w[9]
    pos.append(p)                    # n_last
    p += 1 + 4 * w[p]
    pos.append(p)                    # n_layers
    n_layers = w[p]
    p += 1
    for _ in range(n_layers):
        p += 8
        pos.append(p)                # n_fri_witness
        p += 1 + 4 * w[p]
        pos.append(p)                # n_hash_witness
        p += 1 + 8 * w[p]
        pos.append(p)                # n_column_witness
        p += 1 + w[p]
    assert p == len(w)
    return pos


WRAP_VALUES = [0x20000000, 0x40000000, 0x80000000, 0xC0000000, 0xFFFFFFFF, 0x20000001, 0x3FFFFFFF, 0x10000000,
               0x7FFFFFFF, 0xE0000000]


def count_mutations(raw: bytes, rng, n_random=200):
    """Every count word x every wrapping value, then random mixes (several counts at once, pow_bits 0 so that the
    proof-of-work check does not stop the parser early, as an attacker would set it)."""
    import struct
    pos = count_word_positions(raw)
    out = []
    for p in pos:
        for v in WRAP_VALUES:
            b = bytearray(raw)
            struct.pack_into("<I", b, 4 * p, v)
            out.append(bytes(b))
            b2 = bytearray(b)
            struct.pack_into("<I", b2, 4 * 6, 0)  # pow_bits = 0
            out.append(bytes(b2))
    # the attack itself: a list's DATA removed, its count set so that `words_per_item * count` wraps to 0 (or to a few
    # items) in 32 bits -- the reader then stays in bounds while the walks trust the unwrapped count
    w = list(struct.unpack("<%dI" % (len(raw) // 4), raw))
    per_item = {0: 4, 1: 4}  # n_evals, n_last
    for k, p in enumerate(pos[3:]):
        per_item[3 + k] = (4, 8, 1)[k % 3]
    for k, p in enumerate(pos):
        if k == 2:
            continue  # n_layers has no list of its own
        item = per_item[k]
        for keep in (0, 1, 2):
            if w[p] < keep:
                continue
            cut = w[:p + 1 + item * keep] + w[p + 1 + item * w[p]:]
            for mult in (1, 2, 3):
                cut[p] = ((1 << 32) // item) * mult % (1 << 32) + keep if item > 1 else 0xFFFFFFF0 + keep
                for pow0 in (False, True):
                    c2 = list(cut)
                    if pow0:
                        c2[6] = 0
                    out.append(struct.pack("<%dI" % len(c2), *c2))
    for _ in range(n_random):
        b = bytearray(raw)
        for p in rng.sample(pos, rng.randrange(1, 4)):
            v = rng.choice(WRAP_VALUES + [rng.randrange(1 << 32), rng.randrange(64)])
            struct.pack_into("<I", b, 4 * p, v)
        if rng.random() < 0.7:
            struct.pack_into("<I", b, 4 * 6, 0)
        out.append(bytes(b))
    return out


def test_batch_verifier_core_on_raw_words_survives_wrapping_counts(blob_bytes):
    """n_hash_witness = 0x20000000 made take(8 * n) wrap to take(0).  Host-executed core, raw
    words in, every buffer exactly sized (a heap overrun would be caught by the allocator / crash the worker)."""
    import frieda_b200 as F
    from frieda_b200 import api
    from oracle import oracle as O
    _, opr = O.prove(blob_bytes[:20000], 3, O.make_config(3, 0, 12, 4))
    raw = opr.serialize()
    assert api.verify_core_host_bytes(raw, 3) == 1
    assert api.verify_core_host_bytes(raw, 4) == 0
    muts = count_mutations(raw, random.Random(99))
    outcomes = {1: 0, 0: 0, -1: 0}
    for m in muts:
        r = api.verify_core_host_bytes(m, 3)
        assert r in (1, 0, -1), r
        outcomes[r] += 1
        # whatever the host deserializer accepts must get the same verdict from the host verifier
        try:
            p = F.Proof.deserialize(m)
        except F.FriedaError:
            assert r != 1, "the core accepted bytes the deserializer rejects"
            continue
        try:
            want = int(F.verify_proof(p, 3))
        except F.ReferencePanic:
            want = -1
        assert r == want
    assert outcomes[1] == 0 and outcomes[0] > 100, outcomes
    # capacity limits are an argument error, not "forged"
    import struct
    big = bytearray(raw)
    struct.pack_into("<I", big, 4 * 4, 5000)
    assert api.verify_core_host_bytes(bytes(big), 3) == api.ERR_ARG


GUARD_WORKER = r"""
import ctypes as C, mmap, random, sys
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(root)r + "/tests")
from frieda_b200 import api
from oracle import oracle as O
from test_host_verifier_fuzz import count_mutations
L = api.load_library()
libc = C.CDLL(None, use_errno=True)
libc.mprotect.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
PAGE = mmap.PAGESIZE
data = open(%(root)r + "/tests/golden/blob", "rb").read()[:20000]
_, opr = O.prove(data, 3, O.make_config(3, 0, 12, 4))
raw = opr.serialize()
span = (len(raw) + 64 + PAGE - 1) // PAGE * PAGE
mm = mmap.mmap(-1, span + PAGE)
base = C.addressof(C.c_char.from_buffer(mm))
assert libc.mprotect(base + span, PAGE, 0) == 0        # PROT_NONE guard page right behind the proof
seed = C.c_uint64(3)
n = 0
for m in [raw] + count_mutations(raw, random.Random(5), n_random=400):
    at = base + span - len(m)                          # the proof's last word touches the guard page
    C.memmove(at, m, len(m))
    r = L.frieda_verify_core_host_bytes(at, len(m), C.byref(seed))
    assert r in (1, 0, -1), r
    n += 1
print("GUARD_OK", n)
"""


def test_batch_verifier_core_never_reads_past_the_proof(tmp_path):
    """Each crafted proof ends exactly at a PROT_NONE page: one word read past the proof (what the wrapped
    `take(8 * n_hash)` allowed: hw[8 * used] with `used` bounded only by the unwrapped count) kills the worker."""
    script = tmp_path / "guard.py"
    script.write_text(GUARD_WORKER % {"root": ROOT})
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "GUARD_OK" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-2000:])


@pytest.mark.gpu
def test_verify_batch_bytes_survives_wrapping_counts(blob_bytes):
    """Same mutations through the kernels (frieda_verify_batch_bytes): verdict per proof must equal the host-executed
    core's, the valid proofs interleaved with the crafted ones must still verify, and the context must stay usable
    (an out-of-bounds read would poison it with an illegal-address error)."""
    import numpy as np
    import frieda_b200 as F
    from frieda_b200 import api
    from oracle import oracle as O
    _, opr = O.prove(blob_bytes[:20000], 3, O.make_config(3, 0, 12, 4))
    raw = opr.serialize()
    muts = count_mutations(raw, random.Random(7), n_random=300)
    pieces = []
    for i, m in enumerate(muts):
        pieces.append(m)
        if i % 16 == 0:
            pieces.append(raw)
    offs = np.zeros(len(pieces) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(b) for b in pieces])
    blob = np.frombuffer(b"".join(pieces), dtype=np.uint8)
    ctx = F.Context(0)
    try:
        got = ctx.verify_batch_bytes(blob, offs, [3] * len(pieces))
        want = [api.verify_core_host_bytes(b, 3) for b in pieces]
        assert got == want
        assert sum(1 for g in got if g == 1) == sum(1 for b in pieces if b == raw)
        # still alive
        assert ctx.commit(b"abc", 3) == O.commit(b"abc", 3)
        # offsets beyond the buffer, wrong seed count, capacity: argument errors
        bad = offs.copy()
        bad[-1] += 4
        with pytest.raises(F.FriedaError) as ei:
            ctx.verify_batch_bytes(blob, bad, [3] * len(pieces))
        assert ei.value.code == api.ERR_ARG
        with pytest.raises(F.FriedaError):
            ctx.verify_batch_bytes(blob, offs, [3])
        import struct
        big = bytearray(raw)
        struct.pack_into("<I", big, 4 * 4, 5000)
        with pytest.raises(F.FriedaError) as ei:
            ctx.verify_batch_bytes(np.frombuffer(bytes(big), dtype=np.uint8), np.array([0, len(big)], dtype=np.uint64), [3])
        assert ei.value.code == api.ERR_ARG
    finally:
        ctx.close()
