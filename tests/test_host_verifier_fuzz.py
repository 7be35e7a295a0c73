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
