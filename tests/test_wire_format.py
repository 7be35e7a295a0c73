"""Wire formats of `Proof`: the library's flat "FRDA" encoding and the bincode-1.x layout of the
reference's serde-derived struct (src/proof.rs:19-26).  No GPU: proof objects are built from bytes.

The bincode check is a SPEC test: the expected bytes are produced here, independently of csrc/proof.cpp, from the
declaration order of the Rust structs (serde derives serialise fields in declaration order; bincode 1.x default
options: little-endian, fixed-width integers, u64 length prefix for every Vec, nothing for fixed arrays / newtypes):

  frieda::proof::Proof        { proof, proof_of_work: u64, pcs_config, log_size_bound: u32, evaluations: Vec<QM31> }
  stwo FriProof<H>            { first_layer, inner_layers: Vec<FriLayerProof<H>>, last_layer_poly: LinePoly }
  stwo FriLayerProof<H>       { fri_witness: Vec<QM31>, decommitment, commitment: Blake2sHash }
  stwo MerkleDecommitment<H>  { hash_witness: Vec<Blake2sHash>, column_witness: Vec<M31> }
  stwo LinePoly               { coeffs: Vec<QM31>, log_size: u32 }
  stwo PcsConfig              { pow_bits: u32, fri_config: FriConfig { log_blowup_factor: u32,
                                log_last_layer_degree_bound: u32, n_queries: usize } }
  QM31(CM31(M31, M31), CM31(M31, M31)), M31(u32), Blake2sHash([u8; 32])

It cannot be validated against the Rust crate here (no toolchain; rust/src/lib.rs holds the test a maintainer runs).
"""
import struct

import frieda_b200 as F


def frda(log_size_bound, cfg, pow_nonce, evals, last, layers):
    """The flat encoding of csrc/proof.cpp."""
    out = b"FRDA" + struct.pack("<IIIQIQ", log_size_bound, cfg[0], cfg[1], cfg[2], cfg[3], pow_nonce)
    for lst in (evals, last):
        out += struct.pack("<I", len(lst)) + b"".join(struct.pack("<4I", *q) for q in lst)
    out += struct.pack("<I", len(layers))
    for commitment, fri, hashes, colw in layers:
        out += commitment
        out += struct.pack("<I", len(fri)) + b"".join(struct.pack("<4I", *q) for q in fri)
        out += struct.pack("<I", len(hashes)) + b"".join(hashes)
        out += struct.pack("<I", len(colw)) + b"".join(struct.pack("<I", x) for x in colw)
    return out


def bincode_spec(log_size_bound, cfg, pow_nonce, evals, last, layers):
    u32, u64 = (lambda x: struct.pack("<I", x)), (lambda x: struct.pack("<Q", x))
    vec_qm31 = lambda v: u64(len(v)) + b"".join(struct.pack("<4I", *q) for q in v)  # noqa: E731

    def layer(l):
        commitment, fri, hashes, colw = l
        return (vec_qm31(fri)                                          # FriLayerProof.fri_witness
                + u64(len(hashes)) + b"".join(hashes)                  # .decommitment.hash_witness
                + u64(len(colw)) + b"".join(u32(x) for x in colw)      # .decommitment.column_witness
                + commitment)                                          # .commitment
    out = layer(layers[0])                                             # Proof.proof.first_layer
    out += u64(len(layers) - 1) + b"".join(layer(l) for l in layers[1:])   # .inner_layers
    out += vec_qm31(last) + u32(max(len(last), 1).bit_length() - 1)    # .last_layer_poly { coeffs, log_size }
    out += u64(pow_nonce)                                              # Proof.proof_of_work
    out += u32(cfg[3]) + u32(cfg[0]) + u32(cfg[1]) + u64(cfg[2])       # Proof.pcs_config { pow_bits, fri_config }
    out += u32(log_size_bound)                                         # Proof.log_size_bound
    out += vec_qm31(evals)                                             # Proof.evaluations
    return out


def h(b):
    return bytes([b]) * 32


TINY = dict(log_size_bound=3, cfg=(1, 0, 2, 5), pow_nonce=0x0102030405060708,
            evals=[(1, 2, 3, 4)], last=[(9, 8, 7, 6)],
            layers=[(h(0xAA), [(10, 11, 12, 13)], [h(0x01), h(0x02)], []),
                    (h(0xBB), [], [h(0x03)], [])])

# the same proof written out by hand, field by field (hex, little-endian)
TINY_BINCODE_HEX = (
    # first_layer: fri_witness len 1 + one QM31; hash_witness len 2 + 2 x 32 B; column_witness len 0; commitment
    "0100000000000000" "0a000000" "0b000000" "0c000000" "0d000000"
    "0200000000000000" + "01" * 32 + "02" * 32 +
    "0000000000000000" + "aa" * 32 +
    # inner_layers: len 1; its fri_witness len 0; hash_witness len 1; column_witness len 0; commitment
    "0100000000000000"
    "0000000000000000"
    "0100000000000000" + "03" * 32 +
    "0000000000000000" + "bb" * 32 +
    # last_layer_poly: coeffs len 1 + one QM31, log_size 0
    "0100000000000000" "09000000" "08000000" "07000000" "06000000" "00000000"
    # proof_of_work
    "0807060504030201"
    # pcs_config: pow_bits 5; fri_config: log_blowup 1, log_last 0, n_queries 2 (usize -> u64)
    "05000000" "01000000" "00000000" "0200000000000000"
    # log_size_bound
    "03000000"
    # evaluations: len 1 + one QM31
    "0100000000000000" "01000000" "02000000" "03000000" "04000000"
)


def test_bincode_layout_of_a_hand_written_proof():
    p = F.Proof.deserialize(frda(**TINY))
    got = p.serialize_bincode()
    assert got.hex() == TINY_BINCODE_HEX
    assert got == bincode_spec(**TINY)
    assert p.serialize() == frda(**TINY)        # the flat encoding round-trips byte for byte


def test_bincode_layout_of_a_real_proof(blob_bytes):
    from oracle import oracle as O
    cfg = (3, 1, 12, 4)
    _, opr = O.prove(blob_bytes[:20000], 3, O.make_config(*cfg))
    p = F.Proof.deserialize(opr.serialize())
    c = p.c
    layers = []
    for l in [c.first_layer] + [c.inner_layers[i] for i in range(c.n_inner_layers)]:
        layers.append((bytes(l.commitment), [l.fri_witness[i].tuple() for i in range(l.n_fri_witness)],
                       [bytes(l.hash_witness[32 * i: 32 * i + 32]) for i in range(l.n_hash_witness)],
                       [int(l.column_witness[i]) for i in range(l.n_column_witness)]))
    want = bincode_spec(p.log_size_bound, cfg, p.proof_of_work, p.evaluations, p.last_layer_poly, layers)
    assert p.serialize_bincode() == want
    assert len(p.last_layer_poly) == 2 and want.count(struct.pack("<Q", 2) + struct.pack("<4I", *p.last_layer_poly[0])) == 1
