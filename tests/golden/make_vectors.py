"""Regenerates tests/golden/vectors.json.

Provenance of each group:
  reference_certified : literal from the reference's own test (src/commit.rs:31-37).
  survey_probe        : SURVEY.md Appendix B -- produced by an independent throwaway restatement
                        during the survey (it reproduced the certified root); copied verbatim.
  oracle_generated    : produced here by oracle/frieda_oracle.c (which reproduces both groups
                        above); pins the CUDA path against regressions at more sizes.
The reference itself (Rust + un-vendored stwo) cannot be built in this container.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pattern(n):
    return bytes(i % 256 for i in range(n))


def main():
    blob = open(os.path.join(HERE, "blob"), "rb").read()
    v = {
        "reference_certified": {
            "commit_blob_blowup4": bytes([209, 162, 213, 6, 157, 197, 135, 229, 93, 194, 156, 198, 37, 90, 249, 55,
                                          255, 127, 237, 14, 228, 27, 223, 90, 249, 135, 23, 249, 215, 79, 96,
                                          232]).hex(),
        },
        "survey_probe": {
            "commit_pattern_1024": "636256479b7a848e6664d0d423ec1f84c3d41788c6f783e99ab4e5d058d8264c",
            "commit_pattern_65536": "9dd2ca907612cf593460ee5db5a7f341311f240e0893d2863a71b15e398c452d",
            "commit_pattern_131072": "385fafc45a1dedc1fdcdeafc50e857744151fbe88f2fe00e38301a9d173fd763",
            "commit_zeros_131072": "9c45a4a5c3b94328dd12e41e47f0a3831909341a816ff491efa4379c02e5abfb",
            "commit_e2e_string": "6f457ff594f495af0d391b43086e82d1364e5f7122d5ab731a59c149b0263479",
            "compress_zero": "2ab0266e0d80a7a6897bbb5151db011f2d947b072a12e623e1487da35ec7006e",
            "blob_felts": {"count": 69906, "first4": [942766128, 92852424, 907252563, 202201304], "last": 52444},
            "blob_col0_first4": [656067250, 2045105345, 49624580, 1456778668],
            "blob_leaf0": "60713c4b7369fd93e55886bf16de81a5613500741d0708cccca84fb14a8672a8",
            "blob_layer18_node0": "ce93241410fdb2d3932330d4dfff9fee790ec70446f0dece25fe64e13a420ea7",
            "twiddles_k18": {"first4": [1633461177, 574296567, 1524138903, 924762334], "penultimate": 32768,
                             "last": 1},
            "blob_proof_cfg_4_1_20_20": {
                "inner_root_prefix_suffix": [["f4718527", "31eb"], ["86b301ec", "51f3"], ["f810fb8c", "f2dc"]],
                "last_layer_poly": [[1218608362, 659502217, 149073738, 2047414837],
                                    [718176909, 75873613, 204177336, 893093088]],
                "nonce": 128474,
                "first_queries": [16022, 70843, 74563, 133715, 139281],
                "first_layer": {"positions": 40, "fri_witness": 20, "hash_witness": 262},
                "all_layers_hash_witness": 1858, "inner_fri_witness": 250,
            },
        },
        "oracle_generated": {},
    }
    og = v["oracle_generated"]
    og["commit"] = []
    for name, data, blow in [
        ("empty", b"", 4), ("one_byte", b"\x07", 4), ("15_bytes", pattern(15), 1), ("16_bytes", pattern(16), 2),
        ("pattern_100", pattern(100), 3), ("pattern_4096", pattern(4096), 4), ("pattern_16384", pattern(16384), 4),
        ("pattern_5000_b1", pattern(5000), 1), ("pattern_5000_b6", pattern(5000), 6),
        ("splitmix_c2", O.splitmix64_bytes(0x4652494544410000, 131072), 4),
        ("splitmix_c2_b1", O.splitmix64_bytes(0x4652494544410001, 131072), 4),
        ("splitmix_1MiB_b2", O.splitmix64_bytes(0x4652494544414236, 1 << 20), 2),
    ]:
        og["commit"].append({"name": name, "len": len(data), "log_blowup": blow, "root": O.commit(data, blow).hex()})
    # BASELINE config 5 (64 MiB, blowup 2^2) and an 8 MiB sibling; ~25 s of oracle time
    og["commit_large"] = []
    for name, n in [("c5_splitmix_64MiB_b2", 64 << 20), ("splitmix_8MiB_b2", 8 << 20)]:
        og["commit_large"].append({"name": name, "len": n, "log_blowup": 2, "state0": "0x4652494544414236",
                                   "root": O.commit(O.splitmix64_bytes(0x4652494544414236, n), 2).hex()})
    og["prove"] = []
    import hashlib
    for name, data, seed, cfg in [
        ("blob_4_1_20_20", blob, None, (4, 1, 20, 20)),
        ("blob_4_0_20_20_seedlen", blob, len(blob), (4, 0, 20, 20)),
        ("pattern_1024_seedlen", pattern(1024), 1024, (4, 0, 20, 20)),
        ("pattern_65536_seedlen", pattern(65536), 65536, (4, 0, 20, 20)),
        ("e2e_string", b"This is the original data that needs to be made available.", None, (4, 0, 20, 20)),
        ("splitmix_c4_b0", O.splitmix64_bytes(0x4652494544410000, 131072), 0, (4, 0, 64, 20)),
        ("splitmix_c4_b7", O.splitmix64_bytes(0x4652494544410007, 131072), 7, (4, 0, 64, 20)),
        ("pattern_3000_b2_l2", pattern(3000), 5, (2, 2, 33, 8)),
        # above one shared-memory LDE block (poly_log 17 / 20: strided LDE passes, extra Merkle middle passes)
        ("splitmix_big_1MiB_b2", O.splitmix64_bytes(0x4652494544414236, 1 << 20), 11, (2, 0, 20, 12)),
        ("splitmix_big_8MiB_b2", O.splitmix64_bytes(0x4652494544414236, 8 << 20), None, (2, 0, 20, 12)),
        ("splitmix_big_1MiB_b4_l3", O.splitmix64_bytes(0x4652494544414237, 1 << 20), 3, (4, 3, 64, 10)),
    ]:
        c = O.make_config(*cfg)
        t = O.trace(data, seed, c, with_trees=False)
        og["prove"].append({
            "name": name, "len": len(data), "seed": seed, "cfg": list(cfg), "root": t.root.hex(),
            "layer_roots": None, "alphas": [list(a) for a in t.alphas], "last_layer_poly": [list(q) for q in
                                                                                            t.last_layer_poly],
            "nonce": int(t.nonce), "queries": [int(q) for q in t.queries],
            "proof_sha256": hashlib.sha256(t.proof_bytes).hexdigest(), "proof_len": len(t.proof_bytes),
        })
        roots, _ = O.fri_commit(data, seed, c)
        og["prove"][-1]["layer_roots"] = [r.hex() for r in roots]
    with open(os.path.join(HERE, "vectors.json"), "w") as f:
        json.dump(v, f, indent=1)
    print("wrote vectors.json")


if __name__ == "__main__":
    main()
