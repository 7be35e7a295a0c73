"""Builds and runs the C++ host-API tests (tests/cpp/test_frieda_api.cpp over include/frieda.hpp): the
compiled-language mirror of the reference's interface, tests written after the reference's own."""
import os
import subprocess

import pytest

import frieda_b200 as F
from frieda_b200 import build as fb
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cpp_binary(tmp_path_factory):
    F.load_library()
    out = tmp_path_factory.mktemp("cpp") / "test_frieda_api"
    libdir = os.path.dirname(fb.LIB)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "test_frieda_api.cpp"), "-o", str(out), "-L", libdir, "-lfrieda_b200",
           f"-Wl,-rpath,{libdir}"]
    subprocess.check_call(cmd)
    return str(out)


def test_cpp_api_host_only(cpp_binary, blob_bytes, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device behaviour is checked on the CPU box")
    _, pr = O.prove(blob_bytes, None, O.make_config(4, 1, 20, 20))
    proof_file = tmp_path / "proof.bin"
    proof_file.write_bytes(pr.serialize())
    r = subprocess.run([cpp_binary, os.path.join(ROOT, "tests", "golden", "blob"), "cpu", str(proof_file)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_api_reference_test_suite_on_gpu(cpp_binary):
    r = subprocess.run([cpp_binary, os.path.join(ROOT, "tests", "golden", "blob"), "gpu"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr


# ---- the reference's criterion benches re-expressed over the C++ host API (benches/*.cpp)
@pytest.fixture(scope="module")
def bench_binaries(tmp_path_factory):
    F.load_library()
    libdir = os.path.dirname(fb.LIB)
    outs = {}
    for name in ("commit", "proof"):
        out = tmp_path_factory.mktemp("benches") / name
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                               os.path.join(ROOT, "benches", name + ".cpp"), "-o", str(out), "-L", libdir,
                               "-lfrieda_b200", f"-Wl,-rpath,{libdir}"])
        outs[name] = str(out)
    return outs


def test_benches_build_and_fail_loudly_without_gpu(bench_binaries):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = subprocess.run([bench_binaries["commit"], os.path.join(ROOT, "tests", "golden", "blob")], capture_output=True,
                       text=True, timeout=60)
    assert r.returncode == 1 and "no usable CUDA device" in r.stderr  # no CPU fallback


@pytest.mark.gpu
def test_benches_run_every_reference_group(bench_binaries):
    import json
    env = dict(os.environ, FRIEDA_BENCH_SECONDS="0.02")
    rows = []
    for name in ("commit", "proof"):
        r = subprocess.run([bench_binaries[name], os.path.join(ROOT, "tests", "golden", "blob")], capture_output=True,
                           text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stdout + r.stderr
        rows += [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    groups = {(x["group"], x["param"]) for x in rows}
    sizes = (1024, 4096, 16384, 65536, 262146)
    for g in ("commit", "generate_proof", "commit_and_generate_proof", "verify_proof"):  # benches/*.rs groups
        for s in sizes:
            assert (g, s) in groups
