"""The C++ host mirror of the reference interface (include/zkb200.hpp) over libzkb200.so: tests/cpp/reference_tests.cpp
restates the reference's own BN254 tests (fr.rs:248-416, mod.rs:635-690).  Without a GPU: it compiles, links and
fails loudly (no CPU fallback).  With a GPU: every test passes and the proof made from fixed secrets equals the
oracle's literal restatement of groth16::prove bit for bit."""

import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "zksnark-rs_b200")
EXE = os.path.join(HERE, "cpp", "_reference_tests")


@pytest.fixture(scope="module")
def exe():
    if not os.path.exists(os.path.join(PKG, "libzkb200.so")):
        subprocess.check_call(["python", os.path.join(PKG, "build.py")])
    src = os.path.join(HERE, "cpp", "reference_tests.cpp")
    deps = [src, os.path.join(ROOT, "include", "zkb200.hpp"), os.path.join(ROOT, "include", "zkb200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                               "-L", PKG, "-lzkb200", f"-Wl,-rpath,{PKG}"])
    return EXE


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_cpp_mirror_builds_and_fails_loudly_without_gpu(exe):
    if _has_gpu():
        pytest.skip("GPU present: covered by the gpu test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "zkb200::Error" in r.stdout and "no CPU fallback" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cpp_reference_tests_on_gpu(exe):
    from oracle import groth16 as og, synthetic
    from oracle.fields import FR
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for name in ("single_mult_honest_bn", "qap_from_roots", "bn_encrypt_deg_15_test", "batch_equals_single", "weights_test", "parity_dump"):
        assert f"ok {name}" in r.stdout
    got = {}
    for line in r.stdout.splitlines():
        if line.startswith("proof."):
            k, *limbs = line.split()
            vals = [int(x, 16) for x in limbs]
            got[k] = [sum(vals[4 * i + j] << (64 * j) for j in range(4)) for i in range(len(vals) // 4)]
    # the same proof from the literal restatement of the reference
    P, n = FR.p, 8
    w = synthetic.omega(3)
    rep = synthetic.horner_rep(FR, n, [pow(w, k, P) for k in range(n)])
    cs = [(1000000007 * (k + 3) + k) % P for k in range(n)]
    wit = synthetic.horner_witness(FR, n, 123456789012345678901234567890 % P, cs)
    B = og.BN254Backend()
    dense = og.qap_from_root_rep(FR, rep)
    sig = og.setup(B, dense, (3, 5, 7, 11, 13))
    want = og.prove(B, dense, sig, wit, 17, 19)
    assert got["proof.a"] == list(want.a)
    assert got["proof.b"] == [want.b[0][0], want.b[0][1], want.b[1][0], want.b[1][1]]
    assert got["proof.c"] == list(want.c)


def test_random_elem_comes_from_the_os_csprng(tmp_path):
    """The toxic waste of setup() and the blinding r, s of prove() are drawn with Fr::random_elem (fr.rs:90-99; the
    reference uses rand::thread_rng, an OS-seeded CSPRNG).  The host mirror must read the kernel CSPRNG directly: no
    seeded userspace generator anywhere in the header, values uniform below r, never zero, no repeats."""
    hdr = open(os.path.join(ROOT, "include", "zkb200.hpp")).read()
    for banned in ("mt19937", "random_device", "minstd", "srand(", "rand()", "<random>"):
        assert banned not in hdr, f"zkb200.hpp uses {banned}"
    assert "getrandom" in hdr and "/dev/urandom" in hdr
    src = tmp_path / "rnd.cpp"
    src.write_text('#include <cstdio>\n#include "zkb200.hpp"\nint main() { for (int i = 0; i < 2000; i++) { zkb200::Fr r = zkb200::Fr::random_elem();'
                   ' std::printf("%016lx%016lx%016lx%016lx\\n", (unsigned long)r.l[3], (unsigned long)r.l[2], (unsigned long)r.l[1], (unsigned long)r.l[0]); } }\n')
    exe = tmp_path / "rnd"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", PKG, "-lzkb200", f"-Wl,-rpath,{PKG}"])
    vals = [int(x, 16) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    r = 21888242871839275222246405745257275088548364400416034343698204186575808495617
    assert len(vals) == 2000 and len(set(vals)) == 2000 and all(0 < v < r for v in vals)
    assert max(vals) > r * 3 // 4 and min(vals) < r // 4  # spread over the whole range
    ones = sum(bin(v & ((1 << 250) - 1)).count("1") for v in vals) / (2000 * 250)
    assert 0.48 < ones < 0.52
