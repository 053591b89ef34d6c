"""The C++ mirror of the Go API (secp256k1-voi_b200/host/secp256k1_voi.hpp):
compiles and links against the C-ABI library on CPU; runs on the GPU box."""
import os
import subprocess

import pytest

from conftest import ROOT, load_golden

SRC = os.path.join(ROOT, "tests", "cpp", "test_host_mirror.cpp")
OUT = os.path.join(ROOT, "tests", "cpp", "_build", "test_host_mirror")


def build(s256):
    lib = s256.load_library()  # builds the .so if needed
    libdir = os.path.dirname(s256.library_path())
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", OUT, SRC, "-L", libdir, "-lsecp256k1_b200",
                           f"-Wl,-rpath,{libdir}"])
    return OUT


def test_mirror_compiles_and_links(s256):
    assert os.path.exists(build(s256))


@pytest.mark.gpu
def test_mirror_runs_reference_style_checks(s256):
    exe = build(s256)
    k = load_golden("kats.json")
    row0 = load_golden("bip340.json")["rows"][0]
    rfc = load_golden("rfc6979.json")["rows"][0]
    suite = [s for s in load_golden("h2c.json")["suites"] if s["random_oracle"]][0]
    vec = [v for v in suite["vectors"] if v["msg"] == "abc"][0]
    r = subprocess.run([exe, k["g_uncompressed"], k["g_compressed"], k["libsecp_a"], k["libsecp_xn"], k["libsecp_b"],
                        row0["pk"], row0["sig"], rfc["priv"], rfc["digest"], rfc["r"] + rfc["s"], row0["sk"], row0["aux"],
                        suite["dst"], vec["msg"], vec["Px"] + vec["Py"]], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host mirror ok" in r.stdout


def test_prehash_schnorr_message_matches_hashlib(s256):
    """secec/bitcoin/schnorr.go:56 PreHashSchnorrMessage -- host hashing in the mirror, no device needed."""
    import hashlib
    s256.load_library()
    libdir = os.path.dirname(s256.library_path())
    exe = os.path.join(ROOT, "tests", "cpp", "_build", "prehash_host")
    os.makedirs(os.path.dirname(exe), exist_ok=True)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "prehash_host.cpp"),
                           "-L", libdir, "-lsecp256k1_b200", f"-Wl,-rpath,{libdir}"])

    def run(name: bytes, msg: bytes):
        r = subprocess.run([exe, name.hex() or "", msg.hex() or ""], capture_output=True, text=True, timeout=60)
        assert r.returncode == 0, r.stderr
        return r.stdout.strip()

    def want(name: bytes, msg: bytes):
        t = hashlib.sha256(name).digest()
        return hashlib.sha256(t + t + msg).hexdigest()

    cases = [(b"BIP0340/challenge", b""), (b"my-protocol/v1", b"abc"), ("prot\u00f3colo/\u2713/\U0001f511".encode(), bytes(range(200))),
             (b"x", b"\x00" * 55), (b"x", b"\x00" * 56), (b"x" * 64, b"\xff" * 119), (b"x" * 100, bytes(1000))]
    for name, msg in cases:
        assert run(name, msg) == want(name, msg)
    # refused: empty, and everything Go's strings.ToValidUTF8 would alter
    for bad in [b"", b"\xff", b"ab\xc0\xaf", b"\xed\xa0\x80", b"\xf4\x90\x80\x80", b"\xe2\x82", b"ok\x80", b"\xf8\x88\x80\x80\x80"]:
        assert run(bad, b"m") == "error"
