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
