"""Driven by tests/test_constant_time.py under `ncu`: runs every constant-time entry point once per SECRET set, with
everything that is public (batch size, points, digests, messages) held fixed, so that the per-launch hardware counters
of the same kernel can be compared across the sets.  Prints the order of the sets; the counters come from ncu."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("secp256k1-voi_b200")
N = pkg.synth.N


def rows(vals):
    return np.frombuffer(b"".join(int(v).to_bytes(32, "big") for v in vals), np.uint8).reshape(-1, 32).copy()


def secret_sets(n):
    rnd = np.random.default_rng(12345)
    random = rnd.integers(0, 256, (n, 32), dtype=np.uint8)
    return [
        ("zero", rows([0] * n)),
        ("n_minus_1", rows([N - 1] * n)),
        ("all_ones", rows([2**256 - 1] * n)),                      # >= n: reduced / rejected, still secret-independent
        ("single_bit", rows([1 << (i % 256) for i in range(n)])),
        ("random", random),
        ("small", rows([(i % 7) + 1 for i in range(n)])),
    ]


def main():
    small, big = 4096, 40000        # the lane-split kernel (<= 16384 items) and the throughput kernel (7-bit windows)
    eng = pkg.Engine(device=0, max_batch=big)
    pub_small, _ = eng.scalar_base_mult(pkg.synth.base_mult_scalars(small, start=100))       # fixed PUBLIC points
    pub_small = np.asarray(pub_small).copy()
    pub_small[:8] = pub_small[8:16]                                                           # (rows 0-7 of that stream are edge scalars)
    digest = pkg.synth.base_mult_scalars(small, start=900)
    aux = pkg.synth.base_mult_scalars(small, start=901)
    d_pub, d_dg, d_aux = (torch.from_numpy(x).cuda() for x in (pub_small, digest, aux))
    sets_small = [(nm, torch.from_numpy(v).cuda()) for nm, v in secret_sets(small)]
    sets_big = [(nm, torch.from_numpy(v).cuda()) for nm, v in secret_sets(big)]
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for (nm, k), (_, kb) in zip(sets_small, sets_big):
        eng.scalar_base_mult(k)              # k_base_mult_ct_split + k_finish_affine
        eng.scalar_base_mult(kb)             # k_base_mult_ct
        eng.scalar_mult(k, d_pub)            # k_scalar_mult_ct (GLV, constant time)
        eng.ecdh(k, d_pub)
        eng.ecdsa_sign_rfc6979(k, d_dg)      # nonce generation, k*G, k^-1, s
        eng.schnorr_sign(k, d_dg, d_aux)
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("SETS " + " ".join(nm for nm, _ in sets_small))
    eng.close()


if __name__ == "__main__":
    main()
