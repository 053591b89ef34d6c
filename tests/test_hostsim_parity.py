"""Kernel LOGIC parity on CPU: the item functions of csrc/kernels.cuh compiled
with g++ (portable limb arithmetic) and driven through the same pipelines as
api.cu, checked against the oracle and the golden vectors.  The product (PTX
arithmetic, real launches) is checked by test_gpu_parity.py on the B200."""
import numpy as np
import pytest

import hostsim as hs
import parity_suites as ps


class Backend:
    ecdsa_verify = staticmethod(hs.ecdsa_verify)
    ecdsa_recover = staticmethod(hs.ecdsa_recover)
    schnorr_verify = staticmethod(hs.schnorr_verify)
    double_scalar_mult_basepoint_vartime = staticmethod(hs.double_scalar_mult)
    scalar_base_mult = staticmethod(hs.scalar_base_mult)
    scalar_mult = staticmethod(hs.scalar_mult)
    ecdh = staticmethod(hs.ecdh)
    point_decompress = staticmethod(hs.point_decompress)
    msm = staticmethod(hs.msm)
    msm_partial = staticmethod(hs.msm_partial)
    msm_combine = staticmethod(hs.msm_combine)
    ecdsa_sign_rfc6979 = staticmethod(hs.ecdsa_sign_rfc6979)
    schnorr_sign = staticmethod(hs.schnorr_sign)
    hash_to_curve = staticmethod(hs.hash_to_curve)
    expand_message_xmd = staticmethod(hs.expand_message_xmd)
    debug_field_op = staticmethod(hs.field_op)
    debug_gen_table = staticmethod(hs.gen_table)


be = Backend()


def test_field_ops():
    ps.check_field_ops(be, n=64)


def test_gen_table_matches_reference_bin():
    ps.check_gen_table(be)


def test_base_mult(oracle):
    ps.check_base_mult(be, oracle, n=32)
    ps.check_base_mult(be, oracle, n=8)   # n <= 8 takes the lane-split form in the simulation
    ps.check_base_mult_edges(be, oracle, n=700)


def test_rfc6979(oracle):
    ps.check_rfc6979_and_kats(be, oracle)


def test_wycheproof_ecdsa_subset(oracle):
    ps.check_wycheproof_ecdsa(be, oracle, limit=8)


def test_bip340():
    ps.check_bip340(be)


def test_ecdsa_synth(oracle):
    ps.check_ecdsa_synth(be, oracle, n=64)


def test_schnorr_synth(oracle):
    ps.check_schnorr_synth(be, oracle, n=32)


def test_ecdsa_edges(oracle):
    ps.check_ecdsa_edges(be, oracle)


def test_double_scalar_mult(oracle):
    ps.check_double_scalar_mult(be, oracle, n=40)
    ps.check_double_scalar_mult(be, oracle, n=96)   # with the crafted mid-ladder collision rows
    ps.check_double_scalar_mult_small_multiples(be, oracle)


def test_recover(oracle):
    ps.check_recover_synth(be, oracle, n=16)


def test_scalar_mult_ecdh(oracle):
    ps.check_scalar_mult_ecdh(be, oracle, n=40)


def test_wycheproof_ecdh_subset(oracle):
    ps.check_wycheproof_ecdh(be, oracle, limit=12)


def test_point_decompress(oracle):
    ps.check_point_decompress(be, oracle, n=16)


def test_msm(oracle):
    ps.check_msm(be, oracle, sizes=(0, 1, 2, 31, 32, 33, 64, 200), heavy=4500)


def test_msm_window_sizes(oracle):
    # every Pippenger window width the planner can pick, incl. the carry-only top window (c | 256)
    w = ps.synth.msm_batch(48, ps.oracle_base_mult(oracle))
    exp, est = oracle.msm(w["k32"].tobytes(), w["pt65"].tobytes())
    for c in (4, 5, 7, 8, 11, 13, 16):
        got, st = hs.msm(w["k32"], w["pt65"], force_c=c)
        assert (st, got.tobytes()) == (est, exp), c


def test_msm_sharded(oracle):
    ps.check_msm_sharded(be, oracle, n=128, shards=4)


def test_sign_rfc6979(oracle):
    ps.check_sign_rfc6979(be, oracle, n=24)


def test_schnorr_sign(oracle):
    ps.check_schnorr_sign(be, oracle, n=16)


def test_hash_to_curve(oracle):
    ps.check_hash_to_curve(be, oracle, n=8)
