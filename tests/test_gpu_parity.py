"""Parity of the PRODUCT (sm_100a kernels through the C ABI) with the oracle.
Runs on the B200 box: python -m pytest tests -m gpu"""
import hashlib

import numpy as np
import pytest

import parity_suites as ps

pytestmark = pytest.mark.gpu


def test_field_ops_ptx(engine):
    ps.check_field_ops(engine, n=4096)


def test_gen_table_matches_reference_bin(engine):
    ps.check_gen_table(engine)


def test_base_mult_config1(engine, oracle):
    # BASELINE.json configs[0]: 4096 scalars + affine encode, bit-exact
    ps.check_base_mult(engine, oracle, n=4096)


def test_base_mult_large_batch_kernel_edges(engine, oracle):
    # > 16384 items: the 7-bit kernel with its Jacobian accumulator; recoding-boundary scalars against the oracle
    got, st = ps.check_base_mult_edges(engine, oracle, n=20000)
    # and the same rows through the lane-split kernel (complete formulas) must agree bit for bit
    ks = ps.synth.base_mult_scalars(20000, start=1000)
    got2, st2 = engine.scalar_base_mult(ks[:4096])
    assert np.array_equal(got2[:16], got[:16]) and np.array_equal(st2[:16], st[:16])


def test_rfc6979(engine, oracle):
    ps.check_rfc6979_and_kats(engine, oracle)


def test_wycheproof_ecdsa_all(engine, oracle):
    ps.check_wycheproof_ecdsa(engine, oracle)


def test_bip340(engine):
    ps.check_bip340(engine)


def test_ecdsa_synth(engine, oracle):
    ps.check_ecdsa_synth(engine, oracle, n=4096)


def test_schnorr_synth(engine, oracle):
    ps.check_schnorr_synth(engine, oracle, n=2048)


def test_ecdsa_edges(engine, oracle):
    ps.check_ecdsa_edges(engine, oracle)


def test_double_scalar_mult(engine, oracle):
    ps.check_double_scalar_mult(engine, oracle, n=2048)
    ps.check_double_scalar_mult_small_multiples(engine, oracle)


def test_recover(engine, oracle):
    ps.check_recover_synth(engine, oracle, n=512)


def test_scalar_mult_ecdh(engine, oracle):
    ps.check_scalar_mult_ecdh(engine, oracle, n=2048)


def test_wycheproof_ecdh_all(engine, oracle):
    ps.check_wycheproof_ecdh(engine, oracle)


def test_point_decompress(engine, oracle):
    ps.check_point_decompress(engine, oracle, n=1024)


def test_point_compress(engine, oracle):
    ps.check_point_compress(engine, oracle, n=1024)


def test_msm(engine, oracle):
    ps.check_msm(engine, oracle, sizes=(0, 1, 2, 31, 32, 33, 64, 300, 1000, 4096), big=1 << 17, heavy=20000)


def test_msm_sharded(engine, oracle):
    ps.check_msm_sharded(engine, oracle, n=4096, shards=8)


def test_sign_rfc6979(engine, oracle):
    ps.check_sign_rfc6979(engine, oracle, n=2048)
    # sign on the GPU, verify on the GPU, at a size where every kernel runs full waves
    n = 1 << 16
    priv = ps.synth.base_mult_scalars(n, start=77)
    priv[:, 0] &= 0x7F  # keep every key below n
    dg = ps.synth.base_mult_scalars(n, start=99)
    sig, rec, st = engine.ecdsa_sign_rfc6979(priv, dg)
    assert (st == 1).all()
    pk, _ = engine.scalar_base_mult(priv)
    assert engine.ecdsa_verify(pk, dg, sig, 1).all()
    q, qst = engine.ecdsa_recover(dg, np.concatenate([sig, rec[:, None]], axis=1))
    assert (qst == 1).all() and np.array_equal(q, pk)


def test_schnorr_sign(engine, oracle):
    ps.check_schnorr_sign(engine, oracle, n=1024)
    n = 1 << 15
    priv = ps.synth.base_mult_scalars(n, start=11); priv[:, 0] &= 0x7F
    msg = ps.synth.base_mult_scalars(n, start=12)
    aux = ps.synth.base_mult_scalars(n, start=13)
    sig, st = engine.schnorr_sign(priv, msg, aux)
    assert (st == 1).all()
    pk, _ = engine.scalar_base_mult(priv)
    assert engine.schnorr_verify(pk[:, 1:33].copy(), msg, sig).all()


def test_hash_to_curve(engine, oracle, s256):
    ps.check_hash_to_curve(engine, oracle, n=512)
    with pytest.raises(s256.S256Error):
        engine.hash_to_curve(b"", np.zeros((2, 4), np.uint8))   # errInvalidDomainSep


def test_empty_and_ragged(engine):
    z = np.zeros((0, 32), np.uint8)
    out, st = engine.scalar_base_mult(z)
    assert out.shape == (0, 65) and st.shape == (0,)
    assert engine.ecdsa_verify(np.zeros((0, 65), np.uint8), z, np.zeros((0, 64), np.uint8)).shape == (0,)
    with pytest.raises(ValueError):
        engine.ecdsa_verify(np.zeros((2, 65), np.uint8), np.zeros((1, 32), np.uint8), np.zeros((2, 64), np.uint8))
    # sizes around the warp / CTA / inversion-group boundaries
    for n in (1, 15, 16, 17, 31, 33, 127, 129, 1000):
        ks = ps.synth.base_mult_scalars(n, start=100)
        out, st = engine.scalar_base_mult(ks)
        out2, st2 = engine.scalar_base_mult(ks[::-1].copy())
        assert np.array_equal(out, out2[::-1]) and np.array_equal(st, st2[::-1])


def test_device_pointer_path_matches_host_path(engine, oracle):
    import torch
    w = ps.synth.ecdsa_batch(1024, ps.oracle_base_mult(oracle))
    host = engine.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"])
    dev = engine.ecdsa_verify(torch.from_numpy(w["pk65"]).cuda(), torch.from_numpy(w["digest32"]).cuda(),
                              torch.from_numpy(w["sig64"]).cuda())
    torch.cuda.synchronize()
    assert np.array_equal(dev.cpu().numpy(), host)
    assert np.array_equal(host, w["expected"])


def test_chunking_beyond_capacity(s256, oracle):
    # a context with a tiny capacity must give the same answers chunk by chunk
    eng = s256.Engine(max_batch=192)
    try:
        w = ps.synth.ecdsa_batch(700, ps.oracle_base_mult(oracle))
        assert np.array_equal(eng.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"]), w["expected"])
        ks = ps.synth.base_mult_scalars(500)
        out, st = eng.scalar_base_mult(ks)
        exp, est = oracle.batch_scalar_base_mult(ks)
        assert np.array_equal(out, exp) and np.array_equal(st, est)
    finally:
        eng.close()


def test_full_size_properties(engine):
    """BASELINE config 2 size (2^20): valid-by-construction signatures made with
    the engine's own k*G, 1/16 corrupted -> the boolean vector must equal the
    construction; a checksum pins run-to-run determinism."""
    n = 1 << 20
    w = ps.synth.ecdsa_batch(n, engine.scalar_base_mult)
    a = engine.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"])
    assert np.array_equal(a, w["expected"])
    b = engine.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"])
    assert hashlib.sha256(a.tobytes()).digest() == hashlib.sha256(b.tobytes()).digest()
    assert int(a.sum()) == n - n // 16


def _sample_indices(n, expected, count=1 << 14):
    """Every corrupted item plus `count` valid ones spread over the batch."""
    bad = np.nonzero(np.asarray(expected) == 0)[0]
    good = np.nonzero(np.asarray(expected) == 1)[0]
    good = good[:: max(1, len(good) // count)][:count]
    return np.sort(np.concatenate([bad, good]))


def _checked_base_mult(engine, oracle, sample_of):
    """base_mult callable for synth.*_batch: the engine computes all n multiples, the ORACLE recomputes the ones at
    `sample_of(n)` (so that the sampled inputs do not depend on the product), and both must agree there."""
    def f(k32):
        out, st = engine.scalar_base_mult(k32)
        out, st = np.array(out, np.uint8), np.array(st, np.uint8)
        idx = sample_of(len(out))
        eo, es = oracle.batch_scalar_base_mult(np.ascontiguousarray(k32[idx]))
        assert np.array_equal(out[idx], eo) and np.array_equal(st[idx], es)
        out[idx], st[idx] = eo, es
        return out, st
    return f


def _spread(count):
    # the corruption cycle hits i % 16 == 15; this stride also lands on every residue mod 16
    return lambda n: np.unique(np.concatenate([np.arange(15, n, 16), np.arange(0, n, max(1, n // count))]))


def test_full_size_ecdsa_oracle_sampled(engine, oracle):
    """Config 2 at 2^20 as SURVEY.md 8(d) specifies it: all 65 536 corrupted items and 2^14 valid ones go through the
    oracle (multi-threaded), with keys and nonce points for those items taken from the oracle's own base mult."""
    n = 1 << 20
    w = ps.synth.ecdsa_batch(n, _checked_base_mult(engine, oracle, _spread(1 << 14)))
    got = engine.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"])
    assert np.array_equal(got, w["expected"])
    idx = _sample_indices(n, w["expected"])
    assert len(idx) >= (1 << 16) + (1 << 14) - 16
    exp = oracle.batch_ecdsa_verify(np.ascontiguousarray(w["pk65"][idx]), np.ascontiguousarray(w["digest32"][idx]),
                                    np.ascontiguousarray(w["sig64"][idx]))
    assert np.array_equal(got[idx], exp)


def test_full_size_schnorr_oracle_sampled(engine, oracle):
    """Config 3 at 2^20 (BIP-340 incl. lift_x): by construction on everything, the oracle on all corrupted + 2^14 valid."""
    n = 1 << 20
    w = ps.synth.schnorr_batch(n, _checked_base_mult(engine, oracle, _spread(1 << 14)))
    got = engine.schnorr_verify(w["pkx32"], w["msg"], w["sig64"])
    assert np.array_equal(got, w["expected"])
    idx = _sample_indices(n, w["expected"])
    exp = oracle.batch_schnorr_verify(np.ascontiguousarray(w["pkx32"][idx]), np.ascontiguousarray(w["msg"][idx]),
                                      np.ascontiguousarray(w["sig64"][idx]))
    assert np.array_equal(got[idx], exp)


def test_full_size_ecdh_oracle_sampled(engine, oracle):
    """Config 4 at 2^20 (constant-time GLV ScalarMult / ECDH): the closed form (k d) G on everything, the oracle's own
    ScalarMult on 2^14 items."""
    n = 1 << 20
    w = ps.synth.ecdh_batch(n, _checked_base_mult(engine, oracle, _spread(1 << 14)))
    x, st = engine.ecdh(w["k32"], w["pt65"])
    full, fst = engine.scalar_mult(w["k32"], w["pt65"])
    exp, est = engine.scalar_base_mult(w["closed_form_scalar"])
    assert (np.asarray(st) == 1).all() and (np.asarray(fst) == 1).all()
    assert np.array_equal(full, exp) and np.array_equal(x, np.asarray(exp)[:, 1:33])
    idx = np.arange(0, n, n >> 14)
    ox, ost = oracle.batch_ecdh(np.ascontiguousarray(w["k32"][idx]), np.ascontiguousarray(w["pt65"][idx]))
    assert np.array_equal(np.asarray(x)[idx], ox) and (ost == 1).all()
    of, ofst = oracle.batch_scalar_mult(np.ascontiguousarray(w["k32"][idx[:2048]]), np.ascontiguousarray(w["pt65"][idx[:2048]]))
    assert np.array_equal(np.asarray(full)[idx[:2048]], of)


def test_full_size_msm(engine, oracle):
    """Config 5 at 2^20 on one GPU: the closed form (sum s_i d_i) G for the whole product, the oracle's Straus on a
    2^12 prefix, the sharded shape (8 partials -> combine) and the constant-time flavour at n = 4096."""
    n = 1 << 20
    w = ps.synth.msm_batch(n, _checked_base_mult(engine, oracle, lambda m: np.arange(0, m, max(1, m >> 12))))
    got, st = engine.msm(w["k32"], w["pt65"])
    exp, est = oracle.scalar_base_mult(w["closed_form_scalar"])
    assert st == est and got.tobytes() == exp
    m = 1 << 12
    got, st = engine.msm(w["k32"][:m], w["pt65"][:m])
    exp, est = oracle.msm(w["k32"][:m].tobytes(), w["pt65"][:m].tobytes())
    assert (st, got.tobytes()) == (est, exp)
    got, st = engine.msm(w["k32"][:m], w["pt65"][:m], vartime=False)
    assert (st, got.tobytes()) == (est, exp)
    per = n // 8
    parts = []
    for g in range(8):
        p, pst = engine.msm_partial(w["k32"][g * per:(g + 1) * per], w["pt65"][g * per:(g + 1) * per])
        assert pst == 1
        parts.append(p)
    got, st = engine.msm_combine(np.stack(parts))
    exp, est = oracle.scalar_base_mult(w["closed_form_scalar"])
    assert st == est and got.tobytes() == exp


def test_msm_device_resident_and_sharded_entry_points(s256, oracle):
    """s256_msm_dev (inputs and result in HBM, no host sync) and s256_msm_sharded[_dev] against the host-pointer MSM and
    the oracle; the sharded forms run here with a communicator of ONE rank, which still goes through the library's
    pack -> ncclAllGather -> fold path (two ranks need two GPUs: tests/test_multi_gpu.py)."""
    eng = s256.Engine(device=0, max_batch=1 << 16)
    try:
        w = ps.synth.msm_batch(5000, ps.oracle_base_mult(oracle))
        exp, est = oracle.msm(w["k32"].tobytes(), w["pt65"].tobytes())
        dk, dp = torch_cuda(w["k32"]), torch_cuda(w["pt65"])
        for vt in (True, False):
            if not vt:
                dk, dp = dk[:64].contiguous(), dp[:64].contiguous()
                exp, est = oracle.msm(w["k32"][:64].tobytes(), w["pt65"][:64].tobytes(), vartime=False)
            out, st = eng.msm(dk, dp, vartime=vt)
            assert int(st.cpu()[0]) == est and out.cpu().numpy().tobytes() == exp
        # more items than the context's capacity: chunks accumulate on the device
        big = ps.synth.msm_batch((1 << 16) + 777, eng.scalar_base_mult)
        e2, s2 = oracle.scalar_base_mult(big["closed_form_scalar"])
        out, st = eng.msm(torch_cuda(big["k32"]), torch_cuda(big["pt65"]))
        assert int(st.cpu()[0]) == s2 and out.cpu().numpy().tobytes() == e2
        # an undecodable point poisons the result on the device too
        bad = w["pt65"].copy(); bad[9, 64] ^= 1
        out, st = eng.msm(torch_cuda(w["k32"]), torch_cuda(bad))
        assert int(st.cpu()[0]) == 0 and not out.cpu().numpy().any()
        out, st = eng.msm(torch_cuda(w["k32"][:0]), torch_cuda(w["pt65"][:0]))
        assert int(st.cpu()[0]) == 2 and not out.cpu().numpy().any()
        # sharded entry points: no communicator = plain MSM; then a one-rank communicator through NCCL
        exp, est = oracle.msm(w["k32"].tobytes(), w["pt65"].tobytes())
        out, st = eng.msm_sharded(w["k32"], w["pt65"])
        assert (st, out.tobytes()) == (est, exp)
        eng.comm_init(s256.Engine.comm_unique_id(), 0, 1)
        out, st = eng.msm_sharded(w["k32"], w["pt65"])
        assert (st, out.tobytes()) == (est, exp)
        out, st = eng.msm_sharded(torch_cuda(w["k32"]), torch_cuda(w["pt65"]))
        assert int(st.cpu()[0]) == est and out.cpu().numpy().tobytes() == exp
        out, st = eng.msm_sharded(w["k32"], bad)
        assert st == 0 and not out.any()
        eng.comm_free()
    finally:
        eng.close()


def test_calls_on_different_streams_are_ordered_on_the_device(engine):
    """A *_dev call returns once its kernels are enqueued; the next call (another stream, or a host-pointer call) works in
    the same per-context scratch.  The library orders them on the device (ctx.h scratch_guard); without that the first
    call's ladder would read scalars and tables the second one is already overwriting."""
    import torch
    n = 1 << 17
    wa = ps.synth.ecdsa_batch(n, engine.scalar_base_mult, start=0)
    wb = ps.synth.ecdsa_batch(n, engine.scalar_base_mult, start=5 * n + 3, corrupt_every=7)
    da = [torch_cuda(wa[k]) for k in ("pk65", "digest32", "sig64")]
    db = [torch_cuda(wb[k]) for k in ("pk65", "digest32", "sig64")]
    sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    for _ in range(3):
        with torch.cuda.stream(sa):
            oka = engine.ecdsa_verify(*da)
        with torch.cuda.stream(sb):
            okb = engine.ecdsa_verify(*db)
        okc = engine.ecdsa_verify(wa["pk65"][:4096], wa["digest32"][:4096], wa["sig64"][:4096])   # host-pointer call
        with torch.cuda.stream(sa):
            pk, st = engine.scalar_base_mult(torch_cuda(ps.synth.base_mult_scalars(4096)))
        torch.cuda.synchronize()
        assert np.array_equal(oka.cpu().numpy(), wa["expected"])
        assert np.array_equal(okb.cpu().numpy(), wb["expected"])
        assert np.array_equal(okc, wa["expected"][:4096])


def test_host_pipeline_chunk_shapes(s256):
    """The host-pointer verify path is a two-stage pipeline per chunk (api.cu verify_pipelined): exercise a
    chunk cut in 1/8 + 7/8, a second pipelined chunk of odd size, and a short tail chunk on the plain path."""
    eng = s256.Engine(device=0, max_batch=1 << 18)
    try:
        n = (1 << 18) + (1 << 16) + 777
        w = ps.synth.ecdsa_batch(n, eng.scalar_base_mult)
        for m in (n, (1 << 18) + 1000, 1 << 18):
            got = eng.ecdsa_verify(w["pk65"][:m], w["digest32"][:m], w["sig64"][:m])
            assert np.array_equal(got, w["expected"][:m]), m
        d = [torch_cuda(w[k]) for k in ("pk65", "digest32", "sig64")]
        assert np.array_equal(eng.ecdsa_verify(*d).cpu().numpy(), w["expected"])
    finally:
        eng.close()


def test_host_pipeline_other_paths(s256, oracle):
    """The four-part host pipeline (ctx.h pipelined) of the other entry points: a pipelined chunk of 2^18
    plus a short second chunk, checked by closed forms / round trips on everything and by the oracle on a
    sample that straddles every sub-chunk boundary."""
    cap = 1 << 18
    eng = s256.Engine(device=0, max_batch=cap)
    try:
        n = cap + 5000
        cuts = [0, cap // 16, cap // 4, cap // 8 * 5, cap, n]
        sample = sorted({min(max(c + d, 0), n - 1) for c in cuts for d in (-2, -1, 0, 1, 127, 128)})
        ks = ps.synth.base_mult_scalars(n)
        pts, st = eng.scalar_base_mult(ks)
        eo, es = oracle.batch_scalar_base_mult(ks[sample])
        assert np.array_equal(pts[sample], eo) and np.array_equal(st[sample], es)
        we = ps.synth.ecdh_batch(n, eng.scalar_base_mult)
        x, xst = eng.ecdh(we["k32"], we["pt65"])
        exp, _ = eng.scalar_base_mult(we["closed_form_scalar"])
        assert np.array_equal(x, exp[:, 1:33]) and (xst == 1).all()
        full, fst = eng.scalar_mult(we["k32"], we["pt65"])
        assert np.array_equal(full, exp) and (fst == 1).all()
        ws = ps.synth.schnorr_batch(n, eng.scalar_base_mult)
        assert np.array_equal(eng.schnorr_verify(ws["pkx32"], ws["msg"], ws["sig64"]), ws["expected"])
        w = ps.synth.ecdsa_batch(n, eng.scalar_base_mult, corrupt_every=0)
        priv = np.frombuffer(b"".join(ps.synth._nonzero_mod_n(v).to_bytes(32, "big")
                                      for v in ps.synth._stream_ints(b"key", 0, n, ps.synth.SEED)), np.uint8).reshape(n, 32)
        sig, rec, sst = eng.ecdsa_sign_rfc6979(priv, w["digest32"])
        assert (sst == 1).all()
        es_, er_, est_ = oracle.batch_ecdsa_sign_rfc6979(priv[sample], w["digest32"][sample])
        assert np.array_equal(sig[sample], es_) and np.array_equal(rec[sample], er_)
        assert eng.ecdsa_verify(w["pk65"], w["digest32"], sig).all()
        pk, pst = eng.ecdsa_recover(w["digest32"], np.concatenate([sig, rec.reshape(-1, 1)], axis=1))
        assert np.array_equal(pk, w["pk65"]) and (pst == 1).all()
        ssig, sst2 = eng.schnorr_sign(priv, w["digest32"], ks)
        pub, _ = eng.scalar_base_mult(priv)
        assert (sst2 == 1).all() and eng.schnorr_verify(pub[:, 1:33].copy(), w["digest32"], ssig).all()
        u1 = ps.synth.base_mult_scalars(n, start=5)
        d, dst_ = eng.double_scalar_mult_basepoint_vartime(u1, we["k32"], we["pt65"])
        do, dso = oracle.batch_double_scalar_mult(u1[sample], we["k32"][sample], we["pt65"][sample])
        assert np.array_equal(d[sample], do) and np.array_equal(dst_[sample], dso)
    finally:
        eng.close()


def test_init_out_of_memory_is_an_error_code(s256, oracle):
    """A context too large for the device comes back as an error (no abort, nothing leaked), and the
    next context works."""
    with pytest.raises(s256.S256Error) as ei:
        s256.Engine(device=0, max_batch=1 << 28)   # 2^28 items x 1.7 KB of scratch > 180 GB
    assert "out of memory" in str(ei.value) or "rc=-4" in str(ei.value)
    eng = s256.Engine(device=0, max_batch=1024)
    try:
        ks = ps.synth.base_mult_scalars(16)
        out, st = eng.scalar_base_mult(ks)
        exp, est = oracle.batch_scalar_base_mult(ks)
        assert np.array_equal(out, exp) and np.array_equal(st, est)
    finally:
        eng.close()


def test_pinned_buffers(s256, oracle):
    """s256_host_alloc: page-locked inputs and engine-owned page-locked results give the same bytes as the
    pageable path; a result view is only reused by the next call of the same method."""
    eng = s256.Engine(device=0, max_batch=4096, pinned_outputs=True)
    try:
        ks = ps.synth.base_mult_scalars(3000)
        pk = eng.pinned_empty((3000, 32))
        pk[:] = ks
        out, st = eng.scalar_base_mult(pk)
        exp, est = oracle.batch_scalar_base_mult(ks)
        assert np.array_equal(out, exp) and np.array_equal(st, est)
        keep = out.copy()
        out2, _ = eng.scalar_base_mult(pk[:100])
        assert np.array_equal(out2, keep[:100])
        x, xst = eng.ecdh(pk[8:3000], keep[8:3000])   # another method: its own buffers
        assert np.array_equal(out2, keep[:100]) and (xst == 1).all()
    finally:
        eng.close()


def test_misaligned_device_inputs(engine, oracle):
    """The kernels use 128-bit loads when a row is 16-byte aligned and must fall back to byte loads when the
    caller's device pointer is not (a tensor view at an odd offset)."""
    import torch

    def odd(a):
        a = np.ascontiguousarray(a)
        buf = torch.empty(a.size + 1, dtype=torch.uint8, device="cuda")
        v = buf[1:].view(a.shape)
        v.copy_(torch.from_numpy(a))
        assert v.data_ptr() % 2 == 1 and v.is_contiguous()
        return v
    n = 700
    w = ps.synth.ecdsa_batch(n, ps.oracle_base_mult(oracle))
    ok = engine.ecdsa_verify(odd(w["pk65"]), odd(w["digest32"]), odd(w["sig64"]))
    assert np.array_equal(ok.cpu().numpy(), w["expected"])
    ks = ps.synth.base_mult_scalars(n)
    out, st = engine.scalar_base_mult(odd(ks))
    exp, est = oracle.batch_scalar_base_mult(ks)
    assert np.array_equal(out.cpu().numpy(), exp) and np.array_equal(st.cpu().numpy(), est)
    ws = ps.synth.schnorr_batch(n, ps.oracle_base_mult(oracle))
    ok = engine.schnorr_verify(odd(ws["pkx32"]), odd(ws["msg"]), odd(ws["sig64"]))
    assert np.array_equal(ok.cpu().numpy(), ws["expected"])


def torch_cuda(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_concurrent_callers_share_one_context(engine, oracle):
    """cgo calls arrive on arbitrary OS threads: one context, many threads, mixed entry points."""
    import threading
    w = ps.synth.ecdsa_batch(3000, ps.oracle_base_mult(oracle))
    ks = ps.synth.base_mult_scalars(2000)
    exp_pts, exp_st = oracle.batch_scalar_base_mult(ks)
    errors = []

    def verify_loop():
        try:
            for _ in range(6):
                assert np.array_equal(engine.ecdsa_verify(w["pk65"], w["digest32"], w["sig64"]), w["expected"])
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    def sbm_loop():
        try:
            for _ in range(6):
                out, st = engine.scalar_base_mult(ks)
                assert np.array_equal(out, exp_pts) and np.array_equal(st, exp_st)
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=f) for f in (verify_loop, sbm_loop, verify_loop, sbm_loop)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
