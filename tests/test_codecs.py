"""Host-side codecs in front of the batch: strict-DER signature parsing
(secec.ParseASN1Signature) and the BIP-66 check, pinned by EVERY Wycheproof ECDSA case (996) and the
25 BIP-66 vectors; then Verify(EncodingASN1) / bitcoin.VerifyASN1 end to end on the GPU."""
import numpy as np
import pytest

from conftest import load_golden

H = bytes.fromhex


def test_parse_asn1_matches_reference_expectations(s256):
    der = load_golden("wycheproof_ecdsa_der.json")["cases"]
    reach = {(c["src"], c["tcId"]): c for c in load_golden("wycheproof_ecdsa.json")["cases"]}
    sig, ok = s256.parse_asn1_signatures([H(c["sig"]) for c in der])
    assert len(der) == 996
    n_ok = 0
    for c, s, o in zip(der, sig, ok):
        key = (c["src"], c["tcId"])
        if c["valid"]:
            assert o == 1, c  # a case the reference verifies must parse
        if o:
            n_ok += 1
            # the independent Python parser (tests/golden/make_golden.py) agrees on acceptance and on r, s
            assert key in reach, c
            assert s.tobytes().hex() == reach[key]["r"] + reach[key]["s"]
        else:
            assert key not in reach, c
            assert not s.any()
    assert n_ok == 432


def test_parse_asn1_hand_cases(s256):
    good = H("3006020101020102")
    rows = [good, good + b"\x00", H("30060201010201"), H("3007020101020200 02".replace(" ", "")), H("300702020001020102"),
            H("3006020181020102"), H("3081060201010201 02".replace(" ", "")), H("3006020100020102"), b"", H("30"),
            H("3026022100" + "ff" * 32 + "020101"),  # r >= n
            H("3025022001" + "00" * 31 + "020101")]
    sig, ok = s256.parse_asn1_signatures(rows)
    assert ok.tolist() == [1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1]
    assert sig[0, 31] == 1 and sig[0, 63] == 2


def test_bip66_vectors(s256):
    doc = load_golden("bip66.json")
    valid = [H(v["der"]) + bytes([69]) for v in doc["valid"]]
    invalid = [H(v["der"]) + bytes([69]) for v in doc["invalid"]]
    assert s256.is_valid_signature_encoding_bip0066(valid).tolist() == [1] * 9
    assert s256.is_valid_signature_encoding_bip0066(invalid).tolist() == [0] * 16
    # parseASN1SignatureShitcoin: r, s as listed, except case 8 (r = s = 0 is rejected by the scalar check)
    sig, ok = s256.parse_asn1_signatures([H(v["der"]) for v in doc["valid"]])
    for i, v in enumerate(doc["valid"]):
        r, s = int(v["r"], 16), int(v["s"], 16)
        if i == 8:
            assert r == 0 and s == 0 and ok[i] == 0
        else:
            assert ok[i] == 1 and sig[i].tobytes() == r.to_bytes(32, "big") + s.to_bytes(32, "big")
    # too short / too long / empty rows
    assert s256.is_valid_signature_encoding_bip0066([b"", b"\x30" * 8, b"\x30" * 74]).tolist() == [0, 0, 0]


@pytest.mark.gpu
def test_verify_asn1_all_wycheproof(engine):
    der = load_golden("wycheproof_ecdsa_der.json")["cases"]
    pk = np.frombuffer(b"".join(H(c["pk"]) for c in der), np.uint8).reshape(-1, 65)
    dg = np.frombuffer(b"".join(H(c["digest"])[:32] for c in der), np.uint8).reshape(-1, 32)
    got = engine.ecdsa_verify_asn1(pk, dg, [H(c["sig"]) for c in der])
    exp = np.array([c["valid"] for c in der], np.uint8)
    bad = np.nonzero(got != exp)[0]
    assert len(bad) == 0, [der[i] for i in bad[:3]]


@pytest.mark.gpu
def test_bitcoin_verify_asn1(engine, oracle, s256):
    der = load_golden("wycheproof_ecdsa_der.json")["cases"]
    der = [c for c in der if "sha256" in c["src"]]
    pk = np.frombuffer(b"".join(H(c["pk"]) for c in der), np.uint8).reshape(-1, 65)
    dg = np.frombuffer(b"".join(H(c["digest"])[:32] for c in der), np.uint8).reshape(-1, 32)
    rows = [H(c["sig"]) + b"\x01" for c in der]
    got = engine.bitcoin_verify_asn1(pk, dg, rows)
    # expectation assembled from the pinned parts: BIP-66 syntax, strict DER parse, low-s ECDSA (oracle)
    bip = s256.is_valid_signature_encoding_bip0066(rows)
    sig, ok = s256.parse_asn1_signatures([H(c["sig"]) for c in der])
    low = oracle.batch_ecdsa_verify(pk, dg, sig, 1)
    exp = bip & ok & low
    assert np.array_equal(got, exp)
    assert 0 < int(exp.sum()) < int(np.array([c["valid"] for c in der]).sum())  # high-s valid cases are refused
