"""Falcon wire formats (falcon_r1cs_b200/falcon_codec.py): round trips, malformed inputs, and the end of the chain:
decoded (pk, sig) + hash_to_point(nonce || msg) give coefficient vectors the circuit accepts."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth
from falcon_r1cs_b200 import falcon_codec as fc

Q = 12289


@pytest.mark.parametrize("logn", [9, 10])
def test_public_key_round_trip(logn):
    rng = np.random.default_rng(logn)
    h = rng.integers(0, Q, 1 << logn).astype(np.uint16)
    blob = fc.encode_public_key(logn, h)
    assert len(blob) == fc.PK_BYTES[logn] and blob[0] == logn
    l2, h2 = fc.decode_public_key(blob)
    assert l2 == logn and (h2 == h).all()
    with pytest.raises(fc.FalconFormatError):
        fc.decode_public_key(blob[:-1])
    bad = bytearray(blob)
    bad[1] = 0xFF
    bad[2] = 0xFF  # first coefficient = 0x3fff >= q
    with pytest.raises(fc.FalconFormatError):
        fc.decode_public_key(bytes(bad))


@pytest.mark.parametrize("logn", [9, 10])
def test_signature_round_trip_and_malformed(logn):
    rng = np.random.default_rng(100 + logn)
    s2 = np.rint(rng.normal(0, 165, 1 << logn)).astype(np.int64)
    s2[:4] = [0, -1, 127, -128]
    nonce = bytes(rng.integers(0, 256, 40, dtype=np.uint8))
    blob = fc.encode_signature(logn, nonce, s2)
    assert len(blob) == fc.SIG_BYTES[logn] and blob[0] == 0x30 + logn
    l2, n2, got = fc.decode_signature(blob)
    assert l2 == logn and n2 == nonce and (got == (s2 % Q)).all()
    with pytest.raises(fc.FalconFormatError):
        fc.decode_signature(bytes([0x20 + logn]) + blob[1:])       # wrong encoding tag
    with pytest.raises(fc.FalconFormatError):
        fc.decode_signature(blob[:100])                            # truncated
    bad = bytearray(blob)
    bad[-1] |= 1
    with pytest.raises(fc.FalconFormatError):
        fc.decode_signature(bytes(bad))                            # non-zero padding


def test_circuit_from_wire_formats_matches_direct_inputs(oracle):
    """pk / sig bytes -> build_circuit: same (sig, pk, hm) vectors as the direct construction, and for a
    consistent triple the oracle's circuit is satisfied"""
    logn, n = 9, 512
    c = oracle.Circuit(logn, 0)
    rng = np.random.default_rng(5)
    h = rng.integers(0, Q, n).astype(np.int64)
    msg, nonce = b"testing message", bytes(range(40))
    hm = api.hash_to_point(nonce, msg, n).astype(np.int64)
    # pick s2 small, then s1 = hm - s2*h: a signature in the algebraic sense only if s1 is small too, so instead
    # build the triple the other way round for the satisfiability check
    sig, pk, hm2 = synth.make_signatures(logn, 1, seed=9)
    s2_signed = np.where(sig[0] > Q // 2, sig[0].astype(np.int64) - Q, sig[0].astype(np.int64))
    circ = api.FalconNTTVerificationCircuit.from_bytes(fc.encode_public_key(logn, pk[0]), msg,
                                                       fc.encode_signature(logn, nonce, s2_signed))
    assert (circ.pk == pk[0]).all() and (circ.sig == sig[0]).all() and circ.nonce == nonce
    assert (circ.hm == hm).all() and circ.hm.max() < Q
    z, st, fu = c.witness(sig[0], pk[0], hm2[0], construct_matrices=True)
    assert st == 0 and fu == -1


def test_hash_to_point_known_properties():
    a = api.hash_to_point(b"\x00" * 40, b"msg", 512)
    b = api.hash_to_point(b"\x00" * 40, b"msg", 1024)
    assert a.shape == (512,) and b.shape == (1024,) and a.max() < Q
    assert (b[:512] == a).all()                      # same SHAKE stream, longer output
    assert (api.hash_to_point(b"\x01" + b"\x00" * 39, b"msg", 512) != a).any()
