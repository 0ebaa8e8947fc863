"""Falcon wire formats -> the coefficient vectors the circuit consumes.

The reference obtains them from falcon-rust ([EXT], floating git dependency):
  Polynomial::from(&PublicKey)                      circuits/falcon_ntt.rs:28
  Polynomial::from(&Signature), sig.nonce()         circuits/falcon_ntt.rs:27,44
  Polynomial::from_hash_of_message(msg, nonce)      circuits/falcon_ntt.rs:44   (api.hash_to_point)
falcon-rust wraps the Falcon round-3 reference C code, so the formats restated here are those of the
Falcon specification (section 3.11): public key = header 0x00 + logn, then N coefficients of 14 bits,
big-endian bit packing; signature = header 0x30 + logn ("compressed" encoding), 40-byte nonce, then s2 in
the compressed format: per coefficient a sign bit, the low 7 bits of |s|, and |s| >> 7 in unary (that many
zero bits followed by a one); unused trailing bits must be zero.  Coefficients of s2 are lifted to [0, q)
as the reference's Polynomial does.  Parity is unpinned against falcon-rust itself (not in the tree); the
encoders below exist for round-trip tests.
"""
import numpy as np

Q = 12289
NONCE_LEN = 40
SIG_BYTES = {9: 666, 10: 1280}     # padded signature sizes (falcon-512 / falcon-1024)
PK_BYTES = {9: 897, 10: 1793}


class FalconFormatError(ValueError):
    pass


def decode_public_key(pk_bytes: bytes):
    """-> (logn, h) with h: uint16[N] in [0, q)"""
    if len(pk_bytes) < 1:
        raise FalconFormatError("empty public key")
    logn = pk_bytes[0]
    if logn not in (9, 10) or len(pk_bytes) != PK_BYTES[logn]:
        raise FalconFormatError("bad public key header or length")
    n = 1 << logn
    bits = np.unpackbits(np.frombuffer(pk_bytes, dtype=np.uint8)[1:])
    vals = bits[:14 * n].reshape(n, 14).astype(np.uint32)
    h = (vals << np.arange(13, -1, -1, dtype=np.uint32)).sum(axis=1)
    if (h >= Q).any():
        raise FalconFormatError("public key coefficient out of range")
    return logn, h.astype(np.uint16)


def encode_public_key(logn, h):
    h = np.asarray(h, dtype=np.uint32)
    bits = ((h[:, None] >> np.arange(13, -1, -1, dtype=np.uint32)) & 1).astype(np.uint8).reshape(-1)
    return bytes([logn]) + np.packbits(bits).tobytes()


def decode_signature(sig_bytes: bytes):
    """-> (logn, nonce, s2) with s2: uint16[N] in [0, q) (negative coefficients lifted by q)"""
    if len(sig_bytes) < 1 + NONCE_LEN:
        raise FalconFormatError("signature too short")
    head = sig_bytes[0]
    logn = head & 0x0F
    if head & 0xF0 != 0x30 or logn not in (9, 10):
        raise FalconFormatError("bad signature header (expected compressed encoding 0x30 + logn)")
    n = 1 << logn
    nonce = bytes(sig_bytes[1:1 + NONCE_LEN])
    bits = np.unpackbits(np.frombuffer(sig_bytes, dtype=np.uint8)[1 + NONCE_LEN:])
    out = np.zeros(n, dtype=np.int64)
    pos, total = 0, bits.size
    for i in range(n):
        if pos + 8 > total:
            raise FalconFormatError("truncated signature")
        sign = int(bits[pos])
        low = 0
        for b in bits[pos + 1:pos + 8]:
            low = (low << 1) | int(b)
        pos += 8
        high = 0
        while True:
            if pos >= total:
                raise FalconFormatError("truncated signature")
            if bits[pos]:
                pos += 1
                break
            high += 1
            pos += 1
            if high > 15:
                raise FalconFormatError("coefficient too large")
        mag = (high << 7) | low
        if sign and mag == 0:
            raise FalconFormatError("negative zero")
        out[i] = -mag if sign else mag
    if bits[pos:].any():
        raise FalconFormatError("non-zero padding")
    return logn, nonce, (out % Q).astype(np.uint16)


def encode_signature(logn, nonce, s2_signed, padded=True):
    """s2_signed: integers in (-2048, 2048).  Returns the compressed encoding, zero-padded to the fixed size."""
    assert len(nonce) == NONCE_LEN
    bits = []
    for v in np.asarray(s2_signed, dtype=np.int64):
        mag = int(abs(v))
        bits.append(1 if v < 0 else 0)
        bits.extend((mag >> k) & 1 for k in range(6, -1, -1))
        bits.extend([0] * (mag >> 7))
        bits.append(1)
    body = np.packbits(np.array(bits, dtype=np.uint8)).tobytes()
    out = bytes([0x30 + logn]) + nonce + body
    if padded:
        if len(out) > SIG_BYTES[logn]:
            raise FalconFormatError("signature does not fit the padded size")
        out += bytes(SIG_BYTES[logn] - len(out))
    return out
