"""Synthetic Falcon inputs: there is no Falcon signer in this
environment, and the circuit only sees (sig, pk, hm) coefficient vectors in [0, q).

Per signature: h uniform in [0,q)^N (the public key polynomial), s1, s2 rounded
Gaussians resampled until ||s1||^2+||s2||^2 < SIG_L2_BOUND, sig = s2 mod q,
hm = s1 + s2*h mod (q, x^N+1) so that v = hm - sig*h = s1 (falcon_ntt.rs:47-49).
RNG: numpy Philox keyed by (seed, first index); deterministic for a given call.
"""
import numpy as np

Q = 12289
SIG_L2_BOUND = {9: 34034726, 10: 70265242}  # range_proofs.rs:104,196
SIGMA = {9: 165.7366171829776, 10: 168.38857144654395}


def bitrev(x, bits):
    r = 0
    for i in range(bits):
        r |= ((x >> i) & 1) << (bits - 1 - i)
    return r


def ntt_table(n):
    """NTT_TABLE[i] = 7^bitrev10(i) mod q (script/ntt_param.sage:3-132)."""
    return np.array([pow(7, bitrev(i, 10), Q) for i in range(n)], dtype=np.int64)


def ntt(a, logn):
    """Clear-text forward NTT, loop shape of gadgets/poly.rs:115-149 reduced mod q.
    a: [..., N] int64 array; vectorised over leading dims."""
    n = 1 << logn
    tab = ntt_table(n)
    out = np.array(a, dtype=np.int64) % Q
    t = n
    for l in range(logn):
        m = 1 << l
        ht = t // 2
        x = out.reshape(out.shape[:-1] + (m, 2, ht))
        s = tab[m:2 * m].reshape((m, 1))
        u = x[..., 0, :]
        v = x[..., 1, :] * s % Q
        out = np.stack([(u + v) % Q, (u - v) % Q], axis=-2).reshape(out.shape)
        t = ht
    return out


def intt(a, logn):
    """Inverse of ntt() (Gentleman-Sande with inverse twiddles, scaled by 1/N)."""
    n = 1 << logn
    tab = ntt_table(n)
    inv_tab = np.array([pow(int(x), Q - 2, Q) for x in tab], dtype=np.int64)
    out = np.array(a, dtype=np.int64) % Q
    t = 1
    for l in reversed(range(logn)):
        m = 1 << l
        x = out.reshape(out.shape[:-1] + (m, 2, t))
        s = inv_tab[m:2 * m].reshape((m, 1))
        u = x[..., 0, :]
        v = x[..., 1, :]
        out = np.stack([(u + v) % Q, (u - v) * s % Q], axis=-2).reshape(out.shape)
        t *= 2
    return out * pow(n, Q - 2, Q) % Q


def poly_mul(a, b, logn):
    """a*b in Z_q[x]/(x^N+1) via the NTT."""
    return intt(ntt(a, logn) * ntt(b, logn) % Q, logn)


def make_signatures(logn, count, seed=1, first=0):
    """Returns (sig, pk, hm) as uint16 arrays of shape [count, N]."""
    n = 1 << logn
    rng = np.random.Generator(np.random.Philox(key=[seed, first]))
    pk = rng.integers(0, Q, size=(count, n), dtype=np.int64)
    s1 = np.zeros((count, n), dtype=np.int64)
    s2 = np.zeros((count, n), dtype=np.int64)
    todo = np.arange(count)
    while todo.size:
        a = np.rint(rng.normal(0.0, SIGMA[logn], size=(todo.size, n))).astype(np.int64)
        b = np.rint(rng.normal(0.0, SIGMA[logn], size=(todo.size, n))).astype(np.int64)
        ok = (a * a).sum(1) + (b * b).sum(1) < SIG_L2_BOUND[logn]
        ok &= (np.abs(a).max(1) < 6144) & (np.abs(b).max(1) < 6144)
        s1[todo[ok]] = a[ok]
        s2[todo[ok]] = b[ok]
        todo = todo[~ok]
    sig = s2 % Q
    hm = (s1 + poly_mul(sig, pk, logn)) % Q
    return sig.astype(np.uint16), pk.astype(np.uint16), hm.astype(np.uint16)
