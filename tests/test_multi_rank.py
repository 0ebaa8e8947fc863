"""The single-proof multi-GPU path on the host side, world_size 2 over gloo (no GPU):
each rank forms the MSM sums of its base-range shard (here with the oracle's CPU MSM),
the ranks all_gather the 144-word partials, and frcs_combine_partials (host code of the
product library) adds the shards and finishes create_proof.  The result must be the
oracle's proof under the same (r, s).  On the B200 box the per-shard sums come from
frcs_prove_partial_dev instead (tests/test_gpu_split.py) and the gather is NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def fq_one():
    v = (1 << 384) % P_MOD
    return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(6)], dtype=np.uint64)


def to_xyzz(aff, g2):
    w = 12 if g2 else 6
    out = np.zeros(8 * w // 2 * 1, dtype=np.uint64) if False else np.zeros(4 * w, dtype=np.uint64)
    if not aff.any():
        return out
    out[:2 * w] = aff
    out[2 * w:2 * w + 6] = fq_one()
    out[3 * w:3 * w + 6] = fq_one()
    return out


def shard_partial(O, api, c, P, z, h, r, s, shard, n_shards):
    """MSM sums of one shard, computed with the oracle: A | B1 | L+H | unused | B2"""
    rg = api.shard_ranges(c.n_inst, c.n_wit, c.domain_log2, shard, n_shards)
    zc, hc = O.fr_to_canonical(z), O.fr_to_canonical(h)
    rc, sc = O.fr_to_canonical(r)[0], O.fr_to_canonical(s)[0]
    R = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    ri, si = O.limbs_to_ints(rc.reshape(1, 4))[0], O.limbs_to_ints(sc.reshape(1, 4))[0]
    neg_rs = O.ints_to_limbs([(-ri * si) % R])[0]
    one = O.ints_to_limbs([1])[0]
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")  # alpha, beta, delta | beta2, delta2, gamma2
    (zl, zh), (ll, lh), (hl, hh) = rg["z"], rg["l"], rg["h"]
    first = shard == 0

    def msm(fn, bases, scalars, extra_b, extra_s):
        if first:
            bases = np.concatenate([bases] + [b.reshape(1, -1) for b in extra_b])
            scalars = np.concatenate([scalars] + [x.reshape(1, 4) for x in extra_s])
        return fn(bases, scalars)

    A = msm(O.msm_g1, P.export("a_query")[zl:zh], zc[zl:zh], [g1[0], g1[2]], [one, rc])
    B1 = msm(O.msm_g1, P.export("b_g1_query")[zl:zh], zc[zl:zh], [g1[1], g1[2]], [one, sc])
    B2 = msm(O.msm_g2, P.export("b_g2_query")[zl:zh], zc[zl:zh], [g2[0], g2[1]], [one, sc])
    wit = zc[c.n_inst:]
    LH = msm(O.msm_g1, np.concatenate([P.export("l_query")[ll:lh], P.export("h_query")[hl:hh]]),
             np.concatenate([wit[ll:lh], hc[hl:hh]]), [g1[2]], [neg_rs])
    out = np.zeros(api.PARTIAL_WORDS, dtype=np.uint64)
    out[0:24] = to_xyzz(A, False)
    out[24:48] = to_xyzz(B1, False)
    out[48:72] = to_xyzz(LH, False)
    out[96:144] = to_xyzz(B2, True)
    return out


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from falcon_r1cs_b200 import api, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        logn = 9
        c = O.Circuit(logn, 0)
        P = c.setup(seed=77)
        sig, pk, hm = synth.make_signatures(logn, 1, seed=91)
        z, st, _ = c.witness(sig[0], pk[0], hm[0])
        assert st == 0
        h = c.witness_map(z)
        rng = np.random.default_rng(12)
        r, s = api.fr_rand(rng), api.fr_rand(rng)
        mine = shard_partial(O, api, c, P, z, h, r, s, rank, world)
        t = torch.from_numpy(mine.view(np.int64))
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t)
        if rank == 0:
            parts = np.stack([g.numpy().view(np.uint64) for g in gathered]).reshape(world, 1, api.PARTIAL_WORDS)
            proof = api.combine_partials(parts, r, s)[0]
            want, want_bytes = c.prove(P, z, r, s)
            ok = bool((proof == want).all()) and api.proof_compress(proof) == bytes(want_bytes)
            ok = ok and c.verify_trapdoor(P, z, r, s, proof)
            q.put(("ok" if ok else "mismatch", rank))
        else:
            q.put(("ok", rank))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put(("error: %s\n%s" % (e, traceback.format_exc()), rank))
    finally:
        dist.destroy_process_group()


def test_shard_ranges_cover_everything():
    sys.path.insert(0, ROOT)
    from falcon_r1cs_b200 import api
    for n_sh in (1, 2, 3, 8):
        for key, total in (("z", 1025 + 78386), ("l", 78386), ("h", (1 << 17) - 1)):
            prev = 0
            for k in range(n_sh):
                lo, hi = api.shard_ranges(1025, 78386, 17, k, n_sh)[key]
                assert lo == prev and hi >= lo
                prev = hi
            assert prev == total


def test_split_proof_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[0] == "ok" for r in res), res
