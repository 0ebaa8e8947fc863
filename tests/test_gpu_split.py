"""GPU parity for the split proving key (single proof over several GPUs): the shards are emulated as
contexts on one device (frcs_load_pk_shard k of n), each runs frcs_prove_partial_dev, and
frcs_combine_partials must give the oracle's proof bytes."""
import numpy as np
import pytest

from falcon_r1cs_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_shards", [2, 3])
def test_split_key_proof_equals_oracle(circuits, oracle, n_shards):
    import torch
    logn, n = 9, 2
    c = circuits(logn, 0)
    P = c.setup(seed=2000 + n_shards)
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")
    pk = api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                        a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"),
                        b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"), l_query=P.export("l_query"))
    sig, pkk, hm = synth.make_signatures(logn, n, seed=61)
    rng = np.random.default_rng(4)
    r = np.stack([api.fr_rand(rng) for _ in range(n)])
    s = np.stack([api.fr_rand(rng) for _ in range(n)])
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(x.view(np.int16)).to(dev) for x in (sig, pkk, hm)]
    d_r, d_s = [torch.from_numpy(x.view(np.int64)).to(dev) for x in (r, s)]
    parts = []
    for k in range(n_shards):
        ctx = api.Context(logn)
        try:
            ctx.load_pk_shard(pk, k, n_shards)
            d_part = torch.zeros((n, api.PARTIAL_WORDS), dtype=torch.int64, device=dev)
            d_st = torch.zeros(n, dtype=torch.int32, device=dev)
            ctx.prove_partial_dev(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d_r.data_ptr(), d_s.data_ptr(),
                                  d_part.data_ptr(), d_st.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert int(d_st.abs().sum()) == 0
            parts.append(d_part.cpu().numpy().view(np.uint64))
        finally:
            ctx.close()
    proofs = api.combine_partials(np.stack(parts), r, s)
    for i in range(n):
        z, _, _ = c.witness(sig[i], pkk[i], hm[i])
        want, want_bytes = c.prove(P, z, r[i], s[i])
        assert (proofs[i] == want).all()
        assert api.proof_compress(proofs[i]) == bytes(want_bytes)


@pytest.mark.slow
def test_split_key_schoolbook_1024_proof_equals_oracle(circuits, oracle):
    """BASELINE configs[3]: one Falcon-1024 verify-with-schoolbook proof (1,156,150 constraints, domain 2^21) with the
    proving key split by base range (two shards here, emulated on one device): the combined proof is byte-identical
    to the oracle's create_proof under the same (r, s).  (Was asserted only inside tools/bench_split.py.)"""
    import torch
    from falcon_r1cs_b200 import lib as L
    logn, n_shards = 10, 2
    c = circuits(logn, 1)
    assert (c.n_inst, c.n_wit, c.n_cons) == (2049, 1150004, 1156150)  # README.md:45
    P = c.setup(seed=2100)
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")
    pk = api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                        a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"),
                        b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"), l_query=P.export("l_query"))
    sig, pkk, hm = synth.make_signatures(logn, 1, seed=62)
    rng = np.random.default_rng(5)
    r, s = api.fr_rand(rng)[None], api.fr_rand(rng)[None]
    dev = torch.device("cuda", 0)
    d = [torch.from_numpy(x.view(np.int16)).to(dev) for x in (sig, pkk, hm)]
    d_r, d_s = [torch.from_numpy(x.view(np.int64)).to(dev) for x in (r, s)]
    parts = []
    for k in range(n_shards):
        ctx = api.Context(logn, kind=L.KIND_SCHOOLBOOK)
        try:
            ctx.load_pk_shard(pk, k, n_shards)
            d_part = torch.zeros((1, api.PARTIAL_WORDS), dtype=torch.int64, device=dev)
            d_st = torch.zeros(1, dtype=torch.int32, device=dev)
            ctx.prove_partial_dev(1, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d_r.data_ptr(), d_s.data_ptr(),
                                  d_part.data_ptr(), d_st.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert int(d_st.abs().sum()) == 0
            parts.append(d_part.cpu().numpy().view(np.uint64))
        finally:
            ctx.close()
    proofs = api.combine_partials(np.stack(parts), r, s)
    z, st, _ = c.witness(sig[0], pkk[0], hm[0])
    assert st == 0
    want, want_bytes = c.prove(P, z, r[0], s[0])
    assert (proofs[0] == want).all()
    assert api.proof_compress(proofs[0]) == bytes(want_bytes)


@pytest.mark.parametrize("n_shards", [2, 3, 4])
def test_split_begin_finish_shared_witness_map(circuits, oracle, n_shards):
    """frcs_prove_split_begin_dev / _finish_dev: besides the key, the three ifft + coset_fft transforms of the witness
    map are divided between the shards (vector v on shard v % n_shards).  The exchange (ncclBroadcast in bench.py) is
    emulated with device copies between the shards' buffers; the combined proof equals the oracle's bytes.  Run twice
    on the same contexts: the entry points keep per-context state between the two halves."""
    import torch
    logn = 9
    c = circuits(logn, 0)
    P = c.setup(seed=2200 + n_shards)
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")
    pk = api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                        a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"),
                        b_g2_query=P.export("b_g2_query"), h_query=P.export("h_query"), l_query=P.export("l_query"))
    dev = torch.device("cuda", 0)
    cur = torch.cuda.current_stream().cuda_stream
    ctxs = []
    try:
        for k in range(n_shards):
            ctxs.append(api.Context(logn))
            ctxs[-1].load_pk_shard(pk, k, n_shards)
        dom = 1 << ctxs[0].domain_log2
        abc = [torch.full((3, dom, 4), -1, dtype=torch.int64, device=dev) for _ in range(n_shards)]
        for rep in range(2):
            sig, pkk, hm = synth.make_signatures(logn, 1, seed=63 + rep)
            rng = np.random.default_rng(6 + rep)
            r, s = api.fr_rand(rng)[None], api.fr_rand(rng)[None]
            d = [torch.from_numpy(x.view(np.int16)).to(dev) for x in (sig, pkk, hm)]
            d_r, d_s = [torch.from_numpy(x.view(np.int64)).to(dev) for x in (r, s)]
            d_st = torch.zeros((n_shards, 1), dtype=torch.int32, device=dev)
            d_part = torch.zeros((n_shards, 1, api.PARTIAL_WORDS), dtype=torch.int64, device=dev)
            for k, ctx in enumerate(ctxs):
                ctx.prove_split_begin_dev(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d_r.data_ptr(),
                                          d_s.data_ptr(), abc[k].data_ptr(), d_st[k].data_ptr(), cur)
            for v in range(3):                      # the broadcast of vector v from its owner
                for k in range(n_shards):
                    if k != v % n_shards:
                        abc[k][v].copy_(abc[v % n_shards][v])
            for k, ctx in enumerate(ctxs):
                ctx.prove_split_finish_dev(abc[k].data_ptr(), d_part[k].data_ptr(), cur)
            torch.cuda.synchronize()
            assert int(d_st.abs().sum()) == 0
            proofs = api.combine_partials(d_part.cpu().numpy().view(np.uint64), r, s)
            z, _, _ = c.witness(sig[0], pkk[0], hm[0])
            want, want_bytes = c.prove(P, z, r[0], s[0])
            assert (proofs[0] == want).all()
            assert api.proof_compress(proofs[0]) == bytes(want_bytes)
    finally:
        for ctx in ctxs:
            ctx.close()
