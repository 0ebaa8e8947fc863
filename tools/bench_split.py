#!/usr/bin/env python
"""Single-proof latency with the proving key split by base range over N GPUs (BASELINE.json configs[3] shape,
here on the Falcon-1024 NTT circuit): every rank recomputes z and h, runs its slice of the MSMs
(frcs_prove_partial_dev), one NCCL all_gather moves 144 u64 per proof and rank, rank 0 finishes the proof.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_split.py [--logn 10] [--proofs 1]
Prints one JSON line on rank 0 (ms per proof = max over ranks, device time + gather + host tail)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--logn", type=int, default=10)
    ap.add_argument("--proofs", type=int, default=1)
    ap.add_argument("--kind", type=int, default=0, help="0 = verify-with-NTT circuit, 1 = schoolbook circuit")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import oracle_lib as O
    from falcon_r1cs_b200 import api, synth
    O.lib().orc_set_num_threads(max(1, (os.cpu_count() or 1) // world))
    c = O.Circuit(a.logn, a.kind)
    P = c.setup(7)  # trusted-setup stand-in, identical on every rank
    g1, g2 = P.export("g1_elems"), P.export("g2_elems")
    pk = api.ProvingKey(alpha_g1=g1[0], beta_g1=g1[1], delta_g1=g1[2], beta_g2=g2[0], delta_g2=g2[1],
                        a_query=P.export("a_query"), b_g1_query=P.export("b_g1_query"), b_g2_query=P.export("b_g2_query"),
                        h_query=P.export("h_query"), l_query=P.export("l_query"))
    ctx = api.Context(a.logn, kind=a.kind, device=local)
    ctx.load_pk_shard(pk, rank, world)
    n = a.proofs
    sig, pkk, hm = synth.make_signatures(a.logn, n, seed=5)  # same inputs on every rank
    rng = np.random.default_rng(1)
    r = np.stack([api.fr_rand(rng) for _ in range(n)]); s = np.stack([api.fr_rand(rng) for _ in range(n)])
    d = [torch.from_numpy(x.view(np.int16)).to(dev) for x in (sig, pkk, hm)]
    d_r, d_s = [torch.from_numpy(x.view(np.int64)).to(dev) for x in (r, s)]
    d_part = torch.zeros((n, api.PARTIAL_WORDS), dtype=torch.int64, device=dev)
    d_all = torch.zeros((world, n, api.PARTIAL_WORDS), dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        ctx.prove_partial_dev(n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d_r.data_ptr(), d_s.data_ptr(),
                              d_part.data_ptr(), d_st.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all.view(-1), d_part.view(-1))   # the one NCCL call of the path
        else:
            d_all.copy_(d_part.view(1, n, -1))
        if rank == 0:
            return api.combine_partials(d_all.cpu().numpy().view(np.uint64), r, s)
        torch.cuda.synchronize()
        return None

    proofs = step()
    if rank == 0:
        z, _, _ = c.witness(sig[0], pkk[0], hm[0])
        want, _ = c.prove(P, z, r[0], s[0])
        assert (proofs[0] == want).all(), "split proof differs from the oracle's"
    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        print(json.dumps({"metric": "falcon%d_%s_split_key_proof_latency_ms" % (1 << a.logn, "schoolbook" if a.kind else "ntt"), "constraints": c.n_cons, "value": dt * 1e3 / (a.steps * n),
                          "unit": "ms/proof", "n_gpus": world, "proofs_per_step": n, "steps": a.steps, "byte_identical_to_oracle": True,
                          "gather": "ncclAllGather of %d bytes per rank" % (n * api.PARTIAL_WORDS * 8) if world > 1 else "none"}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
