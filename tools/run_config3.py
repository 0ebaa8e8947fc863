#!/usr/bin/env python
"""BASELINE.json configs[2] at full size: witness generation + R1CS satisfaction for 65,536 synthetic Falcon-1024
signatures, sharded by signature over the ranks (no collective), through the host entry point
frcs_witness_check_batch (H2D of the inputs inside the timed region; the 333 GB of assignments never exist at once).

  python tools/run_config3.py [--count 65536] [--logn 10]
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_config3.py
Inputs: 4096 distinct synthetic signatures (falcon_r1cs_b200/synth.py), repeated to the requested count."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=65536)
    ap.add_argument("--logn", type=int, default=10)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from falcon_r1cs_b200 import api, synth
    ctx = api.Context(a.logn, device=local)
    per = a.count // world
    base = min(4096, per)
    sig, pk, hm = synth.make_signatures(a.logn, base, seed=33, first=rank)
    reps = (per + base - 1) // base
    sig, pk, hm = [np.tile(x, (reps, 1))[:per] for x in (sig, pk, hm)]
    ctx.witness_check_batch(sig[:64], pk[:64], hm[:64])  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fu, st = ctx.witness_check_batch(sig, pk, hm)
    dt = time.perf_counter() - t0
    ok = bool((fu == -1).all() and (st == 0).all())
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    if rank == 0:
        print(json.dumps({"config": "BASELINE configs[2]", "signatures": per * world, "n_gpus": world, "seconds": dt,
                          "witnesses_per_s": per * world / dt, "all_satisfied": ok,
                          "h2d_bytes": per * world * 3 * 2 * (1 << a.logn)}), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
