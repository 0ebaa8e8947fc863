"""Split-proof latency (BASELINE configs[3]) under torchrun, for a few variants of the shard geometry and of the
witness-map division.  python -m torch.distributed.run --nproc-per-node N tools/split_latency.py [variant ...]
variant = <shared 0|1>:<lh window bits>:<z window bits>"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

import bench

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
variants = sys.argv[1:] or ["0:16:8", "1:16:8"]
for v in variants:
    shared, lh, z = v.split(":")
    os.environ["FRCS_LH_WINDOW_BITS"] = lh
    os.environ["FRCS_Z_WINDOW_BITS"] = z
    out = bench.split_proof(rank, world, local, shared_witness_map=shared == "1", reps=10)
    if rank == 0:
        st = {k: round(x, 3) for k, x in out["stages_rank0_ms"].items()}
        print(json.dumps({"variant": v, "n_gpus": world, "ms_per_proof": round(out["ms_per_proof"], 3), "stages": st}),
              flush=True)
dist.destroy_process_group()
