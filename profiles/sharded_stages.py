"""Under torchrun: the ranks prove ONE trace together; every rank prints its per-stage device times and counters.
usage: sharded_stages.py LOG [iters]"""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import zk_symmetric_crypto_b200 as z
from zk_symmetric_crypto_b200 import backend

L = int(sys.argv[1]) if len(sys.argv) > 1 else 20
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid = torch.tensor(list(backend.comm_unique_id()), dtype=torch.uint8, device="cuda")
dist.broadcast(uid, 0)
be = z.Backend(local_rank)
be.comm_init(rank, world, bytes(uid.cpu().tolist()))
key, nonce, counter, pt, ct = bench.synth_inputs(L, 0)
ptb, ctb = pt.tobytes(), ct.tobytes()
times = []
for it in range(iters):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    proof = be.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
    torch.cuda.synchronize()
    dist.barrier()
    times.append((time.perf_counter() - t0) * 1e3)
be.set_profile(True)
t0 = time.perf_counter()
be.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
tp = (time.perf_counter() - t0) * 1e3
st = be.stage_times()
be.set_profile(False)
for r in range(world):
    dist.barrier()
    if r == rank:
        print(json.dumps({"rank": rank, "log": L, "world": world, "ms": [round(t, 1) for t in times], "profiled_ms": round(tp, 1),
                          "stages": {k: round(v, 2) for k, v in st.items()}, "sum_stages": round(sum(st.values()), 1),
                          "counters": be.counters()}), flush=True)
be.comm_destroy()
dist.barrier()
dist.destroy_process_group()
