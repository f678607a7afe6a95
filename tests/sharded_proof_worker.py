"""Worker of tests/test_gpu_sharded.py (run under torchrun, one rank per GPU): the ranks prove ONE ChaCha20 trace together
(column-sharded transforms, NCCL all-to-all of LDE tile row shards, row-sharded leaf hashing and constraint evaluation) and
rank 0 checks that the proof equals the single-GPU proof byte for byte."""
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


def main():
    L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    check = sys.argv[2] if len(sys.argv) > 2 else "single"     # "single": compare with the single-GPU proof; "ref": the
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3       # reference's own verifier accepts it (sizes one GPU cannot hold)
    rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid = torch.tensor(list(backend.comm_unique_id()), dtype=torch.uint8, device="cuda")
    dist.broadcast(uid, 0)
    be = z.Backend(local_rank)
    be.comm_init(rank, world, bytes(uid.cpu().tolist()))
    key, nonce, counter, pt, ct = bench.synth_inputs(L, 0)       # same inputs on every rank
    ptb, ctb = pt.tobytes(), ct.tobytes()
    times = []
    for it in range(iters):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        proof = be.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
        torch.cuda.synchronize()
        dist.barrier()
        times.append(time.perf_counter() - t0)
    if rank == 0 and check == "ref":
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import base64
        import ref_wasm
        t0 = time.perf_counter()
        v = ref_wasm.verify_chacha20_proof(base64.b64encode(proof).decode(), nonce, counter, ptb, ctb)
        assert v == {"algorithm": "chacha20", "valid": True}, v
        print("SHARDED_OK log_n_rows=%d ranks=%d proof_bytes=%d sharded_ms=%s reference_verifier=accepted (%.1f s)" %
              (L, world, len(proof), ["%.1f" % (t * 1e3) for t in times], time.perf_counter() - t0))
    elif rank == 0:
        single = z.Backend(local_rank)
        want = single.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
        t0 = time.perf_counter()
        single.prove_chacha20_raw(key, nonce, counter, ptb, ctb)
        t1 = time.perf_counter() - t0
        assert proof == want, "sharded proof differs from the single-GPU proof"
        print("SHARDED_OK log_n_rows=%d ranks=%d proof_bytes=%d sharded_ms=%.1f single_gpu_ms=%.1f" %
              (L, world, len(proof), min(times) * 1e3, t1 * 1e3))
    else:
        assert proof == b""
    be.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
