"""N>1 host logic on CPU: world_size-2 gloo processes shard a batch of independent proof jobs round-robin, reduce their
step times with max-over-ranks and gather the results on rank 0 (the GPU path uses the same code with nccl)."""
import hashlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import zk_symmetric_crypto_b200.sharding as sh


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_prove(job):
    # stands in for Backend.prove_chacha20_raw on a box without a GPU: deterministic bytes per job
    return hashlib.blake2s(repr(job).encode()).digest()


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        jobs = [("chacha20", seed, 4) for seed in range(7)]
        assert sh.shard_indices(len(jobs), world, rank) == list(range(rank, 7, world))
        out = sh.prove_batch(jobs, _fake_prove)
        ms = sh.max_over_ranks([10.0 + rank, 5.0 - rank])
        q.put((rank, out, ms))
    finally:
        dist.destroy_process_group()


def test_shard_gather_and_max_over_ranks_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {}
    for _ in range(2):
        rank, out, ms = q.get(timeout=120)
        res[rank] = (out, ms)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    jobs = [("chacha20", seed, 4) for seed in range(7)]
    assert res[0][0] == [_fake_prove(j) for j in jobs]
    assert res[1][0] is None
    assert res[0][1] == [11.0, 5.0] and res[1][1] == [11.0, 5.0]


def test_single_process_paths():
    assert sh.world() == (1, 0)
    assert sh.shard_indices(5, 1, 0) == [0, 1, 2, 3, 4]
    assert sh.max_over_ranks([1.5]) == [1.5]
    assert sh.prove_batch([1, 2, 3], lambda j: bytes([j])) == [b"\x01", b"\x02", b"\x03"]
