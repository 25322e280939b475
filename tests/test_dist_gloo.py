"""world_size-2 gloo test of the multi-GPU host logic: contiguous sharding + the final gather (even and ragged)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_total, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "oakink2-tamf_b200"))
    from tamf_b200 import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    r = shard.shard_range(n_total, rank, world)
    # each "sample" is its global sequence index broadcast over [99,1,4]
    local = torch.tensor(list(r), dtype=torch.float32).view(-1, 1, 1, 1).expand(-1, 99, 1, 4).contiguous()
    full = shard.gather_samples(local, n_total)
    ok = full.shape == (n_total, 99, 1, 4) and torch.equal(full[:, 0, 0, 0], torch.arange(n_total, dtype=torch.float32))
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # the max-over-ranks timing reduction bench.py uses
    ok = ok and float(t) == float(world)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in ps)
    [p.join(60) for p in ps]
    assert res == {0: True, 1: True}


def test_gather_even_shards():
    _run(8)


def test_gather_ragged_shards():
    _run(7)
