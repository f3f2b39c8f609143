"""world_size-2 gloo test of the scheduler's shard -> all-gather -> un-permute logic (CPU)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chatterbox_embed_b200 import scheduler, synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards = scheduler.partition(lengths, world)
    mine = shards[rank]
    # stand-in "embedding": a function of the clip index only, so the gather/un-permute is checkable
    local = torch.stack([torch.full((scheduler.EMB,), float(i)) + torch.arange(scheduler.EMB) / 1000 for i in mine])
    full = scheduler.gather_embeddings(local, shards, len(lengths))
    want = torch.stack([torch.full((scheduler.EMB,), float(i)) + torch.arange(scheduler.EMB) / 1000 for i in range(len(lengths))])
    q.put((rank, bool(torch.equal(full, want))))
    dist.destroy_process_group()


def test_gather_world2():
    lengths = synth.ragged_lengths(37)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]
