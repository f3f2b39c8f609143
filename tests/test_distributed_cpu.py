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


class _FakeEmbedder:
    """Stands in for SpeakerEmbedder on CPU: the 'embedding' of a clip is a function of its samples, so the whole job plumbing
    (partition -> batches -> stream -> gather -> un-permute) is checkable without a GPU."""

    def embed_stream(self, batches, pinned=False, **kw):
        for flat, off in batches:
            n = len(off) - 1
            ve = np.stack([np.full(256, flat[off[i]:off[i + 1]].sum(), np.float32) for i in range(n)])
            xv = np.stack([np.full(192, float(off[i + 1] - off[i]), np.float32) for i in range(n)])
            yield ve, xv, np.zeros(n, np.int32)


def _job_worker(rank, world, port, lengths, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.RandomState(0)
    clips = [rng.randn(int(n)).astype(np.float32) for n in lengths]            # every rank can rebuild every clip
    shards = scheduler.partition(lengths, world)
    mine = shards[rank]
    my_len = np.asarray(lengths)[mine]
    fetch = lambda b, i0, i1: np.concatenate([clips[j] for j in mine[i0:i1]])
    local, status = scheduler.embed_shard(_FakeEmbedder(), fetch, my_len, max_clips=4, max_samples=3 * 16000 * 20, pinned=False)
    full = scheduler.gather_embeddings(torch.from_numpy(local), shards, len(lengths)).numpy()
    ok = all(np.allclose(full[i, 0], clips[i].sum(), rtol=1e-5, atol=1e-4) and full[i, 300] == float(lengths[i]) for i in range(len(lengths)))
    q.put((rank, bool(ok and (status == 0).all())))
    dist.destroy_process_group()


def test_voice_bank_job_world2():
    """BASELINE configs[3] on CPU: cbx_partition -> per-rank batches (batch_bounds) -> embed_shard -> one all-gather -> clip order."""
    lengths = [int(x) for x in synth.ragged_lengths(23, lo_s=0.05, hi_s=0.4)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_job_worker, args=(r, 2, port, lengths, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(timeout=60) for p in procs]
    assert res == [(0, True), (1, True)]


def test_batch_bounds():
    rng = np.random.RandomState(3)
    for _ in range(50):
        lens = rng.randint(1, 1000, size=rng.randint(1, 60))
        mc, ms = int(rng.randint(1, 9)), int(rng.randint(1, 3000))
        b = scheduler.batch_bounds(lens, mc, ms)
        assert b[0][0] == 0 and b[-1][1] == len(lens) and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        for i0, i1 in b:
            assert 1 <= i1 - i0 <= mc and (i1 - i0 == 1 or lens[i0:i1].sum() <= ms)
            # greedy: the next clip would not have fitted
            if i1 < len(lens) and i1 - i0 < mc:
                assert lens[i0:i1 + 1].sum() > ms


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
