"""world_size-2 gloo test (CPU) of the sample-parallel driver: sharding + final all_gather must
reproduce the single-process result.  The per-rank sampler here is the CPU oracle -- the test
covers the host-side partition / gather logic, not the kernels."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from protein_redesign_b200 import synthetic as syn
from protein_redesign_b200.sampling import sample_parallel, shard_rows, unshard_rows


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_sampler(batch):
    """Deterministic per-row function of the batch (row independent, like the network)."""
    x = batch["residue_esm"]
    pos = torch.stack([x[..., :3].sum(1).cos() for _ in range(x.shape[1])], 1)
    logits = x[..., :21] * batch["residue_mask"].unsqueeze(-1)
    return pos, logits


def _worker(rank, world, port, rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    batch = syn.make_batch(syn.TINY, [(3, 9)] * rows, seed=9)
    pos, logits = sample_parallel(_fake_sampler, batch, rows)
    if rank == 0:
        q.put((pos, logits))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_roundtrip():
    t = torch.arange(24.).view(6, 4)
    parts = [shard_rows({"t": t}, r, 3)["t"] for r in range(3)]
    assert torch.equal(unshard_rows(parts, 6), t)


def test_two_rank_gather_equals_single_process():
    rows, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    pos, logits = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    batch = syn.make_batch(syn.TINY, [(3, 9)] * rows, seed=9)
    want_pos, want_logits = _fake_sampler(batch)
    assert torch.equal(pos, want_pos)
    assert torch.equal(logits, want_logits)
