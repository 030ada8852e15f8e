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


# ---- the real prepare_batch + the real sharding, with the network replaced by a row-wise function (CPU has no kernels) ----
class _HostModel:
    """ProteinReDiffModel's host-side halves (prepare_batch, schedule) with ``sample`` swapped for a deterministic row-wise
    function of the PREPARED shard: it fails if the shard is masked a second time (sample() used to call
    prepare_batch again on an already prepared shard)."""

    def __init__(self, cfg):
        from protein_redesign_b200.model import ProteinReDiffModel
        self.inner = ProteinReDiffModel(cfg)
        self.setup_schedule = True
        self.prepare_calls = 0

    def prepare_batch(self, batch):
        self.prepare_calls += 1
        return self.inner.prepare_batch(batch)

    def sample(self, batch, noise=None, prepared=False):
        assert prepared, "sample_parallel_model must hand over an already prepared shard"
        keep = batch["residue_extra_mask"]
        pos = torch.stack([batch["residue_esm"].sum(-1), keep, batch["residue_inv_extra_mask"]], -1)
        if noise is not None:
            pos = pos + noise["z_T"]
        return pos, batch["residue_one_hot"].float()


def _worker_model(rank, world, port, rows, q):
    import dataclasses
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from protein_redesign_b200.sampling import sample_parallel_model
    cfg = dataclasses.replace(syn.TINY, mask_prob=0.4)
    batch = syn.make_batch(cfg, [(3, 9), (2, 7), (4, 8), (1, 11)][:rows], seed=9)
    g = torch.Generator().manual_seed(5)
    noise = {"z_T": torch.randn(rows, batch["atom_mask"].shape[1], 3, generator=g)}
    torch.manual_seed(77)  # every rank seeds identically, as under DDP
    m = _HostModel(cfg)
    pos, logits = sample_parallel_model(m, batch, noise=noise)
    assert m.prepare_calls == 1
    if rank == 0:
        q.put((pos, logits))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_real_prepare_batch_is_joint_and_applied_once():
    import dataclasses
    rows, world = 4, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_model, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    pos, logits = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    from protein_redesign_b200.sampling import sample_parallel_model
    cfg = dataclasses.replace(syn.TINY, mask_prob=0.4)
    batch = syn.make_batch(cfg, [(3, 9), (2, 7), (4, 8), (1, 11)], seed=9)
    g = torch.Generator().manual_seed(5)
    noise = {"z_T": torch.randn(rows, batch["atom_mask"].shape[1], 3, generator=g)}
    torch.manual_seed(77)
    want_pos, want_logits = sample_parallel_model(_HostModel(cfg), batch, noise=noise)  # no process group: one process
    assert torch.equal(pos, want_pos)
    assert torch.equal(logits, want_logits)
    # the joint draw masked 40 % of ALL residues of the batch exactly once
    dropped = (want_pos - noise["z_T"])[..., 2].round().sum()
    assert int(dropped) == int(int(batch["residue_mask"].sum()) * 0.4)
