"""-m gpu tests that need TWO GPUs of one box (NCCL over NVLink): the sample-parallel driver with the real model, the
Lightning-free predict loop under a process group, and the data-parallel gradient all-reduce of training_step.  Skipped on
a single-GPU box (run them with ``gpurun --gpus 2``)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import dataclasses
    import sys

    import torch.distributed as dist
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    sys.path.insert(0, here)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from protein_redesign_b200 import synthetic as syn
    from protein_redesign_b200.model import ProteinReDiffModel
    from protein_redesign_b200.sampling import sample_parallel_model
    out = {}
    # ---- sample-parallel == single process, bit for bit, with injected noise --------------------------------------
    T, rows = 4, 4
    cfg = dataclasses.replace(syn.README, num_steps=T, mask_prob=0.3)
    model = ProteinReDiffModel(cfg)
    model.load_state_dict(syn.make_state_dict(cfg, 5), strict=True)
    model = model.to(dev).eval()
    host = syn.make_batch(cfg, [(8, 32), (6, 27), (7, 30), (5, 33)], seed=5)
    B, N = host["atom_mask"].shape
    g = torch.Generator().manual_seed(99)
    noise = {"z_T": torch.randn(B, N, 3, generator=g), "seq_T": torch.randn(B, N, 21, generator=g),
             "steps": torch.randn(T - 1, B, N, 3, generator=g)}
    to_dev = lambda b: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
    torch.manual_seed(7)  # every rank draws the same joint residue mask
    timings = {}
    pos, logits = sample_parallel_model(model, to_dev(host), noise={k: v.to(dev) for k, v in noise.items()}, timings=timings)
    if rank == 0:
        torch.manual_seed(7)
        want_pos, want_logits = model.sample(to_dev(host), noise=noise)
        out["sample_parallel_pos_equal"] = bool(torch.equal(pos, want_pos))
        out["sample_parallel_logits_equal"] = bool(torch.equal(logits, want_logits))
        out["gather_us"] = timings.get("gather_us")
    # ---- predict() under a process group: batches rank::world, results gathered in dataloader order -----------------
    from torch.utils.data import DataLoader
    from protein_redesign_b200.predict import RepeatDataset, collate_fn, predict
    item = syn.make_complex(cfg, 6, 26, seed=3)
    dl = DataLoader(RepeatDataset(item, 4), batch_size=1, collate_fn=collate_fn)
    res = predict(model, dl)
    if rank == 0:
        out["predict_batches"] = len(res)
        out["predict_finite"] = all(bool(torch.isfinite(p).all()) for p, _ in res)
    # ---- data-parallel training: all-reduced gradient == mean of the two ranks' gradients ----------------------------
    tcfg = dataclasses.replace(syn.README, mask_prob=0.15, num_steps=2000)
    tm = ProteinReDiffModel(tcfg)
    tm.load_state_dict(syn.make_state_dict(tcfg, 6), strict=True)
    tm = tm.to(dev).train()
    tb = syn.make_batch(tcfg, [(8, 32), (6, 27)], seed=20 + rank, with_positions=True)
    torch.manual_seed(rank)
    loss = tm.training_step(to_dev(tb), 0)
    loss.backward()
    name = "Denoiser.folding_blocks.1.pair_fc.1.weight"
    mine = dict(tm.named_parameters())[name].grad.clone()
    both = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(both, mine)
    red = mine.clone()
    dist.all_reduce(red)
    red /= world
    if rank == 0:
        want = (both[0] + both[1]) / 2
        out["allreduce_rel"] = float((red - want).norm() / want.norm())
        out["ranks_differ"] = float((both[0] - both[1]).norm() / both[0].norm())
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
def test_two_gpu_sample_parallel_predict_and_gradient_allreduce():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    print("two-GPU results:", out)
    assert out["sample_parallel_pos_equal"] and out["sample_parallel_logits_equal"], out
    assert out["predict_batches"] == 4 and out["predict_finite"], out
    assert out["allreduce_rel"] < 1e-6 and out["ranks_differ"] > 1e-3, out
