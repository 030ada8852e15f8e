"""Sample-parallel generation over the GPUs of one node.

Independent samples / batch rows shard across ranks with no data-path collective; the only
exchange is one final all_gather of ``pos [B_r, N, 3]`` and ``logits [B_r, N, 21]`` (a few MB,
NCCL over NVLink on GPUs, gloo in the CPU tests).

Partitioning follows what Lightning's DistributedSampler does for the reference's batch scripts
(scripts/predict_batch_strc_msk_inp.py:209-225): rank r owns rows r, r+world, r+2*world, ...
Residue masking draws jointly over the whole batch (mask_utils.py:82-99), so ``prepare`` runs on
the full batch BEFORE sharding and the per-rank sampler receives an already prepared shard.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist


def shard_rows(batch: Dict[str, object], rank: int, world: int) -> Dict[str, object]:
    """Rows rank::world of every batched tensor (lists are sliced the same way)."""
    out = {}
    for k, v in batch.items():
        if isinstance(v, torch.Tensor) and v.dim() >= 1:
            out[k] = v[rank::world].contiguous()
        elif isinstance(v, (list, tuple)):
            out[k] = list(v[rank::world])
        else:
            out[k] = v
    return out


def unshard_rows(parts, total_rows: int) -> torch.Tensor:
    """Inverse of shard_rows for the gathered per-rank results."""
    world = len(parts)
    out = parts[0].new_empty((total_rows,) + tuple(parts[0].shape[1:]))
    for r, p in enumerate(parts):
        out[r::world] = p
    return out


def sample_parallel(sample_fn: Callable[[Dict[str, object]], Tuple[torch.Tensor, torch.Tensor]],
                    prepared_batch: Dict[str, object], rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Run ``sample_fn`` on this rank's shard and all_gather the results (every rank returns the full set).

    ``rows`` (the global batch size) must be divisible by the world size so that all ranks exchange
    equally shaped tensors.
    """
    if not (dist.is_available() and dist.is_initialized()):
        return sample_fn(prepared_batch)
    world, rank = dist.get_world_size(), dist.get_rank()
    if rows % world != 0:
        raise ValueError(f"global batch {rows} is not divisible by world size {world}")
    pos, logits = sample_fn(shard_rows(prepared_batch, rank, world))
    pos_parts = [torch.empty_like(pos) for _ in range(world)]
    log_parts = [torch.empty_like(logits) for _ in range(world)]
    dist.all_gather(pos_parts, pos.contiguous())
    dist.all_gather(log_parts, logits.contiguous())
    return unshard_rows(pos_parts, rows), unshard_rows(log_parts, rows)


def shard_noise(noise: Optional[Dict[str, torch.Tensor]], rank: int, world: int) -> Optional[Dict[str, torch.Tensor]]:
    """Rows rank::world of injected sampler noise (``z_T`` [B,N,3], ``seq_T`` [B,N,21], ``steps`` [T-1,B,N,3])."""
    if noise is None:
        return None
    return {k: (v[:, rank::world] if k == "steps" else v[rank::world]).contiguous() for k, v in noise.items()}


def sample_parallel_model(model, batch: Dict[str, object], noise: Optional[Dict[str, torch.Tensor]] = None,
                          timings: Optional[dict] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """What ``Trainer(strategy='ddp').predict`` does for the reference's batch scripts
    (scripts/predict_batch_strc_msk_inp.py:209-229), with the real model: ``prepare_batch`` on the FULL batch (the
    residue-masking draw is joint over all rows, mask_utils.py:82-99 -- every rank must enter with the same CPU RNG state,
    exactly as every DDP rank seeds identically), shard rows rank::world, ``model.sample(shard, prepared=True)`` (the shard
    is NOT masked a second time), one final all_gather.  ``timings`` (a dict) receives the gather time in microseconds."""
    if not model.setup_schedule:
        model.run_setup_schedule()
        model.setup_schedule = True
    full = model.prepare_batch(batch)
    rows = int(full["atom_mask"].shape[0])
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    if rows % world != 0:
        raise ValueError(f"global batch {rows} is not divisible by world size {world}")
    shard = shard_rows(full, rank, world) if world > 1 else full
    pos, logits = model.sample(shard, noise=shard_noise(noise, rank, world) if world > 1 else noise, prepared=True)
    if world == 1:
        return pos, logits
    pos_parts = [torch.empty_like(pos) for _ in range(world)]
    log_parts = [torch.empty_like(logits) for _ in range(world)]
    on_gpu = pos.is_cuda
    if on_gpu and timings is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.all_gather(pos_parts, pos.contiguous())
    dist.all_gather(log_parts, logits.contiguous())
    if on_gpu and timings is not None:
        e1.record()
        torch.cuda.synchronize(pos.device)
        timings["gather_us"] = e0.elapsed_time(e1) * 1e3
    return unshard_rows(pos_parts, rows), unshard_rows(log_parts, rows)
