"""Lightning-free prediction loop and batch assembly.

``predict(model, dataloader)`` reproduces what ``pl.Trainer.predict`` does for the reference's entry points
(generate.py:145-159, scripts/predict_batch_strc_msk_inp.py:209-225): every batch is moved to the model's device,
``model.predict_step(batch, batch_idx)`` runs without autograd (under the EMA weights, reference model.py:249-252) and the
per-batch results are returned as a list.  With a ``torch.distributed`` process group every rank takes batches
``rank::world`` -- Lightning's DistributedSampler for ``strategy='ddp'`` -- and ``gather=True`` all-gathers the results so
that every rank returns all of them in dataloader order.

``collate_fn`` / ``RepeatDataset`` / ``InferenceDataset`` restate the batch assembly either side of the hot path
(reference ProteinReDiff/data.py:80-142,145-170): ligand atoms first, then residues, then padding; ``residue_type`` is
shifted by +1 so that 0 means padding / unknown.  Pinned bit-exactly by tests/golden/collate.npz.
"""
from __future__ import annotations

from typing import Any, Iterable, List, Mapping, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch.utils.data import Dataset
from torch.utils.data.dataloader import default_collate


def collate_fn(data_list: Sequence[Mapping[str, Any]]) -> Mapping[str, Any]:
    """reference data.py:80-142."""
    N = max(d["num_atoms"] + d["num_residues"] for d in data_list)
    batch = {}
    for k, v in data_list[0].items():
        if k.startswith("atom_"):
            pad = (0, 0) * (v.dim() - 1)
            batch[k] = default_collate([F.pad(d[k], pad + (0, N - d["num_atoms"])) for d in data_list])
        elif k.startswith("bond_"):
            pad = (0, 0) * (v.dim() - 2)
            batch[k] = default_collate([F.pad(d[k], pad + (0, N - d["num_atoms"]) * 2) for d in data_list])
        elif k.startswith("residue_"):
            pad = (0, 0) * (v.dim() - 1)
            shift = 1 if k.endswith("_type") else 0  # residue types move to 1..20, 0 = padding / X (data.py:100)
            batch[k] = default_collate([
                F.pad(d[k] + shift if shift else d[k], pad + (d["num_atoms"], N - d["num_atoms"] - d["num_residues"]))
                for d in data_list])
        elif k.endswith("_mol"):
            batch[k] = [d[k] for d in data_list]
        else:
            batch[k] = default_collate([d[k] for d in data_list])
    return batch


class RepeatDataset(Dataset):
    """reference data.py:145-155: the same complex ``repeat`` times (generate.py's num_samples)."""

    def __init__(self, data: Mapping[str, Any], repeat: int):
        super().__init__()
        self.data, self.repeat = data, repeat

    def __len__(self):
        return self.repeat

    def __getitem__(self, index: int) -> Mapping[str, Any]:
        return self.data


class InferenceDataset(Dataset):
    """reference data.py:157-170: a list of featurised complexes (scripts/predict_batch_*.py)."""

    def __init__(self, data: Sequence[Mapping[str, Any]], repeat: int = 0):
        super().__init__()
        self.data = data

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index: int) -> Mapping[str, Any]:
        return self.data[index]


def _to_device(batch: Mapping[str, Any], device) -> dict:
    return {k: (v.to(device, non_blocking=True) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def predict(model, dataloaders: Iterable[Mapping[str, Any]], gather: bool = True) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    """``Trainer.predict(model, dataloaders=...)`` without Lightning: a list of ``predict_step`` results
    ``(pos [B, N, 3] in Angstrom, logits [B, N, 21])`` in dataloader order."""
    device = next(model.parameters()).device
    if device.type != "cuda":
        raise RuntimeError("predict: the model must live on a CUDA sm_100 device (no CPU fallback); call model.cuda() first")
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    was_training = model.training
    model.eval()
    mine, total = [], 0
    with torch.no_grad():
        for idx, batch in enumerate(dataloaders):
            total += 1
            if idx % world != rank:
                continue
            pos, logits = model.predict_step(_to_device(batch, device), idx)
            mine.append((idx, pos, logits))
    model.train(was_training)
    if world == 1 or not gather:
        return [(p, l) for _, p, l in mine]
    # batches differ in shape, so the exchange is object-based (a few MB per batch); results return in dataloader order
    parts: List[Any] = [None] * world
    dist.all_gather_object(parts, [(i, p.cpu(), l.cpu()) for i, p, l in mine])
    merged = sorted((item for part in parts for item in part), key=lambda x: x[0])
    assert len(merged) == total
    return [(p.to(device), l.to(device)) for _, p, l in merged]
