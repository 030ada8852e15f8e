"""Output post-processing on the GPU: what generate.py does with ``trainer.predict``'s results.

* :func:`decode_tokens` / :func:`predict_seq` / :func:`trimmed_sequence` -- reference generate.py:76-91 (argmax of the
  softmax of the sampled logits, 'X' for index 0, leading / trailing 'X' stripped);
* :func:`superpose` -- the role of the TMalign subprocess in generate.py:176-195 (ProteinReDiff/tmalign.py:23-49): rigid
  superposition of every sample (and of its mirror image) onto a reference, the better of the two by TM-score, returned in
  the reference's convention ``aligned = t + pos @ R``.  The correspondence is the residue identity (samples of one
  protein share its residues), so the score is TM-align's TM2 under the identity alignment: a lower bound of what its
  alignment search reports.  Parity: pinned against a float64 SVD Kabsch in the tests (the reference binary is absent).

Both run on hand-written kernels (csrc/prd_post.cu) through the C ABI.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib
from .ops import make_dims
from .models.AF2_modules import _MiniCfg

RESIDUE_TYPES = ["A", "R", "N", "D", "C", "Q", "E", "G", "H", "I", "L", "K", "M", "F", "P", "S", "T", "W", "Y", "V"]  # protein.py:28-31
RESIDUE_TYPES_NEW = ["X"] + RESIDUE_TYPES  # generate.py:80,88

_CFG = _MiniCfg(512, 64, 4)


def decode_tokens(logits: torch.Tensor, residue_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """tokens [B, N] int64 = argmax softmax(logits) (generate.py:79,87); 0 where ``residue_mask`` is 0."""
    B, N, K = logits.shape
    if K != 21:
        raise ValueError(f"expected 21 residue classes, got {K}")
    logits = _lib.check_tensor(logits.contiguous(), torch.float32, "logits")
    mask = None if residue_mask is None else _lib.check_tensor(residue_mask.contiguous(), torch.float32, "residue_mask")
    tokens = torch.empty(B, N, dtype=torch.int64, device=logits.device)
    _lib.call("decode_argmax", make_dims(_CFG, B, N), [logits, mask], [tokens], [])
    return tokens


def predict_seq(logits: torch.Tensor) -> List[List[str]]:
    """generate.py:76-81: one residue letter per token (no trimming)."""
    return [[RESIDUE_TYPES_NEW[i] for i in row] for row in decode_tokens(logits).cpu().tolist()]


def trimmed_sequence(logits: torch.Tensor, residue_mask: Optional[torch.Tensor] = None) -> List[str]:
    """generate.py:86-89: the decoded sequence of every sample with leading / trailing 'X' removed."""
    return ["".join(RESIDUE_TYPES_NEW[i] for i in row).lstrip("X").rstrip("X")
            for row in decode_tokens(logits, residue_mask).cpu().tolist()]


def superpose(pos: torch.Tensor, ref: torch.Tensor, mask: torch.Tensor, mirror: bool = True
              ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Superpose ``pos`` [B, N, 3] onto ``ref`` [N, 3] / [1, N, 3] / [B, N, 3] over the tokens with ``mask`` [B, N] > 0.5.
    Returns ``(tm [B], rmsd [B], t [B, 3], R [B, 3, 3], mirrored [B] bool)`` of the better of (sample, mirror image) by
    TM-score -- generate.py:179-183 -- with ``aligned = t + pos @ R``."""
    B, N, _ = pos.shape
    pos = _lib.check_tensor(pos.contiguous(), torch.float32, "pos")
    ref = ref if ref.dim() == 3 else ref.unsqueeze(0)
    ref = _lib.check_tensor(ref.contiguous(), torch.float32, "ref")
    mask = _lib.check_tensor(mask.contiguous(), torch.float32, "mask")
    if ref.shape[0] not in (1, B) or ref.shape[1:] != pos.shape[1:]:
        raise ValueError(f"reference of shape {tuple(ref.shape)} does not match samples {tuple(pos.shape)}")
    dev = pos.device
    tm = torch.empty(B, 2, device=dev)
    rmsd = torch.empty(B, 2, device=dev)
    R = torch.empty(B, 2, 3, 3, device=dev)
    t = torch.empty(B, 2, 3, device=dev)
    _lib.call("kabsch", make_dims(_CFG, B, N, mode=ref.shape[0]), [pos, ref, mask], [tm, rmsd, R, t], [])
    pick = (tm[:, 1] > tm[:, 0]) if mirror else torch.zeros(B, dtype=torch.bool, device=dev)
    idx = pick.long()
    rows = torch.arange(B, device=dev)
    return tm[rows, idx], rmsd[rows, idx], t[rows, idx], R[rows, idx], pick
