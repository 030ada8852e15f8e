"""Packed-weight cache: fp16 / concatenated copies of module parameters for the CUDA kernels.

The reference overwrites parameters in place between calls (EMA weight swap around
``predict_step``, model.py:249-252), so packed copies are keyed on every source tensor's
``data_ptr``, its in-place ``_version`` and a process-wide weights epoch (bumped around every EMA
swap, because writes through ``param.data`` leave ``_version`` untouched) and rebuilt when any changes.
"""
from __future__ import annotations

import contextlib
from typing import Callable, Sequence

import torch


def half(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float16).contiguous()


def f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def tensor_version(t: torch.Tensor) -> int:
    """In-place modification counter; inference tensors do not track one."""
    try:
        return t._version
    except RuntimeError:
        return -1


# Process-wide "weights epoch".  ``Parameter._version`` does not see writes that go through ``param.data`` (its own
# version counter) -- which is exactly how torch_ema's ``average_parameters()`` swaps the EMA weights in and out
# (``copy_to`` / ``restore`` use ``param.data.copy_``).  Every code path of this package that lets such a writer run
# (the EMA contexts of predict_step / validation_step) bumps the epoch on entry and exit; callers that overwrite
# ``param.data`` themselves call :func:`invalidate_packed_weights`.  The epoch is part of every PackCache key.
_weights_epoch = 0


def weights_epoch() -> int:
    return _weights_epoch


def invalidate_packed_weights() -> None:
    """Force every packed fp16 weight copy (and everything derived from them) to be rebuilt on next use."""
    global _weights_epoch
    _weights_epoch += 1


# "Repack in place" mode (autograd.TrainStepGraph): every PackCache.get() rebuilds its packed copies and writes them INTO
# the tensors it already handed out, so that (a) a CUDA graph that captured those addresses reads the current weights on
# every replay and (b) the packing itself is part of the captured graph.
_repack_in_place = False


@contextlib.contextmanager
def repack_in_place():
    global _repack_in_place
    old, _repack_in_place = _repack_in_place, True
    try:
        yield
    finally:
        _repack_in_place = old


def _copy_into(old, new):
    if isinstance(old, torch.Tensor):
        old.copy_(new)
    elif isinstance(old, dict):
        for k in old:
            _copy_into(old[k], new[k])
    elif isinstance(old, (list, tuple)):
        for a, b in zip(old, new):
            _copy_into(a, b)


class PackCache:
    def __init__(self):
        self._key = None
        self._value = None

    def get(self, sources: Sequence[torch.Tensor], build: Callable[[], object]):
        key = tuple((t.data_ptr(), tensor_version(t), t.device) for t in sources) + (_weights_epoch,)
        if _repack_in_place and self._value is not None:
            with torch.no_grad():
                _copy_into(self._value, build())
            self._key = key
            return self._value
        if key != self._key:
            for t in sources:
                if not t.is_cuda:
                    raise RuntimeError(
                        f"parameter on {t.device}: the B200 denoiser runs on CUDA sm_100 only (no CPU fallback); "
                        "move the module with .cuda() first")
            with torch.no_grad():
                self._value = build()
            self._key = key
        return self._value


def split_rows(w: torch.Tensor) -> torch.Tensor:
    """fp16 pair (hi, lo) of an fp32 weight [rows, K], stacked along rows -> [2*rows, K].

    w ~= hi + lo to ~22 bits: the kernels run the activation against both halves, which removes
    the *systematic* error a single fp16 rounding of the weights would add to every row
    (tools/precision_study.py)."""
    w = w.detach().to(torch.float32)
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    return torch.cat([hi, lo], dim=0).contiguous()


def split_k(w: torch.Tensor) -> torch.Tensor:
    """Same pair stored along K: [rows, 2K] = [hi | lo] (layout of the GEMM's split operand)."""
    w = w.detach().to(torch.float32)
    hi = w.to(torch.float16)
    lo = (w - hi.to(torch.float32)).to(torch.float16)
    return torch.cat([hi, lo], dim=1).contiguous()
