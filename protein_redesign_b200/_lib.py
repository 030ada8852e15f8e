"""ctypes binding of libprd_sm100.so (the C ABI declared in include/prd_denoiser.h).

There is no CPU fallback: if the shared library is missing, or a tensor is not on an sm_100
device, the call raises.  Build the library with ``python -c "import __graft_entry__ as g; g.build()"``
or ``protein_redesign_b200/csrc/build.sh``.
"""
from __future__ import annotations

import ctypes
import os
import threading
from typing import Dict, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libprd_sm100.so")
# A/B of two builds on one box (tools/ab_bench.sh PRD_LIB_PATH ...): needs PRD_ALLOW_STALE_LIB=1 next to it
LIB_PATH = os.environ.get("PRD_LIB_PATH") or LIB_PATH

OPS = (
    "esm_embed", "single_embed", "pair_embed_static", "opm_project", "pair_embed", "pair_bias", "spattention",
    "single_attention", "single_transition", "outer_linear", "triangle_multiplication",
    "triangle_attention", "pair_transition", "symmetrize", "coord_head", "seq_head", "remove_mean",
    "sampler_update", "diffusion_q", "diffusion_loss", "decode_argmax", "kabsch",
)


OPS_BWD = (
    "pair_transition", "single_transition", "seq_head", "coord_head", "triangle_attention", "triangle_multiplication",
    "outer_linear", "single_attention", "spattention", "opm_project", "pair_embed", "single_embed",
)


class PrdDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        "B", "N", "c_s", "c_z", "H", "c", "tf", "esm_dim", "time_dim", "dist_dim",
        "max_bond_distance", "max_relpos", "num_steps", "mode", "residual")]


class PrdGemm(ctypes.Structure):
    _fields_ = [
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32), ("nb1", ctypes.c_int32),
        ("nb2", ctypes.c_int32),
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64), ("a_bs1", ctypes.c_int64), ("a_bs2", ctypes.c_int64),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int64), ("b_bs1", ctypes.c_int64), ("b_bs2", ctypes.c_int64),
        ("alpha", ctypes.c_float), ("act", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("rowscale", ctypes.c_void_p), ("rs_bs1", ctypes.c_int64), ("rs_bs2", ctypes.c_int64),
        ("mul", ctypes.c_void_p), ("ldmul", ctypes.c_int64), ("mul_bs1", ctypes.c_int64), ("mul_bs2", ctypes.c_int64),
        ("add", ctypes.c_void_p), ("ldadd", ctypes.c_int64), ("add_bs1", ctypes.c_int64), ("add_bs2", ctypes.c_int64),
        ("C", ctypes.c_void_p), ("ldc", ctypes.c_int64), ("c_bs1", ctypes.c_int64), ("c_bs2", ctypes.c_int64),
        ("c_fp16", ctypes.c_int32), ("tf32", ctypes.c_int32), ("mul_step", ctypes.c_int32), ("round_tf32", ctypes.c_int32),
    ]


_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()


def source_hash() -> Optional[str]:
    """Hash of the CUDA sources beside this file, computed exactly like csrc/build.sh does (None if they are absent)."""
    import glob
    import hashlib

    csrc = os.path.join(_HERE, "csrc")
    inc = os.path.join(_HERE, "..", "include", "prd_denoiser.h")
    files = sorted(glob.glob(os.path.join(csrc, "*.cu"))) + sorted(glob.glob(os.path.join(csrc, "*.cuh"))) + \
        sorted(glob.glob(os.path.join(csrc, "*.h"))) + [inc]
    if not os.path.exists(inc) or not glob.glob(os.path.join(csrc, "*.cu")):
        return None
    h = hashlib.sha256()
    for f in files:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def load() -> ctypes.CDLL:
    """Load the shared library (once) and declare every prototype of include/prd_denoiser.h."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no CPU fallback. "
                "Run protein_redesign_b200/csrc/build.sh (needs nvcc with sm_100a support).")
        lib = ctypes.CDLL(LIB_PATH)
        lib.prd_source_hash.restype = ctypes.c_char_p
        built, have = lib.prd_source_hash().decode(), source_hash()
        if have is not None and built != have and os.environ.get("PRD_ALLOW_STALE_LIB") != "1":
            raise RuntimeError(
                f"{LIB_PATH} was built from other sources (library {built}, tree {have}): rebuild with "
                "protein_redesign_b200/csrc/build.sh -- a stale kernel must never be tested or benchmarked")
        lib.prd_version.restype = ctypes.c_int
        lib.prd_launch_count.restype = ctypes.c_longlong
        lib.prd_last_error.restype = ctypes.c_char_p
        lib.prd_device_check.restype = ctypes.c_int
        lib.prd_gemm_f16.restype = ctypes.c_int
        lib.prd_gemm_f16.argtypes = [ctypes.POINTER(PrdGemm), ctypes.c_void_p]
        vpp = ctypes.POINTER(ctypes.c_void_p)
        for op in OPS:
            f = getattr(lib, f"prd_{op}_fwd")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(PrdDims), vpp, vpp, vpp, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
            w = getattr(lib, f"prd_{op}_workspace_bytes")
            w.restype = ctypes.c_size_t
            w.argtypes = [ctypes.POINTER(PrdDims)]
        for op in OPS_BWD:
            f = getattr(lib, f"prd_{op}_bwd")
            f.restype = ctypes.c_int
            f.argtypes = [ctypes.POINTER(PrdDims), vpp, vpp, vpp, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
            w = getattr(lib, f"prd_{op}_bwd_workspace_bytes")
            w.restype = ctypes.c_size_t
            w.argtypes = [ctypes.POINTER(PrdDims)]
        _lib = lib
        return lib


def last_error() -> str:
    return load().prd_last_error().decode("utf-8", "replace")


def _ptr_array(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (ctypes.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def check_tensor(t: torch.Tensor, dtype: torch.dtype, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a tensor, got {type(t)}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tensor is on {t.device}; libprd_sm100 runs on CUDA sm_100 only (no CPU fallback)")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: tensor must be contiguous")
    return t


# One grow-only workspace per (device, stream): ops enqueued on different streams of one device never share scratch
# memory.  It must be sized before a CUDA-graph capture starts (Workspace.reserve on the capture stream); growing it
# during capture would allocate.  Growing replaces the buffer object: whoever replays a captured graph keeps the
# tensor returned by reserve() / current() alive next to the graph (ProteinReDiffModel.sample, bench.py do).
class Workspace:
    _buffers: Dict[tuple, torch.Tensor] = {}
    _guard = threading.Lock()

    @staticmethod
    def _key(device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        return idx, int(torch.cuda.current_stream(device).cuda_stream)

    @classmethod
    def reserve(cls, device: torch.device, nbytes: int) -> torch.Tensor:
        key = cls._key(device)
        with cls._guard:
            buf = cls._buffers.get(key)
            if buf is None or buf.numel() < nbytes:
                if torch.cuda.is_current_stream_capturing():
                    raise RuntimeError("prd workspace must be reserved (on the capture stream) before CUDA-graph capture")
                buf = torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)
                cls._buffers[key] = buf
            return buf

    @classmethod
    def current(cls, device: torch.device) -> Optional[torch.Tensor]:
        """The buffer the ops enqueued on the current stream of ``device`` use right now (None before the first call)."""
        return cls._buffers.get(cls._key(device))


def workspace_bytes(op: str, dims: PrdDims) -> int:
    return int(getattr(load(), f"prd_{op}_workspace_bytes")(ctypes.byref(dims)))


def call(op: str, dims: PrdDims, ins: Sequence[Optional[torch.Tensor]], outs: Sequence[Optional[torch.Tensor]],
         weights: Sequence[Optional[torch.Tensor]]) -> None:
    """Enqueue prd_<op>_fwd on the current torch CUDA stream."""
    lib = load()
    dev = None
    for t in list(ins) + list(outs) + list(weights):
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError(f"prd_{op}: tensor on {t.device}; CUDA sm_100 only (no CPU fallback)")
            dev = t.device
            break
    if dev is None:
        raise RuntimeError(f"prd_{op}: no tensors given")
    need = workspace_bytes(op, dims)
    ws = Workspace.reserve(dev, need)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = getattr(lib, f"prd_{op}_fwd")(ctypes.byref(dims), _ptr_array(ins), _ptr_array(outs), _ptr_array(weights),
                                       ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()),
                                       ctypes.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"prd_{op}_fwd failed: {last_error()}")


def call_bwd(op: str, dims: PrdDims, ins: Sequence[Optional[torch.Tensor]], outs: Sequence[Optional[torch.Tensor]],
             weights: Sequence[Optional[torch.Tensor]]) -> None:
    """Enqueue prd_<op>_bwd on the current torch CUDA stream (same conventions as :func:`call`)."""
    lib = load()
    dev = None
    for t in list(ins) + list(outs) + list(weights):
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError(f"prd_{op}_bwd: tensor on {t.device}; CUDA sm_100 only (no CPU fallback)")
            if not t.is_contiguous():
                raise ValueError(f"prd_{op}_bwd: every tensor must be contiguous")
            dev = t.device
    if dev is None:
        raise RuntimeError(f"prd_{op}_bwd: no tensors given")
    need = int(getattr(lib, f"prd_{op}_bwd_workspace_bytes")(ctypes.byref(dims)))
    ws = Workspace.reserve(dev, need)
    stream = torch.cuda.current_stream(dev).cuda_stream
    rc = getattr(lib, f"prd_{op}_bwd")(ctypes.byref(dims), _ptr_array(ins), _ptr_array(outs), _ptr_array(weights),
                                       ctypes.c_void_p(ws.data_ptr()), ctypes.c_size_t(ws.numel()),
                                       ctypes.c_void_p(stream))
    if rc != 0:
        raise RuntimeError(f"prd_{op}_bwd failed: {last_error()}")


def gemm_f16(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor, *, alpha: float = 1.0, bias=None, act: int = 0,
             rowscale=None, mul=None, add=None, mul_step: bool = False, round_tf32: bool = False) -> torch.Tensor:
    """out[..., M, N] = epilogue(alpha * a[..., M, K] @ b[..., N, K]^T); a, b fp16 -- or both fp32, multiplied on
    kind::tf32 (the backward pass's operand type); up to two batch dims (b may omit them = shared weight).  Test hook
    for the tcgen05 + TMA machinery."""
    lib = load()
    tf32 = a.dtype == torch.float32
    check_tensor(a, torch.float32 if tf32 else torch.float16, "a")
    check_tensor(b, torch.float32 if tf32 else torch.float16, "b")
    M, K = a.shape[-2:]
    N = b.shape[-2]
    batch = list(a.shape[:-2])
    while len(batch) < 2:
        batch.insert(0, 1)
    nb2, nb1 = batch
    g = PrdGemm()
    g.M, g.N, g.K, g.nb1, g.nb2 = M, N, K, nb1, nb2
    g.A, g.lda, g.a_bs1, g.a_bs2 = a.data_ptr(), K, M * K, nb1 * M * K
    g.B, g.ldb = b.data_ptr(), K
    if b.dim() > 2:
        g.b_bs1, g.b_bs2 = N * K, nb1 * N * K
    g.alpha, g.act = alpha, act
    g.bias = None if bias is None else bias.data_ptr()
    if rowscale is not None:
        g.rowscale, g.rs_bs1, g.rs_bs2 = rowscale.data_ptr(), M, nb1 * M
    if mul is not None:
        g.mul, g.ldmul, g.mul_bs1, g.mul_bs2 = mul.data_ptr(), N, M * N, nb1 * M * N
    if add is not None:
        g.add, g.ldadd, g.add_bs1, g.add_bs2 = add.data_ptr(), N, M * N, nb1 * M * N
    g.C, g.ldc, g.c_bs1, g.c_bs2 = out.data_ptr(), N, M * N, nb1 * M * N
    g.c_fp16 = 1 if out.dtype == torch.float16 else 0
    g.tf32, g.mul_step, g.round_tf32 = int(tf32), int(mul_step), int(round_tf32)
    rc = lib.prd_gemm_f16(ctypes.byref(g), ctypes.c_void_p(torch.cuda.current_stream(a.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"prd_gemm_f16 failed: {last_error()}")
    return out
