"""Python-side launchers for the module-level C ops (thin: argument checks + pointer packing).

Every function enqueues hand-written sm_100a kernels on the current CUDA stream through
``_lib.call``; none of them computes anything in PyTorch.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import PrdDims

F32, F16, I64 = torch.float32, torch.float16, torch.int64


def make_dims(cfg, B: int, N: int, mode: int = 0, residual: int = 1) -> PrdDims:
    d = PrdDims()
    d.B, d.N = int(B), int(N)
    d.c_s, d.c_z, d.H, d.c, d.tf = cfg.single_dim, cfg.pair_dim, cfg.num_heads, cfg.head_dim, cfg.transition_factor
    d.esm_dim, d.time_dim, d.dist_dim = cfg.esm_dim, cfg.time_dim, cfg.dist_dim
    d.max_bond_distance, d.max_relpos, d.num_steps = cfg.max_bond_distance, cfg.max_relpos, cfg.num_steps
    d.mode, d.residual = int(mode), int(residual)
    return d


def reserve_workspace(cfg, B: int, N: int, device):
    """Size the workspace of the CURRENT stream of ``device`` for every op at (B, N) -- call on the capture stream before
    a CUDA-graph capture.  Returns the buffer: keep it alive for as long as a graph captured against it is replayed."""
    d = make_dims(cfg, B, N)
    need = max(_lib.workspace_bytes(op, d) for op in _lib.OPS)
    return _lib.Workspace.reserve(torch.device(device), need)


def _chk(ts: Sequence[Optional[torch.Tensor]], dtypes, names):
    for t, dt, n in zip(ts, dtypes, names):
        if t is not None:
            _lib.check_tensor(t, dt, n)


def esm_embed(cfg, residue_esm, w_esm_h, out=None):
    B, N, _ = residue_esm.shape
    _chk([residue_esm, w_esm_h], [F32, F16], ["residue_esm", "w_esm_h"])
    out = torch.empty(B, N, cfg.single_dim, dtype=F32, device=residue_esm.device) if out is None else out
    _lib.call("esm_embed", make_dims(cfg, B, N), [residue_esm], [out], [w_esm_h])
    return out


def single_embed(cfg, atom_feats, atom_mask, residue_mask, seq_t, esm_emb, atom_tables, w_type, out=None):
    B, N = atom_mask.shape
    _chk([atom_feats, atom_mask, residue_mask, seq_t, esm_emb, w_type], [I64, F32, F32, F32, F32, F32],
         ["atom_feats", "atom_mask", "residue_mask", "seq_t", "esm_emb", "w_type"])
    out = torch.empty(B, N, cfg.single_dim, dtype=F32, device=atom_mask.device) if out is None else out
    _lib.call("single_embed", make_dims(cfg, B, N), [atom_feats, atom_mask, residue_mask, seq_t, esm_emb], [out],
              list(atom_tables) + [w_type])
    return out


def pair_embed_static(cfg, batch, bond_tables, bdist_table, relpos_table, out=None):
    am = batch["atom_mask"]
    B, N = am.shape
    ins = [am, batch["residue_mask"], batch["bond_mask"], batch["bond_feats"], batch["bond_distance"],
           batch["residue_index"], batch["residue_chain_index"]]
    _chk(ins, [F32, F32, F32, I64, I64, I64, I64],
         ["atom_mask", "residue_mask", "bond_mask", "bond_feats", "bond_distance", "residue_index", "residue_chain_index"])
    out = torch.empty(B, N, N, cfg.pair_dim, dtype=F32, device=am.device) if out is None else out
    _lib.call("pair_embed_static", make_dims(cfg, B, N), ins, [out], list(bond_tables) + [bdist_table, relpos_table])
    return out


def opm_project(cfg, single, mask, weights, out_a=None, out_b=None):
    B, N, _ = single.shape
    _chk([single, mask], [F32, F32], ["single", "mask"])
    od = cfg.single_dim // 4
    out_a = torch.empty(B, N, od, dtype=F32, device=single.device) if out_a is None else out_a
    out_b = torch.empty(B, N, od, dtype=F32, device=single.device) if out_b is None else out_b
    _lib.call("opm_project", make_dims(cfg, B, N), [single, mask], [out_a, out_b], weights)
    return out_a, out_b


def pair_embed(cfg, pair_static, z, mask, t, opm_a, opm_b, weights, out, sampler_state=None, flags: int = 0, rbf_lut=None):
    """``weights``: the six tensors of include/prd_denoiser.h; ``rbf_lut`` (from :func:`rbf_lut_build`) switches the distance
    embedding from the per-pair RBF GEMM to the interpolated table."""
    B, N = mask.shape
    _chk([pair_static, z, mask, t, opm_a, opm_b, out, rbf_lut], [F32, F32, F32, I64, F32, F32, F32, F32],
         ["pair_static", "z", "mask", "t", "opm_a", "opm_b", "pair", "rbf_lut"])
    weights = list(weights)[:6]
    weights += [None] * (6 - len(weights)) + [rbf_lut]  # the C side always reads seven entries
    _lib.call("pair_embed", make_dims(cfg, B, N, mode=flags), [pair_static, z, mask, t, opm_a, opm_b, sampler_state],
              [out], weights)
    return out


def rbf_lut_build(cfg, w_dist, centers, d_max=None):
    """Tabulate d -> W_dist rbf(d) (reference modules.py:73-82 followed by embed_dist's Linear, model.py:360) on the device:
    fp32 weights, PRD_RBF_LUT_POINTS + 1 rows over [0, max center + 0.52 nm]."""
    import ctypes
    _chk([w_dist, centers], [F32, F32], ["w_dist", "centers"])
    lib = _lib.load()
    d = make_dims(cfg, 1, 1)
    lib.prd_rbf_lut_floats.restype = ctypes.c_size_t
    lib.prd_rbf_lut_floats.argtypes = [ctypes.POINTER(PrdDims)]
    lib.prd_rbf_lut_build.restype = ctypes.c_int
    lib.prd_rbf_lut_build.argtypes = [ctypes.POINTER(PrdDims), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float,
                                      ctypes.c_void_p, ctypes.c_void_p]
    lut = torch.empty(int(lib.prd_rbf_lut_floats(ctypes.byref(d))), dtype=F32, device=w_dist.device)
    if d_max is None:
        d_max = float(centers.max()) + 0.52  # host read: callers that rebuild under CUDA-graph capture pass d_max
    rc = lib.prd_rbf_lut_build(ctypes.byref(d), w_dist.data_ptr(), centers.data_ptr(), ctypes.c_float(d_max), lut.data_ptr(),
                               ctypes.c_void_p(torch.cuda.current_stream(w_dist.device).cuda_stream))
    if rc != 0:
        raise RuntimeError(f"prd_rbf_lut_build failed: {_lib.last_error()}")
    return lut


def pair_bias(cfg, pair, proj_a, proj_b=None):
    """The pair-bias projections of the single-representation attentions in ONE pass over the pair tensor.  ``proj_*`` =
    (ln_weight or None, ln_bias or None, w [H, c_z], bias [H] or None), fp32.  Returns (bias_a, bias_b or None), each
    [B, H, N, N]."""
    B, N = pair.shape[:2]
    _chk([pair], [F32], ["pair"])
    H = proj_a[2].shape[0]
    bias_a = torch.empty(B, H, N, N, dtype=F32, device=pair.device)
    bias_b = torch.empty(B, H, N, N, dtype=F32, device=pair.device) if proj_b is not None else None
    weights = list(proj_a) + (list(proj_b) if proj_b is not None else [None] * 4)
    _lib.call("pair_bias", make_dims(cfg, B, N), [pair], [bias_a, bias_b], weights)
    return bias_a, bias_b


def spattention(cfg, single, pair, weights, out=None, bias=None):
    """``bias``: the [B, H, N, N] pair bias precomputed by :func:`pair_bias` (None: projected inside the op)."""
    B, N, _ = single.shape
    _chk([single, pair, bias], [F32, F32, F32], ["single", "pair", "bias"])
    out = torch.empty_like(single) if out is None else out
    _lib.call("spattention", make_dims(cfg, B, N), [single, pair, bias], [out], weights)
    return out


def single_attention(cfg, single, pair, mask, weights, out, residual=1, attn_bias=None):
    B, N, _ = single.shape
    _chk([single, pair, mask, attn_bias, out], [F32, F32, F32, F32, F32], ["single", "pair", "mask", "attn_bias", "out"])
    _lib.call("single_attention", make_dims(cfg, B, N, residual=residual), [single, pair, mask, attn_bias], [out], weights)
    return out


def single_transition(cfg, single, weights, out, residual=1):
    B, N, _ = single.shape
    _chk([single, out], [F32, F32], ["single", "out"])
    _lib.call("single_transition", make_dims(cfg, B, N, residual=residual), [single], [out], weights)
    return out


def outer_linear(cfg, single, pair, weights, out, residual=1):
    B, N, _ = single.shape
    _chk([single, pair, out], [F32, F32, F32], ["single", "pair", "out"])
    _lib.call("outer_linear", make_dims(cfg, B, N, residual=residual), [single, pair], [out], weights)
    return out


def triangle_multiplication(cfg, pair, mask, mode, weights, out, residual=1):
    B, N = mask.shape
    _chk([pair, mask, out], [F32, F32, F32], ["pair", "mask", "out"])
    _lib.call("triangle_multiplication", make_dims(cfg, B, N, mode=mode, residual=residual), [pair, mask], [out], weights)
    return out


def triangle_attention(cfg, pair, mask, mode, weights, out, residual=1, all_valid: bool = False):
    """``all_valid``: the caller's promise that ``mask`` is all ones (bit 1 of PrdDims.mode): a pure performance hint, it
    selects the attention core without the per-sequence handling of ragged batches."""
    B, N = mask.shape
    _chk([pair, mask, out], [F32, F32, F32], ["pair", "mask", "out"])
    _lib.call("triangle_attention", make_dims(cfg, B, N, mode=(mode & 1) | (2 if all_valid else 0), residual=residual),
              [pair, mask], [out], weights)
    return out


def pair_transition(cfg, pair, weights, out, residual=1, next_bias=None):
    """``next_bias`` = (w [H, c_z], b [H]) of the NEXT FoldingBlock's attn_bias: its [B, H, N, N] bias of the updated pair is
    emitted from the same kernel and returned as the second value (None otherwise)."""
    B, N = pair.shape[:2]
    _chk([pair, out], [F32, F32], ["pair", "out"])
    bias_out, extra = None, [None, None]
    if next_bias is not None:
        _chk(list(next_bias), [F32, F32], ["w_bias", "b_bias"])
        bias_out = torch.empty(B, next_bias[0].shape[0], N, N, dtype=F32, device=pair.device)
        extra = list(next_bias)
    _lib.call("pair_transition", make_dims(cfg, B, N, residual=residual), [pair], [out, bias_out], list(weights)[:4] + extra)
    return out if next_bias is None else (out, bias_out)


def symmetrize(cfg, pair):
    B, N = pair.shape[:2]
    _chk([pair], [F32], ["pair"])
    _lib.call("symmetrize", make_dims(cfg, B, N), [], [pair], [])
    return pair


def coord_head(cfg, pair, z, mask, weights, out=None):
    B, N = mask.shape
    _chk([pair, z, mask], [F32, F32, F32], ["pair", "z", "mask"])
    out = torch.empty(B, N, 3, dtype=F32, device=pair.device) if out is None else out
    _lib.call("coord_head", make_dims(cfg, B, N), [pair, z, mask], [out], weights)
    return out


def seq_head(cfg, single, weights, out=None):
    B, N, _ = single.shape
    _chk([single], [F32], ["single"])
    out = torch.empty(B, N, 21, dtype=F32, device=single.device) if out is None else out
    _lib.call("seq_head", make_dims(cfg, B, N), [single], [out], weights)
    return out


def remove_mean(cfg, x, mask):
    """In place; x [R, N, C] with R a multiple of mask.shape[0] (mask rows repeat cyclically)."""
    R, N, C = x.shape
    _chk([x, mask], [F32, F32], ["x", "mask"])
    d = make_dims(cfg, R, N, mode=C)
    d.H = mask.shape[0]
    _lib.call("remove_mean", d, [mask], [x], [])
    return x


def sampler_update(cfg, noise_pred, seq_pred, noise, coef, z, seq_t, state):
    B, N, _ = z.shape
    _chk([noise_pred, seq_pred, noise, coef, z, seq_t], [F32] * 6, ["noise_pred", "seq_pred", "noise", "coef", "z", "seq_t"])
    _lib.check_tensor(state, torch.int32, "sampler_state")
    _lib.call("sampler_update", make_dims(cfg, B, N), [noise_pred, seq_pred, noise, coef], [z, seq_t, state], [])


def diffusion_q(cfg, x, seq, t, noise_z, noise_seq, keep, drop, sched):
    """Forward noising (reference model.py:471-488).  sched: [T, 2] = {sqrt_alphas_cumprod, sqrt_one_minus_alphas_cumprod}."""
    B, N, _ = x.shape
    _chk([x, seq, t, noise_z, noise_seq, keep, drop, sched], [F32, F32, I64, F32, F32, F32, F32, F32],
         ["x", "seq", "t", "noise_z", "noise_seq", "residue_extra_mask", "residue_inv_extra_mask", "sched"])
    z_t, seq_t, seq_t1 = torch.empty_like(x), torch.empty_like(seq), torch.empty_like(seq)
    d = make_dims(cfg, B, N)
    d.num_steps = sched.shape[0]
    _lib.call("diffusion_q", d, [x, seq, t, noise_z, noise_seq, keep, drop, sched], [z_t, seq_t, seq_t1], [])
    return z_t, seq_t, seq_t1


def diffusion_loss(cfg, noise_pred, seq_pred, noise_z, noise_seq, seq_t1, mask, residue_mask, residue_type, t, sched,
                   want_grads: bool = False):
    """Loss terms after the network call + loss = mean(diff_loss / num_nodes) (reference model.py:499-526, 538-541).
    Returns (loss [1], diff_loss [B], terms [B+2] = per-row MSE | KL | CE, d_noise_pred, d_seq_pred); the gradients
    are None unless ``want_grads``."""
    B, N, _ = noise_pred.shape
    _chk([noise_pred, seq_pred, noise_z, noise_seq, seq_t1, mask, residue_mask, residue_type, t, sched],
         [F32, F32, F32, F32, F32, F32, F32, I64, I64, F32],
         ["noise_pred", "seq_pred", "noise_z", "noise_seq", "seq_t1", "mask", "residue_mask", "residue_type", "t", "sched"])
    dev = noise_pred.device
    loss = torch.empty(1, dtype=F32, device=dev)
    diff = torch.empty(B, dtype=F32, device=dev)
    terms = torch.empty(B + 2, dtype=F32, device=dev)
    d_noise = torch.empty_like(noise_pred) if want_grads else None
    d_seq = torch.empty_like(seq_pred) if want_grads else None
    d = make_dims(cfg, B, N)
    d.num_steps = sched.shape[0]
    _lib.call("diffusion_loss", d, [noise_pred, seq_pred, noise_z, noise_seq, seq_t1, mask, residue_mask, residue_type, t, sched],
              [loss, diff, terms, d_noise, d_seq], [])
    return loss, diff, terms, d_noise, d_seq
