"""CPU oracle for the ProteinReDiff denoiser hot path -- TEST INFRASTRUCTURE ONLY.

A plain fp32 PyTorch (CPU) restatement of the reference algorithm, written as stateless
functions over a state-dict.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker / timed
CPU baseline -- never from the product package ``protein_redesign_b200``.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks every function below against
outputs of the unmodified reference (imported from /root/reference with import stubs by
``tests/golden/make_golden.py``; fixtures committed under ``tests/golden/``).

Each function cites the reference file:line it restates (paths relative to the reference
repo root).  Notation follows SURVEY.md Appendix A.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]
LN_EPS = 1e-5  # nn.LayerNorm default, used everywhere
MASK_FILL = -(2.0 ** 15)  # ProteinReDiff/modules.py:177,220
# Memory knob for the large parity cases (N = 1024: the eager logits alone are 17 GB): when set, the row-independent
# ops (triangle attention, OuterLinear) are evaluated ROW_CHUNK pair rows at a time.  Same arithmetic per row.
ROW_CHUNK: Optional[int] = None


def _ln(x: Tensor, w: Optional[Tensor] = None, b: Optional[Tensor] = None) -> Tensor:
    return F.layer_norm(x, x.shape[-1:], w, b, LN_EPS)


# --------------------------------------------------------------------------------------
# schedule / helpers
# --------------------------------------------------------------------------------------
def get_betas(num_steps: int, schedule: str) -> Tensor:
    """ProteinReDiff/difffusion.py:8-26."""
    if schedule == "linear":
        return torch.linspace(0.0001, 0.02, num_steps)
    if schedule == "cosine":
        steps = num_steps + 1
        x = torch.linspace(0, num_steps, steps)
        ac = torch.cos((x / steps) * math.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    raise ValueError(f"Invalid schedule: {schedule}")


def schedule_tables(num_steps: int, schedule: str) -> Dict[str, Tensor]:
    """ProteinReDiff/model.py:172-190 (only the tables the sampler / q() read)."""
    betas = get_betas(num_steps, schedule)
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, 0)
    return {
        "betas": betas,
        "alphas": alphas,
        "alphas_cumprod": ac,
        "sqrt_betas": torch.sqrt(betas),
        "sqrt_alphas": torch.sqrt(alphas),
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
    }


def remove_mean(x: Tensor, mask: Tensor) -> Tensor:
    """ProteinReDiff/utils.py:32-36."""
    m = mask.unsqueeze(-1).expand_as(x)
    s = (m * x).sum(dim=1, keepdim=True)
    n = m.sum(dim=1, keepdim=True)
    return x - m * s / n


# --------------------------------------------------------------------------------------
# batch preparation (index / mask path -- bit exact)
# --------------------------------------------------------------------------------------
def random_residue_mask(residue_mask: Tensor, max_p: float) -> Tuple[Tensor, Tensor]:
    """Deterministic-count branch of RandomMaskingModule (mask_utils.py:77-99, stochastic=False).

    Consumes the global torch CPU generator exactly like the reference (one ``randperm``
    over the number of valid residues of the whole batch).  Returns (keep, drop) masks.
    """
    ones = residue_mask == 1
    num_ones = int(ones.sum().item())
    num_drop = int(num_ones * max_p)
    rows, cols = torch.where(ones)
    pick = torch.randperm(num_ones)[:num_drop]
    keep = residue_mask.clone()
    keep[rows[pick], cols[pick]] = 0
    drop = torch.zeros_like(residue_mask)
    drop[rows[pick], cols[pick]] = 1
    return keep, drop


def prepare_batch(batch: Dict[str, Tensor], mask_prob: float) -> Dict[str, Tensor]:
    """Inference branch of ProteinReDiffModel.prepare_batch (model.py:424-468)."""
    out = dict(batch)
    atom_mask, residue_mask = batch["atom_mask"], batch["residue_mask"]
    ca = batch["residue_atom_pos"][:, :, 1]
    one_hot = F.one_hot(batch["residue_type"], num_classes=21) * 2.0 - 1.0  # model.py:433
    pos = atom_mask.unsqueeze(-1) * batch["atom_pos"] + residue_mask.unsqueeze(-1) * ca
    keep, drop = random_residue_mask(residue_mask, mask_prob)
    out["residue_esm"] = batch["residue_esm"] * keep.unsqueeze(-1)
    out["residue_type_masked"] = (batch["residue_type"] * keep).long()
    out["residue_one_hot"] = one_hot * keep.unsqueeze(-1)
    out["residue_extra_mask"] = keep
    out["residue_inv_extra_mask"] = drop
    out["x"] = 0.1 * pos  # utils.py:24-25
    out["residue_and_atom_mask"] = atom_mask + residue_mask
    return out


# --------------------------------------------------------------------------------------
# input embeddings (A1)
# --------------------------------------------------------------------------------------
def embed_single(sd: SD, batch: Dict[str, Tensor], seq_t: Tensor) -> Tensor:
    """ProteinReDiff/model.py:342-346 with modules.py:35-51 and model.py:89-93,99-102."""
    feats = batch["atom_feats"]
    n_feat = feats.shape[-1]
    scale = 1.0 / math.sqrt(n_feat)
    atom = 0.0
    for f in range(n_feat):
        atom = atom + scale * sd[f"embed_atom_feats.embeddings.{f}.weight"][feats[..., f]]
    res_type = F.relu(F.linear(_ln(seq_t), sd["embed_residue_type.1.weight"]))
    res_esm = F.linear(_ln(batch["residue_esm"]), sd["embed_residue_esm.1.weight"])
    return batch["atom_mask"].unsqueeze(-1) * atom + batch["residue_mask"].unsqueeze(-1) * (res_type + res_esm)


def embed_pair_static(sd: SD, batch: Dict[str, Tensor], max_bond_distance: int, max_relpos: int) -> Tensor:
    """Step-invariant part of the pair embedding (model.py:348-358; modules.py:54-70)."""
    am, rm = batch["atom_mask"], batch["residue_mask"]
    am2 = am.unsqueeze(-1) * am.unsqueeze(-2)
    rm2 = rm.unsqueeze(-1) * rm.unsqueeze(-2)
    bf = batch["bond_feats"]
    scale = 1.0 / math.sqrt(bf.shape[-1])
    bond = 0.0
    for f in range(bf.shape[-1]):
        bond = bond + scale * sd[f"embed_bond_feats.embeddings.{f}.weight"][bf[..., f]]
    bdist = sd["embed_bond_distance.weight"][batch["bond_distance"].clamp(max=max_bond_distance)]
    pair = am2.unsqueeze(-1) * (batch["bond_mask"].unsqueeze(-1) * bond + bdist)
    idx, chain = batch["residue_index"], batch["residue_chain_index"]
    rel = idx.unsqueeze(-1) - idx.unsqueeze(-2)
    same = (chain.unsqueeze(-1) == chain.unsqueeze(-2)).float()
    relemb = sd["embed_relpos.weight"][max_relpos + rel.clamp(min=-max_relpos, max=max_relpos)]
    return pair + rm2.unsqueeze(-1) * (same.unsqueeze(-1) * relemb)


def embed_pair_dynamic(sd: SD, z: Tensor, t: Tensor, mask: Tensor, num_steps: int) -> Tensor:
    """Per-step part: RBF of noisy distances + time embedding (model.py:337-341,359-361;
    modules.py:73-82 RadialBasisProjection, :85-97 SinusoidalProjection)."""
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    d = torch.linalg.norm(z.unsqueeze(-2) - z.unsqueeze(-3), dim=-1)
    center = sd["embed_dist.0.center"]
    rbf_scale = (center.numel() - 1) / 2.0
    rbf = torch.exp(-rbf_scale * torch.square(d.unsqueeze(-1) - center))
    dist = F.linear(rbf, sd["embed_dist.1.weight"])
    scaled_t = t / num_steps  # int64 / int -> float32 true division (model.py:341)
    wx = sd["embed_beta.0.weight"] * scaled_t[:, None, None].unsqueeze(-1)
    beta = F.linear(torch.cat([torch.sin(wx), torch.cos(wx)], dim=-1), sd["embed_beta.1.weight"])
    return m2.unsqueeze(-1) * (dist + beta)


# --------------------------------------------------------------------------------------
# Denoiser prologue (A2)
# --------------------------------------------------------------------------------------
def outer_product_update(sd: SD, single: Tensor, mask: Tensor, prefix: str = "Denoiser.opm.") -> Tensor:
    """AF2_modules.OuterProductUpdate.forward (AF2_modules.py:503-545).

    The einsum at :532 has no contracted index: out[b,i,j,c] = a[b,i,c] * b[b,j,c].
    """
    x = _ln(single, sd[prefix + "layer_norm.weight"], sd[prefix + "layer_norm.bias"])
    m = mask.unsqueeze(-1)
    a = F.linear(x, sd[prefix + "linear_1.weight"], sd[prefix + "linear_1.bias"]) * m
    b = F.linear(x, sd[prefix + "linear_2.weight"], sd[prefix + "linear_2.bias"]) * m
    outer = a.unsqueeze(2) * b.unsqueeze(1)
    outer = F.linear(outer, sd[prefix + "linear_out.weight"], sd[prefix + "linear_out.bias"])
    norm = (m.unsqueeze(2) * m.unsqueeze(1)) + 1e-3
    return outer / norm


def single_pair_attention(sd: SD, single: Tensor, pair: Tensor, num_heads: int,
                          prefix: str = "Denoiser.SPAAttnBlock.") -> Tensor:
    """AF2_modules.SPAttention.forward (:421-473) + Attention (:251-367) + _attention (:613-627).

    Head width is c_s per head (modules.py:366-371); no key mask (mask_bias at :447 is dead);
    the residual is taken on the LayerNorm-ed input (:465-470).
    """
    B, N, cs = single.shape
    H = num_heads
    zb = F.linear(_ln(pair, sd[prefix + "linear_z.0.weight"], sd[prefix + "linear_z.0.bias"]),
                  sd[prefix + "linear_z.1.weight"])  # [B,N,N,H]
    bias = zb.permute(0, 3, 1, 2)  # [B,H,N,N]
    x = _ln(single, sd[prefix + "layer_norm_m.weight"], sd[prefix + "layer_norm_m.bias"])
    q = F.linear(x, sd[prefix + "mha.linear_q.weight"]).view(B, N, H, -1).transpose(1, 2)
    k = F.linear(x, sd[prefix + "mha.linear_k.weight"]).view(B, N, H, -1).transpose(1, 2)
    v = F.linear(x, sd[prefix + "mha.linear_v.weight"]).view(B, N, H, -1).transpose(1, 2)
    q = q / math.sqrt(q.shape[-1])
    a = torch.softmax(torch.matmul(q, k.transpose(-1, -2)) + bias, dim=-1)
    o = torch.matmul(a, v).transpose(1, 2)  # [B,N,H,C]
    g = torch.sigmoid(F.linear(x, sd[prefix + "mha.linear_g.weight"], sd[prefix + "mha.linear_g.bias"]))
    o = (o * g.view(B, N, H, -1)).reshape(B, N, -1)
    return x + F.linear(o, sd[prefix + "mha.linear_o.weight"], sd[prefix + "mha.linear_o.bias"])


# --------------------------------------------------------------------------------------
# FoldingBlock pieces (A3)
# --------------------------------------------------------------------------------------
def gated_attention(sd: SD, prefix: str, x: Tensor, mask: Tensor, num_heads: int,
                    attn_bias: Optional[Tensor] = None) -> Tensor:
    """modules.Attention.forward (modules.py:185-225).  x [..., L, D], mask [..., L]."""
    H = num_heads
    x = _ln(x)
    L = x.shape[-2]

    def heads(y):
        return y.view(*y.shape[:-1], H, -1).transpose(-2, -3)  # [..., H, L, c]

    q = heads(F.linear(x, sd[prefix + "q_proj.weight"]))
    k = heads(F.linear(x, sd[prefix + "k_proj.weight"]))
    v = heads(F.linear(x, sd[prefix + "v_proj.weight"]))
    g = heads(torch.sigmoid(F.linear(x, sd[prefix + "gate_proj.weight"], sd[prefix + "gate_proj.bias"])))
    scale = 1.0 / math.sqrt(q.shape[-1])
    logits = torch.matmul(scale * q, k.transpose(-1, -2))
    if attn_bias is not None:
        logits = logits + attn_bias
    key_mask = mask[..., None, None, :]
    logits = logits.masked_fill(key_mask < 0.5, MASK_FILL)
    attn = torch.softmax(logits, dim=-1)
    o = g * torch.matmul(attn, v)
    o = o.transpose(-2, -3).reshape(*x.shape[:-1], -1)
    return F.linear(o, sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def attn_bias_from_pair(sd: SD, prefix: str, pair: Tensor) -> Tensor:
    """FoldingBlock.attn_bias (modules.py:300-304): LN -> c_z->H (+bias) -> [B,H,N,N]."""
    return F.linear(_ln(pair), sd[prefix + "attn_bias.1.weight"], sd[prefix + "attn_bias.1.bias"]).permute(0, 3, 1, 2)


def transition(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """single_fc / pair_fc (modules.py:306-311, 321-326)."""
    h = F.relu(F.linear(_ln(x), sd[prefix + "1.weight"], sd[prefix + "1.bias"]))
    return F.linear(h, sd[prefix + "3.weight"], sd[prefix + "3.bias"])


def outer_linear(sd: SD, prefix: str, single: Tensor) -> Tensor:
    """modules.OuterLinear.forward (modules.py:283-287)."""
    x = _ln(single)
    xj = x.unsqueeze(-3)
    n = x.shape[-2]
    step = ROW_CHUNK or n
    outs = []
    for r0 in range(0, n, step):
        xi = x[..., r0:r0 + step, :].unsqueeze(-2)
        outs.append(F.linear(torch.cat([xi * xj, (xi - xj).expand(*xi.shape[:-3], xi.shape[-3], xj.shape[-2], -1)], dim=-1),
                             sd[prefix + "linear.weight"], sd[prefix + "linear.bias"]))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=-3)


def triangle_multiplication(sd: SD, prefix: str, pair: Tensor, mask_2d: Tensor, mode: str) -> Tensor:
    """modules.TriangleMultiplication.forward (modules.py:262-274)."""
    p = _ln(pair)
    ab = mask_2d.unsqueeze(-1) * torch.sigmoid(F.linear(p, sd[prefix + "ab_gate.weight"], sd[prefix + "ab_gate.bias"])) \
        * F.linear(p, sd[prefix + "ab_proj.weight"], sd[prefix + "ab_proj.bias"])
    a, b = torch.chunk(ab, 2, dim=-1)
    if mode == "outgoing":
        x = torch.einsum("bikd,bjkd->bijd", a, b)
    elif mode == "incoming":
        x = torch.einsum("bkid,bkjd->bijd", a, b)
    else:
        raise ValueError(f"Invalid mode: {mode}")
    gate = torch.sigmoid(F.linear(p, sd[prefix + "out_gate.weight"], sd[prefix + "out_gate.bias"]))
    return gate * F.linear(_ln(x), sd[prefix + "out_proj.weight"], sd[prefix + "out_proj.bias"])


def triangle_attention(sd: SD, prefix: str, pair: Tensor, mask_2d: Tensor, num_heads: int, mode: str) -> Tensor:
    """modules.TriangleAttention.forward (modules.py:236-243): rows ("starting") or columns
    ("ending") of the pair tensor are independent sequences; no triangle bias."""
    if mode == "ending":
        pair, mask_2d = pair.transpose(1, 2), mask_2d.transpose(1, 2)
    elif mode != "starting":
        raise ValueError(f"Invalid mode: {mode}")
    n = pair.shape[1]
    step = ROW_CHUNK or n
    outs = [gated_attention(sd, prefix + "attn.", pair[:, r0:r0 + step], mask_2d[:, r0:r0 + step], num_heads)
            for r0 in range(0, n, step)]
    out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
    return out.transpose(1, 2) if mode == "ending" else out


def folding_block(sd: SD, k: int, single: Tensor, pair: Tensor, mask: Tensor, num_heads: int,
                  probe: Optional[Callable[[str, Tensor], None]] = None) -> Tuple[Tensor, Tensor]:
    """modules.FoldingBlock.forward (modules.py:328-343) -- eight residual updates in order."""
    p = f"Denoiser.folding_blocks.{k}."
    rec = probe or (lambda n, t: None)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    bias = attn_bias_from_pair(sd, p, pair)
    single = single + gated_attention(sd, p + "single_attn.", single, mask, num_heads, bias)
    rec(p + "single_attn", single)
    single = single + transition(sd, p + "single_fc.", single)
    rec(p + "single_fc", single)
    pair = pair + outer_linear(sd, p + "outer_linear.", single)
    rec(p + "outer_linear", pair)
    pair = pair + triangle_multiplication(sd, p + "pair_mul_outgoing.", pair, m2, "outgoing")
    rec(p + "pair_mul_outgoing", pair)
    pair = pair + triangle_multiplication(sd, p + "pair_mul_incoming.", pair, m2, "incoming")
    rec(p + "pair_mul_incoming", pair)
    pair = pair + triangle_attention(sd, p + "pair_attn_starting.", pair, m2, num_heads, "starting")
    rec(p + "pair_attn_starting", pair)
    pair = pair + triangle_attention(sd, p + "pair_attn_ending.", pair, m2, num_heads, "ending")
    rec(p + "pair_attn_ending", pair)
    pair = pair + transition(sd, p + "pair_fc.", pair)
    rec(p + "pair_fc", pair)
    return single, pair


def denoiser_trunk(sd: SD, cfg, single: Tensor, pair: Tensor, mask: Tensor,
                   probe: Optional[Callable[[str, Tensor], None]] = None) -> Tuple[Tensor, Tensor]:
    """modules.Denoiser.forward (modules.py:391-404)."""
    rec = probe or (lambda n, t: None)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    pair = pair + m2.unsqueeze(-1) * outer_product_update(sd, single, mask)
    rec("Denoiser.opm", pair)
    single = single_pair_attention(sd, single, pair, cfg.num_heads)
    rec("Denoiser.SPAAttnBlock", single)
    for k in range(cfg.num_blocks):
        single, pair = folding_block(sd, k, single, pair, mask, cfg.num_heads, probe)
    pair = 0.5 * (pair + pair.transpose(1, 2))
    rec("Denoiser.out_pair", pair)
    return single, pair


# --------------------------------------------------------------------------------------
# heads (A4) and the full step
# --------------------------------------------------------------------------------------
def coord_head(sd: SD, pair: Tensor, z: Tensor, mask: Tensor) -> Tensor:
    """model.py:364-373: radial weights x unit difference vectors, summed over j, mean removed."""
    h = F.relu(F.linear(_ln(pair), sd["weight_radial.1.weight"], sd["weight_radial.1.bias"]))
    w = F.linear(h, sd["weight_radial.3.weight"])  # [B,N,N,1]
    dz = z.unsqueeze(-2) - z.unsqueeze(-3)
    r = dz * torch.rsqrt(torch.sum(torch.square(dz), -1, keepdim=True) + 1e-4)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    eps = (m2.unsqueeze(-1) * w * r).sum(dim=2)
    return remove_mean(eps, mask)


def seq_head(sd: SD, single: Tensor) -> Tensor:
    """model.py:374 with seq_mlp (:117-122)."""
    h = F.relu(F.linear(_ln(single), sd["seq_mlp.1.weight"], sd["seq_mlp.1.bias"]))
    return F.linear(h, sd["seq_mlp.3.weight"])


def denoiser_step(sd: SD, cfg, batch: Dict[str, Tensor], z: Tensor, seq_t: Tensor, mask: Tensor, t: Tensor,
                  probe: Optional[Callable[[str, Tensor], None]] = None) -> Tuple[Tensor, Tensor]:
    """ProteinReDiffModel.sample_step == forward (model.py:318-375 / :254-316)."""
    rec = probe or (lambda n, t_: None)
    single = embed_single(sd, batch, seq_t)
    rec("embed_single", single)
    pair = embed_pair_static(sd, batch, cfg.max_bond_distance, cfg.max_relpos)
    pair = pair + embed_pair_dynamic(sd, z, t, mask, cfg.num_steps)
    rec("embed_pair", pair)
    single, pair = denoiser_trunk(sd, cfg, single, pair, mask, probe)
    noise_pred = coord_head(sd, pair, z, mask)
    seq_pred = seq_head(sd, single)
    return noise_pred, seq_pred


# --------------------------------------------------------------------------------------
# sampler (A5)
# --------------------------------------------------------------------------------------
def sample(sd: SD, cfg, batch: Dict[str, Tensor], randn_like: Callable[[Tensor], Tensor] = torch.randn_like,
           num_steps: Optional[int] = None, trace: Optional[list] = None) -> Tuple[Tensor, Tensor]:
    """ProteinReDiffModel.sample (model.py:377-422).  ``randn_like`` lets tests inject
    pre-generated noise in the reference's draw order: z_T, seq_T, then one draw per
    non-final step."""
    T = cfg.num_steps if num_steps is None else num_steps
    tab = schedule_tables(T, cfg.diffusion_schedule)
    batch = prepare_batch(batch, cfg.mask_prob)
    x, mask = batch["x"], batch["residue_and_atom_mask"]
    residue_mask, seq = batch["residue_mask"], batch["residue_one_hot"]
    keep, drop = batch["residue_extra_mask"], batch["residue_inv_extra_mask"]
    B = x.shape[0]
    time_steps = torch.linspace(T - 1, 0, steps=T).long()
    z = remove_mean(randn_like(x), mask)
    seq_t = remove_mean(randn_like(seq), residue_mask)
    seq_t = keep.unsqueeze(-1) * seq + drop.unsqueeze(-1) * seq_t
    step_cfg = cfg if T == cfg.num_steps else _with_steps(cfg, T)
    seq_pred = None
    for i in range(T):
        t = torch.broadcast_to(time_steps[i], (B,))
        w_noise = (1.0 - tab["alphas"][t]) / tab["sqrt_one_minus_alphas_cumprod"][t]
        noise_pred, seq_pred = denoiser_step(sd, step_cfg, batch, z, seq_t, mask, t)
        mean = (1.0 / tab["sqrt_alphas"][t])[:, None, None] * (z - w_noise[:, None, None] * noise_pred)
        seq_t = torch.softmax(seq_pred, dim=-1) * 2 - 1
        if bool((t == 0).all()):
            z = mean
        else:
            noise = remove_mean(randn_like(x), mask)
            z = mean + tab["sqrt_betas"][t][:, None, None] * noise
        if trace is not None:
            trace.append((z.clone(), seq_pred.clone(), noise_pred.clone()))
    return 10.0 * z, residue_mask.unsqueeze(-1) * seq_pred


def _with_steps(cfg, T):
    import dataclasses

    return dataclasses.replace(cfg, num_steps=T)


# --------------------------------------------------------------------------------------
# training objective
# --------------------------------------------------------------------------------------
def q_sample(tab: Dict[str, Tensor], x: Tensor, seq: Tensor, t: Tensor, noise_z: Tensor, noise_seq: Tensor,
             keep: Tensor, drop: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """ProteinReDiffModel.q (model.py:471-488): forward noising of the coordinates and of the +-1 one-hot
    sequence; known residues (``keep`` = residue_extra_mask) stay clean, ``drop`` rows are noised."""
    sa, s1 = tab["sqrt_alphas_cumprod"], tab["sqrt_one_minus_alphas_cumprod"]
    z_t = sa[t][:, None, None] * x + s1[t][:, None, None] * noise_z
    seq_t = sa[t][:, None, None] * seq + s1[t][:, None, None] * noise_seq
    seq_t = keep.unsqueeze(-1) * seq + drop.unsqueeze(-1) * seq_t
    t1 = (t - 1).clamp(min=0)
    seq_t1 = sa[t1][:, None, None] * seq + s1[t1][:, None, None] * noise_seq
    return z_t, seq_t, seq_t1, t1


def loss_from_outputs(tab: Dict[str, Tensor], batch: Dict[str, Tensor], noise_pred: Tensor, seq_pred: Tensor,
                      noise_z: Tensor, noise_seq: Tensor, seq_t1: Tensor, t1: Tensor) -> Tensor:
    """The three terms of ProteinReDiffModel.diffusion_loss after the network call (model.py:499-526):
    per-row noise MSE (sum) + KL(softmax(seq_{t-1}) || softmax(seq_pred_{t-1})) + CE((seq_pred+1)/2, type;
    ignore_index 0) * mask; the KL and CE terms are scalar sums added to every batch row (:512-525).
    Returns diff_loss [B]."""
    mask, residue_mask = batch["residue_and_atom_mask"], batch["residue_mask"]
    sa, s1 = tab["sqrt_alphas_cumprod"], tab["sqrt_one_minus_alphas_cumprod"]
    seq_pred_t1 = sa[t1][:, None, None] * seq_pred + s1[t1][:, None, None] * noise_seq
    diff = (mask.unsqueeze(-1) * torch.square(noise_pred - noise_z)).sum(dim=(1, 2))
    diff = diff + F.kl_div(torch.log_softmax(seq_pred_t1, dim=-1) * residue_mask.unsqueeze(-1),
                           torch.softmax(seq_t1, dim=-1) * residue_mask.unsqueeze(-1), reduction="none").sum()
    logits = (seq_pred + 1) / 2
    ce = F.cross_entropy(logits.reshape(-1, 21), batch["residue_type"].reshape(-1), reduction="none", ignore_index=0)
    return diff + (ce * mask.reshape(-1)).sum()


def diffusion_loss(sd: SD, cfg, batch: Dict[str, Tensor], x: Tensor, mask: Tensor, t: Tensor,
                   randn_like: Callable[[Tensor], Tensor] = torch.randn_like,
                   outputs: Optional[list] = None) -> Tensor:
    """ProteinReDiffModel.diffusion_loss (model.py:490-526) on a prepared batch; draw order: noise_z, noise_seq.
    ``outputs`` (a list) receives (noise_pred, seq_pred) so that tests can differentiate with respect to them."""
    tab = schedule_tables(cfg.num_steps, cfg.diffusion_schedule)
    seq, residue_mask = batch["residue_one_hot"], batch["residue_mask"]
    noise_z = remove_mean(randn_like(x), mask)
    noise_seq = remove_mean(randn_like(seq), residue_mask)
    z_t, seq_t, seq_t1, t1 = q_sample(tab, x, seq, t, noise_z, noise_seq, batch["residue_extra_mask"],
                                      batch["residue_inv_extra_mask"])
    noise_pred, seq_pred = denoiser_step(sd, cfg, batch, z_t, seq_t, mask, t)
    if outputs is not None:
        noise_pred, seq_pred = noise_pred.detach().requires_grad_(), seq_pred.detach().requires_grad_()
        outputs.extend([noise_pred, seq_pred])
    return loss_from_outputs(tab, batch, noise_pred, seq_pred, noise_z, noise_seq, seq_t1, t1)


def training_loss(sd: SD, cfg, batch: Dict[str, Tensor], randn_like: Callable[[Tensor], Tensor] = torch.randn_like,
                  outputs: Optional[list] = None) -> Tuple[Tensor, Tensor, Tensor]:
    """ProteinReDiffModel.training_step (model.py:528-549) with training_mode=False: prepare_batch
    (one CPU randperm), t ~ randint(0, T, (B,)) from the global CPU generator, loss = mean(diff_loss / num_nodes).
    Returns (loss, diff_loss [B], t)."""
    batch = prepare_batch(batch, cfg.mask_prob)
    x, mask = batch["x"], batch["residue_and_atom_mask"]
    num_nodes = (mask > 0.5).sum(dim=1)
    t = torch.randint(0, cfg.num_steps, size=(x.shape[0],))
    diff = diffusion_loss(sd, cfg, batch, x, mask, t, randn_like, outputs)
    return torch.mean(diff / num_nodes), diff, t
