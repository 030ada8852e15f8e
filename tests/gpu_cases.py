"""GPU parity cases: every C-ABI op (and the whole step / sampler) against the CPU oracle.

Each case returns ``{metric_name: (value, bound)}``; ``tests/test_gpu_parity.py`` asserts
value <= bound, ``tools/gpu_diag.py`` runs every case in its own process (a CUDA fault poisons the
context) and logs the numbers.

Tolerances: the kernels use fp16 tensor-core operands with fp32 accumulation and fp32 LayerNorm /
softmax statistics; operand rounding is 2^-11 relative, so single ops land at ~1e-4..5e-4 relative
L2 and the full 4-block step at <= 1e-3 (north_star tolerance) on coordinates and logits.
"""
from __future__ import annotations

import dataclasses
import math

import torch
import torch.nn.functional as F

from oracle import denoiser_ref as ref
from protein_redesign_b200 import _lib, ops
from protein_redesign_b200 import synthetic as syn
from protein_redesign_b200.model import ProteinReDiffModel

DEV = "cuda:0"
OP_TOL = 2e-3      # single fused op, relative L2 vs fp32 oracle
STEP_TOL = 1e-3    # whole denoiser step (north_star: <= 1e-3 relative on coords and logits)


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(cfg, seed):
    sd = syn.make_state_dict(cfg, seed)
    m = ProteinReDiffModel(cfg)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV).eval(), sd


def _to_dev(batch):
    return {k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def _pair_inputs(cfg, B, N, seed, pad=0):
    g = torch.Generator().manual_seed(seed)
    pair = torch.randn(B, N, N, cfg.pair_dim, generator=g) * 1.5 + 0.3
    single = torch.randn(B, N, cfg.single_dim, generator=g)
    mask = torch.ones(B, N)
    if pad:
        mask[:, N - pad:] = 0
        if B > 1:
            mask[1, N - 2 * pad:] = 0
    return single, pair, mask


# ------------------------------------------------------------------------------------------
# tcgen05 + TMA machinery
# ------------------------------------------------------------------------------------------
def case_gemm(M=256, N=128, K=64, nb1=1, nb2=1, epilogue=False, out_fp16=False, seed=0, bias_act=0):
    g = torch.Generator().manual_seed(seed)
    a = (torch.randn(nb2, nb1, M, K, generator=g)).half()
    b = (torch.randn(nb2, nb1, N, K, generator=g)).half()
    want = torch.matmul(a.float(), b.float().transpose(-1, -2))
    kw = {}
    if epilogue:
        bias = torch.randn(N, generator=g)
        rowscale = torch.rand(nb2, nb1, M, generator=g)
        mul = torch.randn(nb2, nb1, M, N, generator=g)
        add = torch.randn(nb2, nb1, M, N, generator=g)
        want = torch.relu(0.5 * want + bias) * rowscale.unsqueeze(-1) * mul + add
        kw = dict(alpha=0.5, bias=bias.to(DEV), act=1, rowscale=rowscale.to(DEV).contiguous(),
                  mul=mul.to(DEV).contiguous(), add=add.to(DEV).contiguous())
    if bias_act:  # alpha + bias + activation only: the full-line fp16 store path of the single-representation projections
        bias = torch.randn(N, generator=g)
        want = 0.25 * want + bias
        want = torch.relu(want) if bias_act == 1 else torch.sigmoid(want)
        kw = dict(alpha=0.25, bias=bias.to(DEV), act=bias_act)
    out = torch.full((nb2, nb1, M, N), float("nan"), dtype=torch.float16 if out_fp16 else torch.float32, device=DEV)
    _lib.gemm_f16(a.to(DEV).contiguous(), b.to(DEV).contiguous(), out, **kw)
    torch.cuda.synchronize()
    return {"rel": (rel(out.float(), want), 2e-3 if out_fp16 else 1e-5 * math.sqrt(K) + 1e-6)}


# ------------------------------------------------------------------------------------------
# module-level ops vs oracle
# ------------------------------------------------------------------------------------------
def _block_prefix(k=0):
    return f"Denoiser.folding_blocks.{k}."


def case_pair_transition(cfg=syn.PAPER, B=1, N=72, seed=0):
    m, sd = _model(cfg, seed)
    _, pair, _ = _pair_inputs(cfg, B, N, seed)
    want = pair + ref.transition(sd, _block_prefix() + "pair_fc.", pair)
    blk = m.Denoiser.folding_blocks[0]
    p = pair.to(DEV).contiguous()
    ops.pair_transition(cfg, p, blk.pair_fc.packed_pair(), p)
    upd = blk.pair_fc(pair.to(DEV))  # module-level form returns the update
    torch.cuda.synchronize()
    return {"rel": (rel(p, want), OP_TOL), "rel_update": (rel(upd, want - pair), OP_TOL)}


def case_trimul(cfg=syn.PAPER, B=2, N=72, mode="outgoing", seed=0, pad=5):
    m, sd = _model(cfg, seed)
    _, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    upd = ref.triangle_multiplication(sd, _block_prefix() + f"pair_mul_{mode}.", pair, m2, mode)
    mod = getattr(m.Denoiser.folding_blocks[0], f"pair_mul_{mode}")
    p = pair.to(DEV).contiguous()
    mod.apply_(cfg, p, mask.to(DEV))
    got_upd = mod(pair.to(DEV), m2.to(DEV))
    torch.cuda.synchronize()
    return {"rel": (rel(p, pair + upd), OP_TOL), "rel_update": (rel(got_upd, upd), 2 * OP_TOL)}


def case_triattn(cfg=syn.PAPER, B=2, N=72, mode="starting", seed=0, pad=5):
    m, sd = _model(cfg, seed)
    _, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    upd = ref.triangle_attention(sd, _block_prefix() + f"pair_attn_{mode}.", pair, m2, cfg.num_heads, mode)
    mod = getattr(m.Denoiser.folding_blocks[0], f"pair_attn_{mode}")
    p = pair.to(DEV).contiguous()
    mod.apply_(cfg, p, mask.to(DEV))
    got_upd = mod(pair.to(DEV), m2.to(DEV))
    torch.cuda.synchronize()
    return {"rel": (rel(p, pair + upd), OP_TOL), "rel_update": (rel(got_upd, upd), 2 * OP_TOL)}


def case_outer_linear(cfg=syn.PAPER, B=2, N=72, seed=0):
    m, sd = _model(cfg, seed)
    single, pair, _ = _pair_inputs(cfg, B, N, seed)
    upd = ref.outer_linear(sd, _block_prefix() + "outer_linear.", single)
    mod = m.Denoiser.folding_blocks[0].outer_linear
    p = pair.to(DEV).contiguous()
    mod.apply_(cfg, single.to(DEV).contiguous(), p)
    got_upd = mod(single.to(DEV))
    torch.cuda.synchronize()
    return {"rel": (rel(p, pair + upd), OP_TOL), "rel_update": (rel(got_upd, upd), OP_TOL)}


def case_single_attention(cfg=syn.PAPER, B=2, N=72, seed=0, pad=5):
    m, sd = _model(cfg, seed)
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    p = _block_prefix()
    bias = ref.attn_bias_from_pair(sd, p, pair)
    want = single + ref.gated_attention(sd, p + "single_attn.", single, mask, cfg.num_heads, bias)
    blk = m.Denoiser.folding_blocks[0]
    s = single.to(DEV).contiguous()
    ops.single_attention(cfg, s, pair.to(DEV).contiguous(), mask.to(DEV), blk.single_attn.packed_single(blk.attn_bias[1]), s)
    upd = blk.single_attn(single.to(DEV), mask.to(DEV), attn_bias=bias.to(DEV).contiguous())
    torch.cuda.synchronize()
    return {"rel": (rel(s, want), OP_TOL), "rel_update": (rel(upd, want - single), OP_TOL)}


def case_single_transition(cfg=syn.PAPER, B=2, N=72, seed=0):
    m, sd = _model(cfg, seed)
    single, _, _ = _pair_inputs(cfg, B, N, seed)
    want = single + ref.transition(sd, _block_prefix() + "single_fc.", single)
    blk = m.Denoiser.folding_blocks[0]
    s = single.to(DEV).contiguous()
    ops.single_transition(cfg, s, blk.single_fc.packed_single(), s)
    torch.cuda.synchronize()
    return {"rel": (rel(s, want), OP_TOL)}


def case_spattention(cfg=syn.PAPER, B=2, N=72, seed=0):
    m, sd = _model(cfg, seed)
    single, pair, mask = _pair_inputs(cfg, B, N, seed)
    want = ref.single_pair_attention(sd, single, pair, cfg.num_heads)
    got = m.Denoiser.SPAAttnBlock(single.to(DEV), pair.to(DEV), mask.to(DEV), cfg=cfg)
    torch.cuda.synchronize()
    return {"rel": (rel(got, want), OP_TOL)}


def case_opm(cfg=syn.PAPER, B=2, N=72, seed=0, pad=5):
    m, sd = _model(cfg, seed)
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    want = ref.outer_product_update(sd, single, mask)
    got = m.Denoiser.opm(single.to(DEV), mask.to(DEV), cfg=cfg)
    torch.cuda.synchronize()
    return {"rel": (rel(got, want), OP_TOL)}


def case_embeddings(cfg=syn.PAPER, sizes=((9, 50), (12, 60)), seed=0):
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed, two_chains=True)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    want_single = ref.embed_single(sd, pb, seq_t)
    want_static = ref.embed_pair_static(sd, pb, cfg.max_bond_distance, cfg.max_relpos)
    want_pair = want_static + ref.embed_pair_dynamic(sd, z, t, mask, cfg.num_steps)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    want_pair_opm = want_pair + m2.unsqueeze(-1) * ref.outer_product_update(sd, want_single, mask)
    db = _to_dev(pb)
    w = m._weights()
    esm_emb, pair_static = m._static_embeddings(db)
    single = ops.single_embed(cfg, db["atom_feats"], db["atom_mask"], db["residue_mask"], seq_t.to(DEV), esm_emb,
                              w["atom_tabs"], w["w_type"])
    a, b = m.Denoiser.opm.project(cfg, single, mask.to(DEV))
    _, (w_o, b_o) = m.Denoiser.opm.packed_weights()
    pair = torch.empty_like(pair_static)
    ops.pair_embed(cfg, pair_static, z.to(DEV), mask.to(DEV), t.to(DEV), a, b, w["pair_dyn"] + [w_o, b_o], pair)
    # the same with the distance embedding read from the interpolated table instead of the per-pair RBF GEMM
    pair_lut = torch.empty_like(pair_static)
    ops.pair_embed(cfg, pair_static, z.to(DEV), mask.to(DEV), t.to(DEV), a, b, w["pair_dyn"] + [w_o, b_o], pair_lut,
                   rbf_lut=w["rbf_lut"])
    torch.cuda.synchronize()
    # the table itself against the reference's RadialBasisProjection + Linear in fp64 at off-grid distances
    lut = w["rbf_lut"].cpu().double()
    cz = cfg.pair_dim
    inv_h, M = float(lut[0]), int(lut[1])
    dq = torch.linspace(0.0, 2.7, 4001, dtype=torch.float64)
    u = torch.clamp(dq * inv_h, max=float(M))
    mi = torch.clamp(u.floor().long(), max=M - 1)
    fr = (u - mi).unsqueeze(-1)
    rows = lut[cz:].view(M + 1, cz)
    got_tab = rows[mi] * (1 - fr) + rows[mi + 1] * fr
    cen = sd["embed_dist.0.center"].double()
    scale = (cfg.dist_dim - 1) / 2.0
    want_tab = torch.exp(-scale * (dq.unsqueeze(-1) - cen) ** 2) @ sd["embed_dist.1.weight"].double().t()
    return {"single": (rel(single, want_single), OP_TOL), "pair_static": (rel(pair_static, want_static), 1e-6),
            "pair": (rel(pair, want_pair_opm), OP_TOL), "pair_rbf_lut": (rel(pair_lut, want_pair_opm), OP_TOL),
            "rbf_lut_table": (rel(got_tab, want_tab), 2e-5)}


def case_heads(cfg=syn.PAPER, B=2, N=72, seed=0, pad=5):
    m, sd = _model(cfg, seed)
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    g = torch.Generator().manual_seed(seed + 1)
    z = torch.randn(B, N, 3, generator=g)
    sym = 0.5 * (pair + pair.transpose(1, 2))
    want_noise = ref.coord_head(sd, sym, z, mask)
    want_seq = ref.seq_head(sd, single)
    w = m._weights()
    noise = ops.coord_head(cfg, pair.to(DEV).contiguous(), z.to(DEV), mask.to(DEV), w["coord"])
    seq = ops.seq_head(cfg, single.to(DEV).contiguous(), w["seq"])
    p2 = pair.to(DEV).contiguous()
    ops.symmetrize(cfg, p2)
    torch.cuda.synchronize()
    return {"noise": (rel(noise, want_noise), OP_TOL), "seq": (rel(seq, want_seq), OP_TOL),
            "sym_maxabs": (float((p2.cpu() - sym).abs().max()), 0.0)}


def case_pair_bias(cfg=syn.PAPER, B=2, N=72, seed=0):
    m, sd = _model(cfg, seed)
    _, pair, _ = _pair_inputs(cfg, B, N, seed)
    want = ref.attn_bias_from_pair(sd, _block_prefix(), pair)
    blk = m.Denoiser.folding_blocks[0]
    # reach the stream kernel through single_attention's workspace is awkward; use the op with zero single
    # instead: compare through the full single_attention case.  Here: SPA bias path (affine LN) via spattention
    # is covered by case_spattention.  Direct check uses the same kernel via a 1-token trick is not possible,
    # so this case only checks shapes/values through single_attention with single = 0.
    single = torch.zeros(B, N, cfg.single_dim)
    mask = torch.ones(B, N)
    wantd = ref.gated_attention(sd, _block_prefix() + "single_attn.", single, mask, cfg.num_heads, want)
    s = single.to(DEV).contiguous()
    ops.single_attention(cfg, s, pair.to(DEV).contiguous(), mask.to(DEV), blk.single_attn.packed_single(blk.attn_bias[1]), s)
    torch.cuda.synchronize()
    return {"rel": (rel(s, wantd), OP_TOL)}


# ------------------------------------------------------------------------------------------
# whole step / sampler
# ------------------------------------------------------------------------------------------
def case_step(cfg=syn.PAPER, sizes=((12, 60), (9, 50)), seed=3, n_total=None, golden=None, probes=False, row_chunk=None):
    """``row_chunk``: evaluate the oracle's row-independent ops that many pair rows at a time (host memory at N = 1024)."""
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed, n_total=n_total)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    want_probes = {}
    ref.ROW_CHUNK = row_chunk
    try:
        with torch.inference_mode():
            want_noise, want_seq = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t,
                                                     probe=(lambda n, v: want_probes.__setitem__(n, v.clone())) if probes else None)
    finally:
        ref.ROW_CHUNK = None
    torch.manual_seed(seed)
    db = m.prepare_batch(_to_dev(batch))
    got_probes = {}
    with torch.inference_mode():
        noise, seq = m._denoise(db, z.to(DEV), seq_t.to(DEV), mask.to(DEV), t.to(DEV),
                                probe=(lambda n, v: got_probes.__setitem__(n, v.clone())) if probes else None)
    torch.cuda.synchronize()
    out = {"noise": (rel(noise, want_noise), STEP_TOL), "seq": (rel(seq, want_seq), STEP_TOL),
           "prep_keep": (float((db["residue_extra_mask"].cpu() - pb["residue_extra_mask"]).abs().max()), 0.0),
           "pad_noise": (float((noise.cpu() * (1 - mask).unsqueeze(-1)).abs().max()), 0.0)}
    if golden is not None:
        out["noise_vs_reference"] = (rel(noise, torch.from_numpy(golden["noise_pred"])), STEP_TOL)
        out["seq_vs_reference"] = (rel(seq, torch.from_numpy(golden["seq_pred"])), STEP_TOL)
    for n, v in got_probes.items():
        if n in want_probes:
            out["probe:" + n] = (rel(v, want_probes[n]), 5e-3)
    return out


def case_sample(cfg=None, sizes=((8, 32), (6, 27)), seed=5, T=8, graph=True):
    cfg = cfg or dataclasses.replace(syn.README, num_steps=T, mask_prob=0.3)
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed)
    B, N = batch["atom_mask"].shape
    g = torch.Generator().manual_seed(seed + 31337)
    draws = {"z_T": torch.randn(B, N, 3, generator=g), "seq_T": torch.randn(B, N, 21, generator=g),
             "steps": torch.randn(T - 1, B, N, 3, generator=g)}
    order = [draws["z_T"], draws["seq_T"]] + [draws["steps"][i] for i in range(T - 1)]
    it = iter(order)
    torch.manual_seed(seed)
    trace = []
    with torch.inference_mode():
        want_pos, want_logits = ref.sample(sd, cfg, batch, randn_like=lambda x: next(it).clone(), trace=trace)
    torch.manual_seed(seed)
    pos, logits = m.sample(_to_dev(batch), noise=draws, use_cuda_graph=graph)
    torch.cuda.synchronize()
    rmsd = float(((pos.cpu() - want_pos) ** 2).sum(-1).mean().sqrt())
    return {"pos_rel": (rel(pos, want_pos), 5e-3), "logits_rel": (rel(logits, want_logits), 5e-3),
            "rmsd_angstrom": (rmsd, 0.05)}


def case_sample_T50(seed=5, T=50, sizes=((8, 32), (6, 27))):
    """north_star: a 50-step fixed-noise trajectory.  The sampling map of the random-init network is expanding: in the
    fp32 CPU oracle itself a 1e-4 relative perturbation of the weights grows to 0.048 A RMSD / 57 % of the logits after
    50 steps, so the free-running trajectory is held to the stated coordinate bound (0.1 A RMSD, positions
    have 15 A rms) and the per-step contract is checked teacher-forced: the network at the oracle's own state at every
    7th step (<= 1e-3), and the on-device DDPM update at all 50 steps (schedule / noise indexing, last-step branch)."""
    cfg = dataclasses.replace(syn.README, num_steps=T, mask_prob=0.3)
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed)
    B, N = batch["atom_mask"].shape
    g = torch.Generator().manual_seed(seed + 31337)
    draws = {"z_T": torch.randn(B, N, 3, generator=g), "seq_T": torch.randn(B, N, 21, generator=g),
             "steps": torch.randn(T - 1, B, N, 3, generator=g)}
    it = iter([draws["z_T"], draws["seq_T"]] + [draws["steps"][i] for i in range(T - 1)])
    torch.manual_seed(seed)
    trace = []
    with torch.inference_mode():
        want_pos, _ = ref.sample(sd, cfg, batch, randn_like=lambda x: next(it).clone(), trace=trace)
    torch.manual_seed(seed)
    pos, _ = m.sample(_to_dev(batch), noise=draws, use_cuda_graph=True)
    torch.cuda.synchronize()
    out = {"free_running_pos_rel": (rel(pos, want_pos), 5e-3),
           "free_running_rmsd_angstrom": (float(((pos.cpu() - want_pos) ** 2).sum(-1).mean().sqrt()), 0.1)}
    # teacher-forced: the oracle's state before step i
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    mask = pb["residue_and_atom_mask"]
    z0 = ref.remove_mean(draws["z_T"], mask)
    s0 = pb["residue_extra_mask"].unsqueeze(-1) * pb["residue_one_hot"] + \
        pb["residue_inv_extra_mask"].unsqueeze(-1) * ref.remove_mean(draws["seq_T"], pb["residue_mask"])
    states = [(z0, s0.float())] + [(tr[0], torch.softmax(tr[1], -1) * 2 - 1) for tr in trace[:-1]]
    torch.manual_seed(seed)
    db = m.prepare_batch(_to_dev(batch))
    from protein_redesign_b200 import ops
    steps_dev = torch.stack([ref.remove_mean(draws["steps"][i], mask) for i in range(T - 1)]).to(DEV).contiguous()
    worst_n = worst_s = worst_u = 0.0
    with torch.inference_mode():
        for i in range(T):
            z_i, s_i = states[i]
            t_i = torch.full((B,), T - 1 - i, dtype=torch.int64)
            if i % 7 == 0 or i == T - 1:
                n, sp = m.sample_step(db, z_i.to(DEV).contiguous(), s_i.to(DEV).contiguous(), mask.to(DEV), t_i.to(DEV))
                worst_n = max(worst_n, rel(n, trace[i][2]))
                worst_s = max(worst_s, rel(sp, trace[i][1]))
            z_dev, s_dev = z_i.to(DEV).contiguous().clone(), s_i.to(DEV).contiguous().clone()
            state = torch.tensor([T - 1 - i, i], dtype=torch.int32, device=DEV)
            ops.sampler_update(cfg, trace[i][2].to(DEV).contiguous(), trace[i][1].to(DEV).contiguous(), steps_dev, m._coef,
                               z_dev, s_dev, state)
            worst_u = max(worst_u, rel(z_dev, trace[i][0]))
            if i + 1 < T:
                worst_u = max(worst_u, rel(s_dev, states[i + 1][1]))
    torch.cuda.synchronize()
    out["teacher_forced_noise_pred"] = (worst_n, STEP_TOL)
    out["teacher_forced_seq_pred"] = (worst_s, STEP_TOL)
    out["sampler_update_all_steps"] = (worst_u, 1e-5)
    return out


def case_loss(cfg, sizes, seed, golden=None, **batch_kw):
    """a18: training_step's objective on device (q(), the network, the three loss terms, loss = mean(diff_loss / num_nodes))
    and d loss / d (noise_pred, seq_pred) vs the oracle (autograd on the CPU) and vs the reference golden."""
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed, with_positions=True, **batch_kw)
    B, N = batch["atom_mask"].shape
    g = torch.Generator().manual_seed(seed + 4242)
    draws = {"z": torch.randn(B, N, 3, generator=g), "seq": torch.randn(B, N, 21, generator=g)}
    it = iter([draws["z"], draws["seq"]])
    outs = []
    torch.manual_seed(seed)
    want_loss, want_diff, want_t = ref.training_loss(sd, cfg, batch, randn_like=lambda x: next(it).clone(), outputs=outs)
    want_loss.backward()
    detail = {}
    torch.manual_seed(seed)
    with torch.no_grad():
        loss = m.training_step(_to_dev(batch), 0, noise=draws, detail=detail)
    torch.manual_seed(seed)
    vloss = m.validation_step(_to_dev(batch), 0, noise=draws)  # reference model.py:226-247: same objective, no graph
    torch.cuda.synchronize()
    out = {"t_exact": (float((detail["t"].cpu() - want_t).abs().max()), 0.0),
           "validation_step": (abs(float(vloss) - float(loss)), 0.0),
           "loss": (abs(float(loss) - float(want_loss.detach())) / abs(float(want_loss.detach())), STEP_TOL),
           "noise_pred": (rel(detail["noise_pred"], outs[0]), STEP_TOL),
           "seq_pred": (rel(detail["seq_pred"], outs[1]), STEP_TOL),
           "d_noise_pred": (rel(detail["d_noise_pred"], outs[0].grad), 2 * STEP_TOL),
           "d_seq_pred": (rel(detail["d_seq_pred"], outs[1].grad), 2 * STEP_TOL)}
    # the loss kernels alone, on the oracle's own network outputs: tight tolerance (fp32 reductions only)
    from protein_redesign_b200 import ops
    tab = ref.schedule_tables(cfg.num_steps, cfg.diffusion_schedule)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    nz = ref.remove_mean(draws["z"], pb["residue_and_atom_mask"])
    ns = ref.remove_mean(draws["seq"], pb["residue_mask"])
    z_t, seq_t, seq_t1, t1 = ref.q_sample(tab, pb["x"], pb["residue_one_hot"], want_t, nz, ns, pb["residue_extra_mask"],
                                          pb["residue_inv_extra_mask"])
    sched = torch.stack([tab["sqrt_alphas_cumprod"], tab["sqrt_one_minus_alphas_cumprod"]], 1).contiguous().to(DEV)
    dz, dseq, dseq1 = ops.diffusion_q(cfg, pb["x"].to(DEV), pb["residue_one_hot"].float().to(DEV), want_t.to(DEV), nz.to(DEV),
                                      ns.to(DEV), pb["residue_extra_mask"].to(DEV), pb["residue_inv_extra_mask"].to(DEV), sched)
    out["q_z_t"] = (rel(dz, z_t), 1e-6)
    out["q_seq_t"] = (rel(dseq, seq_t), 1e-6)
    out["q_seq_t1"] = (rel(dseq1, seq_t1), 1e-6)
    l2, diff2, terms2, dn2, ds2 = ops.diffusion_loss(
        cfg, outs[0].detach().to(DEV), outs[1].detach().to(DEV), nz.to(DEV), ns.to(DEV), seq_t1.to(DEV),
        pb["residue_and_atom_mask"].to(DEV), pb["residue_mask"].to(DEV), pb["residue_type"].to(DEV), want_t.to(DEV), sched,
        want_grads=True)
    torch.cuda.synchronize()
    out["kernel_diff_loss"] = (rel(diff2, want_diff), 1e-5)
    out["kernel_loss"] = (abs(float(l2) - float(want_loss.detach())) / abs(float(want_loss.detach())), 1e-5)
    out["kernel_d_noise_pred"] = (rel(dn2, outs[0].grad), 1e-5)
    out["kernel_d_seq_pred"] = (rel(ds2, outs[1].grad), 1e-5)
    if golden is not None:
        out["loss_vs_reference"] = (abs(float(loss) - float(golden["loss"])) / abs(float(golden["loss"])), STEP_TOL)
        out["diff_loss_vs_reference"] = (rel(diff2, torch.from_numpy(golden["diff_loss"])), 1e-5)
        out["d_seq_pred_vs_reference"] = (rel(ds2, torch.from_numpy(golden["d_seq_pred"])), 1e-5)
        out["d_noise_pred_vs_reference"] = (rel(dn2, torch.from_numpy(golden["d_noise_pred"])), 1e-5)
        out["z_t_vs_reference"] = (rel(detail["z_t"], torch.from_numpy(golden["z_t"])), 1e-5)
        out["seq_t_vs_reference"] = (rel(detail["seq_t"], torch.from_numpy(golden["seq_t"])), 1e-5)
    return out


def case_invariants(cfg=syn.PAPER, sizes=((10, 54),), seed=7):
    """E(3) equivariance of noise_pred / invariance of seq_pred under a rigid motion of z, and batch-row
    independence: size-independent properties checked on the CUDA path itself."""
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    db = m.prepare_batch(_to_dev(batch))
    g = torch.Generator().manual_seed(0)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    shift = torch.randn(1, 1, 3, generator=g)
    with torch.inference_mode():
        n0, s0 = m._denoise(db, z.to(DEV), seq_t.to(DEV), mask.to(DEV), t.to(DEV))
        n0, s0 = n0.clone(), s0.clone()
        n1, s1 = m._denoise(db, (z @ q + shift).to(DEV).contiguous(), seq_t.to(DEV), mask.to(DEV), t.to(DEV))
    torch.cuda.synchronize()
    mean = (mask.unsqueeze(-1) * n0.cpu()).sum(1) / mask.sum(1, keepdim=True)
    return {"equivariance": (rel(n1, n0.cpu() @ q), 2e-3), "invariance": (rel(s1, s0), 2e-3),
            "zero_mean": (float(mean.abs().max()), 1e-5)}


def case_batch_rows(cfg=syn.PAPER, sizes=((32, 480), (20, 400), (25, 487), (10, 380), (32, 480), (8, 300), (30, 482), (16, 430)),
                    seed=21, oracle_row=3, check_rows=(0, 3, 7)):
    """The benchmarked shape (B = 8, N = 512), ragged: row k of the batched output must equal the B = 1 output of complex k
    (the network never mixes batch rows), and one padded row is compared with the oracle directly -- so the B = 8
    run inherits the B = 1 oracle checks."""
    m, sd = _model(cfg, seed)
    batch = syn.make_batch(cfg, list(sizes), seed=seed, n_total=512)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    db = _to_dev(pb)
    with torch.inference_mode():
        n8, s8 = m._denoise(db, z.to(DEV), seq_t.to(DEV), mask.to(DEV), t.to(DEV))
        n8, s8 = n8.clone(), s8.clone()
        worst_n = worst_s = 0.0
        for k in check_rows:
            row = {key: (v[k:k + 1].contiguous() if isinstance(v, torch.Tensor) and v.dim() >= 1 else v) for key, v in db.items()}
            n1, s1 = m._denoise(row, z[k:k + 1].to(DEV), seq_t[k:k + 1].to(DEV), mask[k:k + 1].to(DEV), t[k:k + 1].to(DEV))
            worst_n = max(worst_n, rel(n8[k:k + 1], n1))
            worst_s = max(worst_s, rel(s8[k:k + 1], s1))
        k = oracle_row
        prow = {key: (v[k:k + 1] if isinstance(v, torch.Tensor) and v.dim() >= 1 else v) for key, v in pb.items()}
        want_n, want_s = ref.denoiser_step(sd, cfg, prow, z[k:k + 1], seq_t[k:k + 1], mask[k:k + 1], t[k:k + 1])
    torch.cuda.synchronize()
    return {"row_independence_noise": (worst_n, 1e-5), "row_independence_seq": (worst_s, 1e-5),
            "b8_row_vs_oracle_noise": (rel(n8[k:k + 1], want_n), STEP_TOL), "b8_row_vs_oracle_seq": (rel(s8[k:k + 1], want_s), STEP_TOL),
            "pad_noise": (float((n8.cpu() * (1 - mask).unsqueeze(-1)).abs().max()), 0.0)}


def case_denoiser_forward(cfg=syn.PAPER, B=2, N=72, seed=13, pad=5):
    """The public entry modules.Denoiser.forward(batch, z, t, single, pair, cache) (reference modules.py:391-404): OPM
    added to the caller's pair in place, SPAttention, the folding blocks, symmetrisation."""
    m, sd = _model(cfg, seed)
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    with torch.inference_mode():
        want_single, want_pair = ref.denoiser_trunk(sd, cfg, single, pair, mask)
        p = pair.to(DEV).contiguous()
        got_single, got_pair, cache = m.Denoiser({"residue_and_atom_mask": mask.to(DEV)}, None, None, single.to(DEV), p, "cache")
    torch.cuda.synchronize()
    return {"single": (rel(got_single, want_single), STEP_TOL), "pair": (rel(got_pair, want_pair), STEP_TOL),
            "in_place": (0.0 if got_pair.data_ptr() == p.data_ptr() and cache == "cache" else 1.0, 0.0)}


def case_folding_block_forward(cfg=syn.PAPER, B=2, N=72, seed=14, pad=5):
    """modules.FoldingBlock.forward(single, pair, mask) (reference modules.py:328-343) returns NEW tensors."""
    m, sd = _model(cfg, seed)
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    with torch.inference_mode():
        want_single, want_pair = ref.folding_block(sd, 1, single, pair, mask, cfg.num_heads)
        s_in, p_in = single.to(DEV), pair.to(DEV)
        got_single, got_pair = m.Denoiser.folding_blocks[1](s_in, p_in, mask.to(DEV))
    torch.cuda.synchronize()
    return {"single": (rel(got_single, want_single), OP_TOL), "pair": (rel(got_pair, want_pair), OP_TOL),
            "inputs_untouched": (float((p_in.cpu() - pair).abs().max() + (s_in.cpu() - single).abs().max()), 0.0)}


class _ShadowEMA:
    """torch_ema.ExponentialMovingAverage's weight swap, restated: average_parameters() copies the shadow weights in
    through ``param.data.copy_`` (no version counter sees it) and restores the originals on exit."""

    def __init__(self, params, shadow):
        self.params, self.shadow = list(params), [s.clone() for s in shadow]

    def average_parameters(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            saved = [p.data.clone() for p in self.params]
            for p, s in zip(self.params, self.shadow):
                p.data.copy_(s)
            try:
                yield
            finally:
                for p, s in zip(self.params, saved):
                    p.data.copy_(s)
        return ctx()


def case_predict_step(seed=6, T=6, sizes=((8, 32), (6, 27))):
    """predict_step (reference model.py:249-252) = sample() under the EMA weights: checked against the oracle run with the
    SHADOW weights, then sample() again must be back on the raw weights (packed fp16 copies invalidated both ways)."""
    cfg = dataclasses.replace(syn.README, num_steps=T, mask_prob=0.3)
    m, sd = _model(cfg, seed)
    sd_ema = syn.make_state_dict(cfg, seed + 100)
    batch = syn.make_batch(cfg, list(sizes), seed=seed)
    B, N = batch["atom_mask"].shape
    g = torch.Generator().manual_seed(seed + 31337)
    draws = {"z_T": torch.randn(B, N, 3, generator=g), "seq_T": torch.randn(B, N, 21, generator=g),
             "steps": torch.randn(T - 1, B, N, 3, generator=g)}

    def oracle(weights):
        it = iter([draws["z_T"], draws["seq_T"]] + [draws["steps"][i] for i in range(T - 1)])
        torch.manual_seed(seed)
        with torch.inference_mode():
            return ref.sample(weights, cfg, batch, randn_like=lambda x: next(it).clone())

    want_raw, want_ema = oracle(sd), oracle(sd_ema)
    torch.manual_seed(seed)
    pos_a, log_a = m.sample(_to_dev(batch), noise=draws)  # first pack happens on the RAW weights
    names = [n for n, _ in m.named_parameters()]
    m.ema = _ShadowEMA([p for _, p in m.named_parameters()], [sd_ema[n].to(DEV) for n in names])
    torch.manual_seed(seed)
    pos_b, log_b = m.predict_step(_to_dev(batch), 0, noise=draws)
    torch.manual_seed(seed)
    pos_c, log_c = m.sample(_to_dev(batch), noise=draws)
    torch.cuda.synchronize()
    return {"raw_pos": (rel(pos_a, want_raw[0]), 5e-3), "raw_logits": (rel(log_a, want_raw[1]), 5e-3),
            "ema_pos": (rel(pos_b, want_ema[0]), 5e-3), "ema_logits": (rel(log_b, want_ema[1]), 5e-3),
            "restored_pos": (rel(pos_c, pos_a), 0.0), "restored_logits": (rel(log_c, log_a), 0.0)}


def case_back_to_back_batches(cfg=syn.README, sizes=((8, 32), (6, 27)), seed=15):
    """Two DIFFERENT same-shape batches through one model, the first one freed before the second is created (so the
    caching allocator hands the same addresses out again): each must match the oracle (the step-invariant
    embedding cache was keyed on raw data_ptr values)."""
    m, sd = _model(cfg, seed)
    out = {}
    for tag, bseed in (("first", seed), ("second", seed + 1)):
        batch = syn.make_batch(cfg, list(sizes), seed=bseed)
        z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, bseed)
        torch.manual_seed(bseed)
        pb = ref.prepare_batch(batch, cfg.mask_prob)
        with torch.inference_mode():
            want_n, want_s = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
            torch.manual_seed(bseed)
            db = m.prepare_batch(_to_dev(batch))
            n, s_ = m.sample_step(db, z.to(DEV), seq_t.to(DEV), mask.to(DEV), t.to(DEV))
            out[tag + "_noise"] = (rel(n, want_n), STEP_TOL)
            out[tag + "_seq"] = (rel(s_, want_s), STEP_TOL)
        del db, n, s_
        torch.cuda.synchronize()
    return out


CASES = {
    "gemm_basic": lambda: case_gemm(256, 128, 64),
    "gemm_k512": lambda: case_gemm(384, 256, 512),
    "gemm_tails": lambda: case_gemm(200, 72, 136),
    "gemm_n48": lambda: case_gemm(130, 48, 64),
    "gemm_batch": lambda: case_gemm(128, 128, 128, nb1=3, nb2=2),
    "gemm_epilogue": lambda: case_gemm(200, 136, 128, nb1=2, epilogue=True),
    "gemm_fp16_out": lambda: case_gemm(256, 128, 256, out_fp16=True),
    "gemm_fp16_relu": lambda: case_gemm(384, 512, 128, out_fp16=True, bias_act=1),
    "gemm_fp16_sigmoid": lambda: case_gemm(256, 256, 64, nb1=2, out_fp16=True, bias_act=2),
    "pair_transition": lambda: case_pair_transition(),
    "pair_transition_readme": lambda: case_pair_transition(syn.README, 2, 40),
    "trimul_outgoing": lambda: case_trimul(mode="outgoing"),
    "trimul_incoming": lambda: case_trimul(mode="incoming"),
    "trimul_readme": lambda: case_trimul(syn.README, 2, 40, "incoming"),
    # N = 140: plane_ld = 192, the second 64-column half of the second k-tile lies outside the padded plane row
    "trimul_n140": lambda: case_trimul(B=1, N=140, mode="incoming", pad=7),
    "trimul_n300": lambda: case_trimul(B=1, N=300, mode="outgoing", pad=11),
    # N % 128 == 0: TMA-staged contraction-result tiles and the two-threads-per-row output kernel
    "trimul_n256": lambda: case_trimul(B=2, N=256, mode="incoming", pad=9),
    "trimul_n128": lambda: case_trimul(B=1, N=128, mode="outgoing", pad=3),
    "triattn_starting": lambda: case_triattn(mode="starting"),
    "triattn_ending": lambda: case_triattn(mode="ending"),
    "triattn_n200": lambda: case_triattn(B=1, N=200, mode="ending", pad=9),
    "triattn_readme": lambda: case_triattn(syn.README, 2, 40, "starting"),
    # N = 140: ragged 16-byte vectors in the transposed v tile, two query tiles, padded last key tile
    "triattn_n140": lambda: case_triattn(B=1, N=140, mode="starting", pad=6),
    "triattn_n300": lambda: case_triattn(B=1, N=300, mode="ending", pad=13),
    # N = 512: four query tiles per sequence -> the four-group kernel (prd_triattn4.cu), with masked key tiles
    "triattn_n512": lambda: case_triattn(B=1, N=512, mode="starting", pad=70),
    "pair_transition_n140": lambda: case_pair_transition(syn.PAPER, 1, 140),
    "outer_linear_n300": lambda: case_outer_linear(syn.PAPER, 1, 300),
    "outer_linear": lambda: case_outer_linear(),
    "outer_linear_readme": lambda: case_outer_linear(syn.README, 2, 140),
    "single_attention": lambda: case_single_attention(),
    "single_transition": lambda: case_single_transition(),
    "spattention": lambda: case_spattention(),
    "opm": lambda: case_opm(),
    "embeddings": lambda: case_embeddings(),
    "embeddings_readme": lambda: case_embeddings(syn.README),
    # N % 128 == 0 with padding: the resident-operand table kernel (pair_embed_lut_kernel)
    "embeddings_n128": lambda: case_embeddings(sizes=((16, 112), (12, 100)), seed=3),
    "embeddings_n256": lambda: case_embeddings(sizes=((30, 226),), seed=4),
    "heads": lambda: case_heads(),
    "pair_bias": lambda: case_pair_bias(),
    "step_paper_n72": lambda: case_step(probes=True),
    "step_readme_n40": lambda: case_step(syn.README, ((8, 32), (6, 27)), seed=2),
    "step_paper_n128": lambda: case_step(syn.PAPER, ((16, 112),), seed=4),
    "sample_eager": lambda: case_sample(graph=False),
    "sample_graph": lambda: case_sample(graph=True),
    "invariants": lambda: case_invariants(),
    # BASELINE.json's full sizes, through size-independent properties (no oracle at these sizes): config 2 (~300 tokens),
    # config 3 (batch of 8 x 512 tokens), config 5 (1024 tokens = 1023 residues + 1 dummy atom)
    "invariants_n300": lambda: case_invariants(syn.PAPER, sizes=((30, 270),), seed=8),
    "invariants_n512_b8": lambda: case_invariants(syn.PAPER, sizes=((32, 480),) * 8, seed=9),
    "invariants_n1024": lambda: case_invariants(syn.PAPER, sizes=((1, 1023),), seed=10),
    "sample_graph_T50": lambda: case_sample_T50(),
    # ---- round 2: oracle parity at BASELINE.json's sizes ----
    "step_n512": lambda: case_step(syn.PAPER, ((32, 480),), seed=31),                                  # config 3, one complex
    "step_n512_b2_ragged": lambda: case_step(syn.PAPER, ((32, 430), (20, 371)), seed=32, n_total=512),   # masked key tiles of the g4 core
    "step_n300": lambda: case_step(syn.PAPER, ((30, 270),), seed=33),                                  # config 2
    "step_n1024": lambda: case_step(syn.PAPER, ((1, 1023),), seed=34, row_chunk=64),                   # config 5
    "batch_rows_b8_n512": lambda: case_batch_rows(),
    "trimul_n512_outgoing": lambda: case_trimul(B=1, N=512, mode="outgoing", pad=37),
    "trimul_n512_incoming": lambda: case_trimul(B=1, N=512, mode="incoming", pad=37),
    "triattn_n512_ending": lambda: case_triattn(B=1, N=512, mode="ending", pad=70),
    "outer_linear_n512": lambda: case_outer_linear(syn.PAPER, 1, 512),
    "pair_transition_n512": lambda: case_pair_transition(syn.PAPER, 1, 512),
    "heads_n512": lambda: case_heads(syn.PAPER, 1, 512, seed=3, pad=41),
    "embeddings_n512": lambda: case_embeddings(sizes=((32, 450), (20, 492)), seed=5),
    "denoiser_forward": lambda: case_denoiser_forward(),
    "folding_block_forward": lambda: case_folding_block_forward(),
    "predict_step_ema": lambda: case_predict_step(),
    "back_to_back_batches": lambda: case_back_to_back_batches(),
    "loss_paper_n72": lambda: case_loss(dataclasses.replace(syn.PAPER, mask_prob=0.15, num_steps=2000),
                                        ((12, 60), (9, 50)), seed=11),
}


# ------------------------------------------------------------------------------------------
# backward pass: every prd_<op>_bwd against autograd through the oracle, then the whole training step
# ------------------------------------------------------------------------------------------
BWD_TOL = 3e-3   # per-tensor relative L2 of a gradient (tf32 operand products, fp32 accumulation and reductions)


def _bwd_setup(cfg, seed, names_prefixes):
    """(model on the GPU, P, G, oracle state-dict whose tensors under the given prefixes require grad)."""
    from protein_redesign_b200 import autograd as ag  # noqa: F401
    m, sd = _model(cfg, seed)
    sd = {k: v.clone() for k, v in sd.items()}
    for k in sd:
        if any(k.startswith(p) for p in names_prefixes) and sd[k].is_floating_point() and k not in ("embed_beta.0.weight", "embed_dist.0.center"):
            sd[k].requires_grad_()
    P = {n: p.detach().contiguous() for n, p in m.named_parameters()}
    G = {n: torch.zeros_like(p) for n, p in P.items()}
    return m, sd, P, G


ZERO_GRAD_FLOOR = 1e-3  # see _grad_rel


def _grad_rel(got, want, global_norm):
    """Relative L2 error of one gradient tensor.  Five tensors of the model have an analytically ZERO gradient (a constant
    shift of attention logits along the key axis cancels in the softmax: FoldingBlock.attn_bias.1.bias x4,
    SPAttention.linear_z.0.bias): what any implementation returns for them is rounding noise, so the denominator is
    floored at ZERO_GRAD_FLOOR of the norm of the whole gradient."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).norm() / max(float(want.norm()), ZERO_GRAD_FLOOR * global_norm, 1e-30))


def _param_metrics(out, sd, G, tol=BWD_TOL):
    req = {k: v for k, v in sd.items() if v.requires_grad}
    for k, v in req.items():
        assert v.grad is not None, k
    gn = math.sqrt(sum(float(v.grad.double().norm()) ** 2 for v in req.values()))
    for k, v in req.items():
        out["dW:" + k] = (_grad_rel(G[k], v.grad, gn), tol)
    return out


def _rand_like(x, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(x.shape, generator=g)


def case_bwd_transition(cfg=syn.PAPER, B=2, N=40, seed=40, which="pair_fc"):
    from protein_redesign_b200 import autograd as ag
    prefix = _block_prefix() + which + "."
    m, sd, P, G = _bwd_setup(cfg, seed, [prefix])
    single, pair, _ = _pair_inputs(cfg, B, N, seed)
    x = (pair if which == "pair_fc" else single).clone().requires_grad_()
    out = x + ref.transition(sd, prefix, x)
    dy = _rand_like(out, seed + 1)
    out.backward(dy)
    d = dy.to(DEV).contiguous()
    ag.transition_bwd(cfg, P, G, prefix, x.detach().to(DEV).contiguous(), d)
    torch.cuda.synchronize()
    return _param_metrics({"dx": (rel(d, x.grad), BWD_TOL)}, sd, G)


def case_bwd_triattn(cfg=syn.PAPER, B=2, N=40, mode="starting", seed=41, pad=5):
    from protein_redesign_b200 import autograd as ag
    prefix = _block_prefix() + f"pair_attn_{mode}."
    m, sd, P, G = _bwd_setup(cfg, seed, [prefix])
    _, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    x = pair.clone().requires_grad_()
    out = x + ref.triangle_attention(sd, prefix, x, m2, cfg.num_heads, mode)
    dy = _rand_like(out, seed + 1)
    out.backward(dy)
    d = dy.to(DEV).contiguous()
    ag.triangle_attention_bwd(cfg, P, G, prefix + "attn.", 1 if mode == "ending" else 0, pair.to(DEV).contiguous(), mask.to(DEV), d)
    torch.cuda.synchronize()
    return _param_metrics({"dx": (rel(d, x.grad), BWD_TOL)}, sd, G)


def case_bwd_trimul(cfg=syn.PAPER, B=2, N=40, mode="outgoing", seed=42, pad=5):
    from protein_redesign_b200 import autograd as ag
    prefix = _block_prefix() + f"pair_mul_{mode}."
    m, sd, P, G = _bwd_setup(cfg, seed, [prefix])
    _, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    x = pair.clone().requires_grad_()
    out = x + ref.triangle_multiplication(sd, prefix, x, m2, mode)
    dy = _rand_like(out, seed + 1)
    out.backward(dy)
    d = dy.to(DEV).contiguous()
    ag.triangle_multiplication_bwd(cfg, P, G, prefix, 1 if mode == "incoming" else 0, pair.to(DEV).contiguous(), mask.to(DEV), d)
    torch.cuda.synchronize()
    return _param_metrics({"dx": (rel(d, x.grad), BWD_TOL)}, sd, G)


def case_bwd_outer_linear(cfg=syn.PAPER, B=2, N=40, seed=43):
    from protein_redesign_b200 import autograd as ag
    prefix = _block_prefix() + "outer_linear."
    m, sd, P, G = _bwd_setup(cfg, seed, [prefix])
    single, pair, _ = _pair_inputs(cfg, B, N, seed)
    s = single.clone().requires_grad_()
    out = ref.outer_linear(sd, prefix, s)
    dy = _rand_like(out, seed + 1)
    out.backward(dy)
    d_single = torch.zeros(B, N, cfg.single_dim, device=DEV)
    ag.outer_linear_bwd(cfg, P, G, prefix, single.to(DEV).contiguous(), dy.to(DEV).contiguous(), d_single)
    torch.cuda.synchronize()
    return _param_metrics({"d_single": (rel(d_single, s.grad), BWD_TOL)}, sd, G)


def case_bwd_single_attention(cfg=syn.PAPER, B=2, N=40, seed=44, pad=5):
    from protein_redesign_b200 import autograd as ag
    p = _block_prefix()
    m, sd, P, G = _bwd_setup(cfg, seed, [p + "single_attn.", p + "attn_bias."])
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    s, pr = single.clone().requires_grad_(), pair.clone().requires_grad_()
    out = s + ref.gated_attention(sd, p + "single_attn.", s, mask, cfg.num_heads, ref.attn_bias_from_pair(sd, p, pr))
    dy = _rand_like(out, seed + 1)
    out.backward(dy)
    d_single = dy.to(DEV).contiguous()
    d_pair = torch.zeros_like(pair, device=DEV)
    ag.single_attention_bwd(cfg, P, G, p, single.to(DEV).contiguous(), pair.to(DEV).contiguous(), mask.to(DEV), d_single, d_pair)
    torch.cuda.synchronize()
    return _param_metrics({"d_single": (rel(d_single, s.grad), BWD_TOL), "d_pair": (rel(d_pair, pr.grad), BWD_TOL)}, sd, G)


def case_bwd_spattention(cfg=syn.PAPER, B=2, N=40, seed=45):
    from protein_redesign_b200 import autograd as ag
    m, sd, P, G = _bwd_setup(cfg, seed, ["Denoiser.SPAAttnBlock."])
    single, pair, mask = _pair_inputs(cfg, B, N, seed)
    s, pr = single.clone().requires_grad_(), pair.clone().requires_grad_()
    out = ref.single_pair_attention(sd, s, pr, cfg.num_heads)
    dy = _rand_like(out, seed + 1)
    out.backward(dy)
    d_single = dy.to(DEV).contiguous()
    d_pair = torch.zeros_like(pair, device=DEV)
    ag.spattention_bwd(cfg, P, G, single.to(DEV).contiguous(), pair.to(DEV).contiguous(), d_single, d_pair)
    torch.cuda.synchronize()
    return _param_metrics({"d_single": (rel(d_single, s.grad), BWD_TOL), "d_pair": (rel(d_pair, pr.grad), BWD_TOL)}, sd, G)


def case_bwd_heads(cfg=syn.PAPER, B=2, N=40, seed=46, pad=5):
    from protein_redesign_b200 import autograd as ag
    m, sd, P, G = _bwd_setup(cfg, seed, ["weight_radial.", "seq_mlp."])
    single, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
    z = _rand_like(torch.empty(B, N, 3), seed + 2)
    s, pr = single.clone().requires_grad_(), pair.clone().requires_grad_()
    noise = ref.coord_head(sd, 0.5 * (pr + pr.transpose(1, 2)), z, mask)
    seq = ref.seq_head(sd, s)
    dn, dsq = _rand_like(noise, seed + 3), _rand_like(seq, seed + 4)
    (noise * dn).sum().backward()
    (seq * dsq).sum().backward()
    d_single = torch.empty(B, N, cfg.single_dim, device=DEV)
    d_pair = torch.empty_like(pair, device=DEV)
    ag.seq_head_bwd(cfg, P, G, single.to(DEV).contiguous(), dsq.to(DEV).contiguous(), d_single)
    ag.coord_head_bwd(cfg, P, G, pair.to(DEV).contiguous(), z.to(DEV), mask.to(DEV), dn.to(DEV).contiguous(), d_pair)
    torch.cuda.synchronize()
    return _param_metrics({"d_single": (rel(d_single, s.grad), BWD_TOL), "d_pair": (rel(d_pair, pr.grad), BWD_TOL)}, sd, G)


def case_bwd_embeddings(cfg=syn.PAPER, sizes=((9, 30), (12, 26)), seed=47):
    """pair_embed_bwd + opm_project_bwd + single_embed_bwd: everything upstream of the trunk (reference model.py:332-361,
    modules.py:391-397)."""
    from protein_redesign_b200 import autograd as ag
    m, sd, P, G = _bwd_setup(cfg, seed, ["Denoiser.opm.", "embed_"])
    batch = syn.make_batch(cfg, list(sizes), seed=seed, two_chains=True)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    single = ref.embed_single(sd, pb, seq_t)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    pair = ref.embed_pair_static(sd, pb, cfg.max_bond_distance, cfg.max_relpos) + ref.embed_pair_dynamic(sd, z, t, mask, cfg.num_steps) \
        + m2.unsqueeze(-1) * ref.outer_product_update(sd, single, mask)
    dp, ds = _rand_like(pair, seed + 1), _rand_like(single, seed + 2)
    ((pair * dp).sum() + (single * ds).sum()).backward()
    db = _to_dev(pb)
    w = m._weights()
    esm_emb, _ = m._static_embeddings(db)
    single_dev = ops.single_embed(cfg, db["atom_feats"], db["atom_mask"], db["residue_mask"], seq_t.to(DEV), esm_emb, w["atom_tabs"], w["w_type"])
    a, b = m.Denoiser.opm.project(cfg, single_dev, mask.to(DEV))
    d_single = ds.to(DEV).contiguous()
    d_a, d_b = ag.pair_embed_bwd(cfg, P, G, db, dp.to(DEV).contiguous(), z.to(DEV), mask.to(DEV), t.to(DEV), a, b)
    ag.opm_project_bwd(cfg, P, G, single_dev, mask.to(DEV), d_a, d_b, d_single)
    ag.single_embed_bwd(cfg, P, G, db, seq_t.to(DEV), d_single)
    torch.cuda.synchronize()
    return _param_metrics({}, sd, G)


STEP_GRAD_TOL = 3e-2     # worst per-tensor relative L2 of a parameter gradient through the whole step (see case_train_step)
STEP_GRAD_MEDIAN = 3e-3  # median over the 240 tensors


def case_train_step(cfg, sizes, seed, golden=None, tol=STEP_GRAD_TOL, top=4, **batch_kw):
    """training_step under autograd: loss.backward() through the CUDA backward kernels; every parameter gradient against
    autograd through the oracle (per-tensor relative L2) and against the REFERENCE's gradient fingerprints (norm +
    seeded random projection of all 240 tensors, tests/golden/loss_*.npz; bound 1e-3, the acceptance criterion).

    Why the per-tensor L2 bound through the whole step is 3e-2 while every op's own backward meets 3e-3 (bwd_* cases): the
    network has ReLU layers (single_fc, pair_fc, seq_mlp, weight_radial).  The forward activations differ from the fp32
    oracle's by eps ~ 1e-4 (the step tolerance), so a fraction ~0.8 eps of the ReLU pre-activations has the other sign, and
    each flipped element carries a full-size error in d(hidden): relative L2 ~ sqrt(0.8 eps) ~ 1e-2 on those layers'
    gradients and, diluted, on everything upstream.  The errors are sparse and random-signed, so they vanish in the
    fingerprints (norm, projection) and in any sum over rows; the reference's own training runs fp16 AMP (train.py:37),
    whose eps ~ 1e-3 puts it at ~3e-2 by the same argument."""
    m, sd = _model(cfg, seed)
    m.train()
    sdg = {k: (v.clone().requires_grad_() if v.is_floating_point() and k not in ("embed_beta.0.weight", "embed_dist.0.center") else v)
           for k, v in sd.items()}
    batch = syn.make_batch(cfg, list(sizes), seed=seed, with_positions=True, **batch_kw)
    B, N = batch["atom_mask"].shape
    g = torch.Generator().manual_seed(seed + 4242)
    draws = {"z": torch.randn(B, N, 3, generator=g), "seq": torch.randn(B, N, 21, generator=g)}
    it = iter([draws["z"], draws["seq"]])
    torch.manual_seed(seed)
    want_loss, _, _ = ref.training_loss(sdg, cfg, batch, randn_like=lambda x: next(it).clone())
    want_loss.backward()
    torch.manual_seed(seed)
    with torch.enable_grad():
        loss = m.training_step(_to_dev(batch), 0, noise=draws)
        loss.backward()
    torch.cuda.synchronize()
    out = {"loss": (abs(float(loss) - float(want_loss.detach())) / abs(float(want_loss.detach())), STEP_TOL)}
    grads = {n: p.grad for n, p in m.named_parameters() if p.requires_grad}
    req = {k: v for k, v in sdg.items() if v.requires_grad}
    gn = math.sqrt(sum(float(v.grad.double().norm()) ** 2 for v in req.values()))
    errs = []
    for k, v in req.items():
        assert grads.get(k) is not None, f"no gradient for {k}"
        errs.append((_grad_rel(grads[k], v.grad, gn), k))
    errs.sort(reverse=True)
    for e, k in errs[:top]:
        out[f"grad_rel_l2[{k}]"] = (e, tol)
    out["median_grad_rel_l2"] = (errs[len(errs) // 2][0], STEP_GRAD_MEDIAN)
    out["tensors_without_grad"] = (float(240 - len(errs)), 0.0)
    if golden is not None:
        gp = torch.Generator().manual_seed(seed + 777)
        gnorm = math.sqrt(sum(float(x) ** 2 for x in golden["grad_norms"]))
        fw, fname = 0.0, ""
        for n, want_norm, want_proj in zip([str(x) for x in golden["grad_names"]], golden["grad_norms"], golden["grad_projs"]):
            grad = grads[n].detach().cpu()
            dvec = torch.randn(grad.shape, generator=gp)
            scale = max(float(want_norm), ZERO_GRAD_FLOOR * gnorm, 1e-12)
            e = max(abs(float(grad.norm()) - float(want_norm)) / scale,
                    abs(float((grad * dvec).sum()) - float(want_proj)) / (scale * float(dvec.norm())))
            if e > fw:
                fw, fname = e, n
        out[f"reference_gradient_fingerprints[worst: {fname}]"] = (fw, 1e-3)
    return out


CASES.update({
    "gemm_tf32": lambda: case_gemm_tf32(),
    "gemm_tf32_batch_tails": lambda: case_gemm_tf32(200, 72, 100, nb1=3),
    "bwd_pair_fc": lambda: case_bwd_transition(which="pair_fc"),
    "bwd_single_fc": lambda: case_bwd_transition(which="single_fc"),
    "bwd_triattn_starting": lambda: case_bwd_triattn(mode="starting"),
    "bwd_triattn_ending": lambda: case_bwd_triattn(mode="ending"),
    "bwd_triattn_n140": lambda: case_bwd_triattn(B=1, N=140, mode="ending", pad=9),
    "bwd_trimul_outgoing": lambda: case_bwd_trimul(mode="outgoing"),
    "bwd_trimul_incoming": lambda: case_bwd_trimul(mode="incoming"),
    "bwd_trimul_n75": lambda: case_bwd_trimul(B=1, N=75, mode="incoming", pad=4),
    "bwd_outer_linear": lambda: case_bwd_outer_linear(),
    "bwd_single_attention": lambda: case_bwd_single_attention(),
    "bwd_spattention": lambda: case_bwd_spattention(),
    "bwd_heads": lambda: case_bwd_heads(),
    "bwd_embeddings": lambda: case_bwd_embeddings(),
    "bwd_embeddings_readme": lambda: case_bwd_embeddings(syn.README),
    "train_step_readme_n40": lambda: case_train_step(dataclasses.replace(syn.README, mask_prob=0.15, num_steps=2000),
                                                     ((8, 32), (6, 27)), 10),
    "train_step_paper_n72": lambda: case_train_step(dataclasses.replace(syn.PAPER, mask_prob=0.15, num_steps=2000),
                                                    ((12, 60), (9, 50)), 11),
})


def case_gemm_tf32(M=256, N=128, K=64, nb1=1, seed=0):
    """fp32 operands on kind::tf32: operands pre-rounded to tf32 make every product exact, so only the fp32 accumulation
    order differs from the float64 reference."""
    g = torch.Generator().manual_seed(seed)

    def r_tf32(x):  # round to nearest, ties away (cvt.rna): add half an ulp of the 10-bit mantissa, clear 13 bits
        i = x.view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

    a = r_tf32(torch.randn(1, nb1, M, K, generator=g))
    b = r_tf32(torch.randn(1, nb1, N, K, generator=g))
    want = torch.matmul(a.double(), b.double().transpose(-1, -2))
    out = torch.full((1, nb1, M, N), float("nan"), dtype=torch.float32, device=DEV)
    _lib.gemm_f16(a.to(DEV).contiguous(), b.to(DEV).contiguous(), out)
    torch.cuda.synchronize()
    return {"rel": (rel(out, want), 2e-6)}


# ------------------------------------------------------------------------------------------
# Lightning-free predict loop, GPU post-processing
# ------------------------------------------------------------------------------------------
def _kabsch_np(P, Q):
    """float64 SVD Kabsch (row convention: aligned = t + P @ R) -> (R, t, rmsd)."""
    import numpy as np
    cp, cq = P.mean(0), Q.mean(0)
    H = (P - cp).T @ (Q - cq)
    U, S, Vt = np.linalg.svd(H)
    d = np.sign(np.linalg.det(U @ Vt))
    D = np.diag([1.0, 1.0, d])
    R = U @ D @ Vt
    t = cq - cp @ R
    diff = t + P @ R - Q
    return R, t, float(np.sqrt((diff ** 2).sum(-1).mean()))


def case_postprocess(B=5, N=150, seed=50):
    import numpy as np
    from protein_redesign_b200 import postprocess as post
    g = torch.Generator().manual_seed(seed)
    out = {}
    # ---- decode (reference generate.py:76-91) ----
    logits = torch.randn(B, N, 21, generator=g)
    rmask = torch.ones(B, N)
    rmask[:, :7] = 0
    rmask[1, N - 20:] = 0
    tokens = post.decode_tokens(logits.to(DEV), rmask.to(DEV)).cpu()
    want_tok = torch.argmax(torch.softmax(logits, dim=-1), dim=-1) * rmask.long()
    out["tokens_exact"] = (float((tokens != want_tok).sum()), 0.0)
    letters = ["X"] + post.RESIDUE_TYPES
    want_seq = ["".join(letters[i] for i in row).lstrip("X").rstrip("X") for row in want_tok.tolist()]
    out["sequences_exact"] = (float(sum(a != b for a, b in zip(post.trimmed_sequence(logits.to(DEV), rmask.to(DEV)), want_seq))), 0.0)
    out["predict_seq_exact"] = (float(post.predict_seq(logits.to(DEV)) != [[letters[i] for i in row] for row in
                                                                            torch.argmax(logits, -1).tolist()]), 0.0)
    # ---- superposition (generate.py:176-195, tmalign.py:23-49) ----
    ref = 10.0 * torch.randn(N, 3, generator=g, dtype=torch.float64)
    mask = torch.ones(B, N)
    mask[:, :7] = 0  # ligand tokens do not take part
    mask[2, N - 30:] = 0
    pos = torch.empty(B, N, 3, dtype=torch.float64)
    for b in range(B):
        q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g, dtype=torch.float64))
        if torch.det(q) < 0:
            q[:, 0] = -q[:, 0]
        pos[b] = (ref + (0.5 + b) * torch.randn(N, 3, generator=g, dtype=torch.float64)) @ q + 5.0 * torch.randn(1, 3, generator=g, dtype=torch.float64)
    pos[3, :, 2] = -pos[3, :, 2]  # a mirror image: only the mirrored superposition can fit it
    tm, rmsd, t, R, mirrored = post.superpose(pos.float().to(DEV), ref.float().to(DEV), mask.to(DEV))
    torch.cuda.synchronize()
    worst_r = worst_R = worst_tm = worst_fit = 0.0
    for b in range(B):
        sel = mask[b] > 0.5
        P = pos[b][sel].numpy().copy()
        Q = ref[sel].numpy()
        if bool(mirrored[b]):
            P[:, 2] = -P[:, 2]
        Rw, tw, rw = _kabsch_np(P, Q)
        L = int(sel.sum())
        d0 = max(0.5, 1.24 * (L - 15) ** (1.0 / 3.0) - 1.8)
        d2 = ((tw + P @ Rw - Q) ** 2).sum(-1)
        tmw = float((1.0 / (1.0 + d2 / d0 ** 2)).sum() / L)
        Rg = R[b].double().cpu().numpy()
        if bool(mirrored[b]):
            Rg = np.diag([1.0, 1.0, -1.0]) @ Rg  # back to the rotation that acts on the mirrored sample
        worst_r = max(worst_r, abs(float(rmsd[b]) - rw) / rw)
        worst_R = max(worst_R, float(np.abs(Rg - Rw).max()))
        worst_tm = max(worst_tm, abs(float(tm[b]) - tmw))
        # the returned (t, R) act on the ORIGINAL sample: aligned = t + pos @ R (generate.py:186)
        al = t[b].double().cpu().numpy() + pos[b][sel].numpy() @ R[b].double().cpu().numpy()
        worst_fit = max(worst_fit, abs(float(np.sqrt(((al - Q) ** 2).sum(-1).mean())) - rw) / rw)
    out.update({"rmsd_rel": (worst_r, 1e-4), "rotation_maxabs": (worst_R, 1e-4), "tm_abs": (worst_tm, 1e-4),
                "aligned_rmsd_rel": (worst_fit, 1e-4), "mirror_detected": (0.0 if bool(mirrored[3]) and not bool(mirrored[0]) else 1.0, 0.0)})
    return out


def case_predict_loop(seed=52, T=4):
    """predict(model, dataloader) == Trainer.predict for generate.py:145-159: RepeatDataset + collate_fn + predict_step under
    no_grad; the result of every batch equals a direct predict_step call with the same generator state."""
    from torch.utils.data import DataLoader
    from protein_redesign_b200.predict import RepeatDataset, collate_fn, predict
    from protein_redesign_b200 import postprocess as post
    cfg = dataclasses.replace(syn.README, num_steps=T, mask_prob=0.3)
    m, sd = _model(cfg, seed)
    item = syn.make_complex(cfg, 6, 26, seed=seed)
    dl = DataLoader(RepeatDataset(item, 5), batch_size=2, collate_fn=collate_fn)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    res = predict(m, dl)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    direct = [m.predict_step({k: (v.to(DEV) if isinstance(v, torch.Tensor) else v) for k, v in b.items()}, i) for i, b in enumerate(dl)]
    torch.cuda.synchronize()
    same = all(torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) for a, b in zip(res, direct))
    shapes = [tuple(p.shape) for p, _ in res] == [(2, 32, 3), (2, 32, 3), (1, 32, 3)]
    finite = all(bool(torch.isfinite(p).all()) and bool(torch.isfinite(l).all()) for p, l in res)
    pos = torch.cat([p for p, _ in res])
    logits = torch.cat([l for _, l in res])
    rmask = torch.cat([torch.zeros(5, 6), torch.ones(5, 26)], 1).to(DEV)
    seqs = post.trimmed_sequence(logits, rmask)
    tm, rmsd, t, R, mir = post.superpose(pos, pos[0], rmask)  # generate.py:171-175: the first sample as the reference
    return {"equals_direct_predict_step": (0.0 if same else 1.0, 0.0), "shapes": (0.0 if shapes else 1.0, 0.0),
            "finite": (0.0 if finite else 1.0, 0.0), "sequence_lengths": (float(sum(len(s) > 26 for s in seqs)), 0.0),
            "self_tm_is_one": (abs(float(tm[0]) - 1.0), 1e-5), "self_rmsd_is_zero": (float(rmsd[0]), 1e-3)}


CASES.update({"postprocess": lambda: case_postprocess(), "predict_loop": lambda: case_predict_loop()})


def case_train_modes(cfg=None, sizes=((8, 32), (6, 27)), seed=60):
    """The three ways to run a training step give the same loss and gradients: training_step + loss.backward() (autograd),
    autograd.training_step_manual (no autograd engine), autograd.TrainStepGraph (device side as one CUDA graph) -- and the
    graph picks up an in-place weight update between replays (its fp16 weight packs are rebuilt inside the graph)."""
    from protein_redesign_b200 import autograd as ag
    cfg = cfg or dataclasses.replace(syn.README, mask_prob=0.15, num_steps=2000)
    m, sd = _model(cfg, seed)
    m.train()
    batch = syn.make_batch(cfg, list(sizes), seed=seed, with_positions=True)
    B, N = batch["atom_mask"].shape
    g = torch.Generator().manual_seed(seed + 1)
    draws = {"z": torch.randn(B, N, 3, generator=g).to(DEV), "seq": torch.randn(B, N, 21, generator=g).to(DEV)}
    names = [n for n, _ in ag.trainable_parameters(m)]

    def flat_of(model):
        return torch.cat([p.grad.reshape(-1) for n, p in model.named_parameters() if p.requires_grad]).clone()

    torch.manual_seed(seed)
    with torch.enable_grad():
        loss_a = m.training_step(_to_dev(batch), 0, noise=draws)
        loss_a.backward()
    ga = flat_of(m)
    grads = ag.FlatGrads(m).attach()
    torch.manual_seed(seed)
    loss_b = ag.training_step_manual(m, _to_dev(batch), grads, noise=draws)
    gb = grads.flat.clone()
    tsg = ag.TrainStepGraph(m, _to_dev(batch), grads, inject_noise=True)
    torch.manual_seed(seed)
    loss_c = tsg.step(_to_dev(batch), noise=draws).clone()
    gc = grads.flat.clone()
    # an optimiser-like in-place update of EVERY parameter, then the same step again: graph == manual on the new weights
    with torch.no_grad():
        for _, p in ag.trainable_parameters(m):
            p.mul_(1.02)
    torch.manual_seed(seed)
    loss_d = tsg.step(_to_dev(batch), noise=draws).clone()
    gd = grads.flat.clone()
    torch.manual_seed(seed)
    loss_e = ag.training_step_manual(m, _to_dev(batch), grads, noise=draws)
    ge = grads.flat.clone()
    torch.cuda.synchronize()
    return {"manual_vs_autograd_loss": (abs(float(loss_b) - float(loss_a)) / abs(float(loss_a)), 1e-6),
            "manual_vs_autograd_grads": (rel(gb, ga), 1e-5),
            "graph_vs_manual_loss": (abs(float(loss_c) - float(loss_b)) / abs(float(loss_b)), 1e-6),
            "graph_vs_manual_grads": (rel(gc, gb), 1e-5),
            "weights_changed_loss_moved": (0.0 if abs(float(loss_d) - float(loss_c)) / abs(float(loss_c)) > 1e-4 else 1.0, 0.0),
            "graph_after_update_vs_manual_loss": (abs(float(loss_d) - float(loss_e)) / abs(float(loss_e)), 1e-6),
            "graph_after_update_vs_manual_grads": (rel(gd, ge), 1e-5),
            "parameters": (float(len(names) != 240), 0.0)}


CASES["train_modes"] = lambda: case_train_modes()


def case_dw_tc():
    """The weight-gradient reduction dW = dY^T X on tcgen05 with both operands MN-major tf32 (csrc/prd_bwd_dw.cu) through
    its C-ABI hook, against float64 on operands pre-rounded to tf32 (every product is then exact): full tiles, column tails
    (Nout, K not multiples of the tile), row counts that are not multiples of the pipeline block, K > 256 (several tiles)."""
    import ctypes
    lib = _lib.load()
    lib.prd_dw_acc.restype = ctypes.c_int
    lib.prd_dw_acc.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                               ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]

    def r_tf32(x):
        i = x.view(torch.int32)
        return ((i + 0x1000) & ~0x1FFF).view(torch.float32)

    out = {}
    g = torch.Generator().manual_seed(3)
    for R, Nout, K in ((512, 128, 64), (4096, 256, 64), (4096, 64, 256), (1000, 320, 64), (3108, 128, 512), (2000, 100, 72)):
        dY, X = r_tf32(torch.randn(R, Nout, generator=g)), r_tf32(torch.randn(R, K, generator=g))
        dYd, Xd = dY.to(DEV), X.to(DEV)
        dW = torch.ones(Nout, K, device=DEV)  # the kernel accumulates
        db = torch.zeros(Nout, device=DEV)
        rc = lib.prd_dw_acc(dYd.data_ptr(), Nout, Xd.data_ptr(), K, R, Nout, K, dW.data_ptr(), K, db.data_ptr(), 0.5, 2, None)
        assert rc == 0, _lib.last_error()
        torch.cuda.synchronize()
        out[f"dW_{R}x{Nout}x{K}"] = (rel(dW, 1.0 + 0.5 * dY.double().t() @ X.double()), 2e-6)
        out[f"db_{R}x{Nout}x{K}"] = (rel(db, 0.5 * dY.double().sum(0)), 2e-6)
    return out


CASES["dw_tc"] = lambda: case_dw_tc()


def case_mask_contract(cfg=syn.PAPER, B=1, N=40, seed=61):
    """TriangleAttention.forward(pair, mask_2d) accepts m (x) m and refuses a general pair mask."""
    m, sd = _model(cfg, seed)
    _, pair, mask = _pair_inputs(cfg, B, N, seed, pad=3)
    m2 = (mask.unsqueeze(-1) * mask.unsqueeze(-2)).to(DEV)
    mod = m.Denoiser.folding_blocks[0].pair_attn_starting
    ok = mod(pair.to(DEV), m2)
    bad = m2.clone()
    bad[0, 1, 2] = 0.0
    refused = 0.0
    try:
        mod(pair.to(DEV), bad)
        refused = 1.0
    except ValueError:
        pass
    torch.cuda.synchronize()
    return {"finite": (0.0 if bool(torch.isfinite(ok).all()) else 1.0, 0.0), "general_mask_refused": (refused, 0.0)}


CASES["mask_contract"] = lambda: case_mask_contract()


def case_fused_bias(cfg=syn.PAPER, B=2, N=72, seed=62):
    """The pair-bias projections that no longer get their own pass over the pair tensor: ops.pair_bias (SPAttention's
    linear_z and the first block's attn_bias in one stream) and the next block's attn_bias out of pair_fc's epilogue."""
    m, sd = _model(cfg, seed)
    _, pair, _ = _pair_inputs(cfg, B, N, seed)
    den = m.Denoiser
    p_spa = "Denoiser.SPAAttnBlock."
    want_spa = F.linear(ref._ln(pair, sd[p_spa + "linear_z.0.weight"], sd[p_spa + "linear_z.0.bias"]),
                        sd[p_spa + "linear_z.1.weight"]).permute(0, 3, 1, 2)
    want_b0 = ref.attn_bias_from_pair(sd, _block_prefix(0), pair)
    got_spa, got_b0 = ops.pair_bias(cfg, pair.to(DEV).contiguous(), den.SPAAttnBlock.bias_projection(),
                                    den.folding_blocks[0].bias_projection())
    only_spa, none = ops.pair_bias(cfg, pair.to(DEV).contiguous(), den.SPAAttnBlock.bias_projection())
    updated = pair + ref.transition(sd, _block_prefix(0) + "pair_fc.", pair)
    want_b1 = ref.attn_bias_from_pair(sd, _block_prefix(1), updated)
    proj = den.folding_blocks[1].bias_projection()
    p = pair.to(DEV).contiguous()
    _, got_b1 = ops.pair_transition(cfg, p, den.folding_blocks[0].pair_fc.packed_pair(), p, next_bias=(proj[2], proj[3]))
    torch.cuda.synchronize()
    return {"spa_bias": (rel(got_spa, want_spa), 1e-5), "block0_bias": (rel(got_b0, want_b0), 1e-5),
            "single_projection": (rel(only_spa, want_spa) + (0.0 if none is None else 1.0), 1e-5),
            "pair_fc_rows": (rel(p, updated), OP_TOL), "next_block_bias": (rel(got_b1, want_b1), OP_TOL)}


CASES["fused_bias"] = lambda: case_fused_bias()
CASES["fused_bias_readme"] = lambda: case_fused_bias(syn.README, 2, 40)


# ------------------------------------------------------------------------------------------
# round 2, second half: the GEMM's TMA-store epilogue, the tensor-core backward attention, batched weight preparation
# ------------------------------------------------------------------------------------------
def case_gemm_tf32_epilogue(M=197, N=100, K=64, nb1=2, seed=3):
    """fp32 C through the tensor-map store (ldc % 4 == 0) on tiles that are ragged in M and N (boxes clipped by the TMA unit),
    with every epilogue operand: bias, [M, N] gate (ReLU backward) and [M, N] addend, tf32-rounded result."""
    g = torch.Generator().manual_seed(seed)

    def r_tf32(x):
        return ((x.view(torch.int32) + 0x1000) & ~0x1FFF).view(torch.float32)

    a = r_tf32(torch.randn(1, nb1, M, K, generator=g))
    b = r_tf32(torch.randn(1, nb1, N, K, generator=g))
    bias = torch.randn(N, generator=g)
    gate = torch.randn(1, nb1, M, N, generator=g)
    add = torch.randn(1, nb1, M, N, generator=g)
    prod = torch.matmul(a.double(), b.double().transpose(-1, -2)) + bias.double()
    want = r_tf32((torch.where(gate > 0, prod, torch.zeros_like(prod)) + add.double()).float())
    out = torch.full((1, nb1, M, N), float("nan"), dtype=torch.float32, device=DEV)
    _lib.gemm_f16(a.to(DEV).contiguous(), b.to(DEV).contiguous(), out, bias=bias.to(DEV), mul=gate.to(DEV).contiguous(), mul_step=True,
                  add=add.to(DEV).contiguous(), round_tf32=True)
    torch.cuda.synchronize()
    # a result within half a tf32 ulp of a rounding boundary may land on the other side: compare at tf32 resolution
    return {"rel": (rel(out, want), 2e-4), "finite": (0.0 if bool(torch.isfinite(out).all()) else 1.0, 0.0)}


def case_attn_tc_vs_simt(cfg=syn.PAPER, B=2, N=77, mode="ending", seed=61, pad=6):
    """The tensor-core backward attention (mma.sync tf32) against the exact fp32 SIMT kernels it replaced, on a length that
    is no multiple of any tile (77 = 2 x 32 + 13), with padded tokens: same op, PRD_ATTN_SIMT toggled."""
    import os
    from protein_redesign_b200 import autograd as ag
    prefix = _block_prefix() + f"pair_attn_{mode}."
    res = {}
    for tag, env in (("tc", "0"), ("simt", "1")):
        os.environ["PRD_ATTN_SIMT"] = env
        try:
            m, sd, P, G = _bwd_setup(cfg, seed, [prefix])
            _, pair, mask = _pair_inputs(cfg, B, N, seed, pad)
            d = _rand_like(pair, seed + 1).to(DEV).contiguous()
            ag.triangle_attention_bwd(cfg, P, G, prefix + "attn.", 1 if mode == "ending" else 0, pair.to(DEV).contiguous(), mask.to(DEV), d)
            torch.cuda.synchronize()
            res[tag] = (d.clone(), {k: v.clone() for k, v in G.items()})
        finally:
            os.environ.pop("PRD_ATTN_SIMT", None)
    out = {"dx": (rel(res["tc"][0], res["simt"][0]), 1e-3)}
    for k in res["tc"][1]:
        if k.startswith(prefix) and float(res["simt"][1][k].norm()) > 0:
            out["dW:" + k] = (rel(res["tc"][1][k], res["simt"][1][k]), 2e-3)
    return out


CASES["gemm_tf32_epilogue"] = lambda: case_gemm_tf32_epilogue()
CASES["gemm_tf32_epilogue_unaligned"] = lambda: case_gemm_tf32_epilogue(M=130, N=70, K=96, nb1=1, seed=4)   # ldc % 4 != 0: no tensor map
CASES["attn_tc_vs_simt"] = lambda: case_attn_tc_vs_simt()
CASES["attn_tc_vs_simt_starting"] = lambda: case_attn_tc_vs_simt(mode="starting", N=45, seed=62, pad=0)
