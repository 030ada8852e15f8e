"""Backward pass of the denoiser on the CUDA kernels (reference: Lightning's ``loss.backward()`` through
``ProteinReDiffModel.training_step`` model.py:528-549 with per-block checkpointing, modules.py:399-401).

Structure
* forward: the fused inference kernels, keeping only block-boundary checkpoints (single, pair before every FoldingBlock,
  the embedding outputs and the OuterProductUpdate operands);
* backward: per block, the forward of the block is re-run once to recover the input of each of its eight residual
  updates, then ``prd_<op>_bwd`` (include/prd_denoiser.h) is called in reverse order; each of those recomputes its own
  intermediates from its input, so nothing else is ever stored;
* parameter gradients are fp32 tensors shaped like the reference parameters, accumulated by the kernels and handed to
  autograd by :class:`DenoiserFunction` -- ``loss.backward()``, optimisers and gradient all-reduce work unchanged.

Nothing here computes in PyTorch: the functions pack pointers and call the C ABI.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib, ops
from .ops import make_dims


def _bwd(op, dims, ins, outs, weights):
    _lib.call_bwd(op, dims, ins, outs, weights)


# --------------------------------------------------------------------------------------------------------------
# per-op launchers.  P: name -> fp32 parameter (contiguous, CUDA), G: name -> fp32 gradient buffer (accumulated)
# --------------------------------------------------------------------------------------------------------------
def _names(prefix: str, *suffixes: str) -> List[str]:
    return [prefix + s for s in suffixes]


def transition_bwd(cfg, P, G, prefix: str, x_in: torch.Tensor, d_io: torch.Tensor) -> None:
    """single_fc / pair_fc (reference modules.py:306-311, 321-326); ``d_io``: d out -> d in."""
    n = _names(prefix, "1.weight", "1.bias", "3.weight", "3.bias")
    op = "single_transition" if x_in.dim() == 3 else "pair_transition"
    B, N = x_in.shape[:2]
    _bwd(op, make_dims(cfg, B, N), [x_in], [d_io] + [G[k] for k in n], [P[k] for k in n])


def seq_head_bwd(cfg, P, G, single: torch.Tensor, d_seq_pred: torch.Tensor, d_single: torch.Tensor) -> None:
    """seq_mlp (reference model.py:374); ``d_single`` is written."""
    n = ["seq_mlp.1.weight", "seq_mlp.1.bias", "seq_mlp.3.weight"]
    B, N = single.shape[:2]
    _bwd("seq_head", make_dims(cfg, B, N), [single, d_seq_pred], [d_single] + [G[k] for k in n], [P[k] for k in n])


def coord_head_bwd(cfg, P, G, pair, z, mask, d_noise_pred, d_pair) -> None:
    """weight_radial + equivariant sum + remove_mean + the trunk's symmetrisation (reference model.py:364-373,
    modules.py:403); ``pair`` is the tensor BEFORE symmetrisation, ``d_pair`` is written."""
    n = ["weight_radial.1.weight", "weight_radial.1.bias", "weight_radial.3.weight"]
    B, N = mask.shape
    _bwd("coord_head", make_dims(cfg, B, N), [pair, z, mask, d_noise_pred], [d_pair] + [G[k] for k in n], [P[k] for k in n])


_ATTN = ("q_proj.weight", "k_proj.weight", "v_proj.weight", "gate_proj.weight", "gate_proj.bias", "out_proj.weight", "out_proj.bias")


def triangle_attention_bwd(cfg, P, G, prefix: str, mode: int, pair_in, mask, d_pair) -> None:
    """reference modules.py:236-243 (``prefix`` ends with ``attn.``)."""
    n = _names(prefix, *_ATTN)
    B, N = mask.shape
    _bwd("triangle_attention", make_dims(cfg, B, N, mode=mode), [pair_in, mask], [d_pair] + [G[k] for k in n], [P[k] for k in n])


def single_attention_bwd(cfg, P, G, block_prefix: str, single_in, pair, mask, d_single, d_pair) -> None:
    """attn_bias + single_attn of a FoldingBlock (reference modules.py:300-304, 185-225, 335)."""
    n = _names(block_prefix, "attn_bias.1.weight", "attn_bias.1.bias") + _names(block_prefix + "single_attn.", *_ATTN)
    B, N = mask.shape
    _bwd("single_attention", make_dims(cfg, B, N), [single_in, pair, mask], [d_single, d_pair] + [G[k] for k in n], [P[k] for k in n])


def triangle_multiplication_bwd(cfg, P, G, prefix: str, mode: int, pair_in, mask, d_pair) -> None:
    """reference modules.py:262-274."""
    n = _names(prefix, "ab_proj.weight", "ab_proj.bias", "ab_gate.weight", "ab_gate.bias", "out_proj.weight", "out_proj.bias",
               "out_gate.weight", "out_gate.bias")
    B, N = mask.shape
    _bwd("triangle_multiplication", make_dims(cfg, B, N, mode=mode), [pair_in, mask], [d_pair] + [G[k] for k in n],
         [P[k] for k in n])


def outer_linear_bwd(cfg, P, G, prefix: str, single_in, d_pair, d_single) -> None:
    """reference modules.py:283-287; ``d_pair`` is read only (the residual passes through), ``d_single`` accumulates."""
    n = _names(prefix, "linear.weight", "linear.bias")
    B, N = single_in.shape[:2]
    _bwd("outer_linear", make_dims(cfg, B, N), [single_in, d_pair], [d_single, G[n[0]], G[n[1]]], [P[n[0]]])


_SPA = ("layer_norm_m.weight", "layer_norm_m.bias", "linear_z.0.weight", "linear_z.0.bias", "linear_z.1.weight",
        "mha.linear_q.weight", "mha.linear_k.weight", "mha.linear_v.weight", "mha.linear_g.weight", "mha.linear_g.bias",
        "mha.linear_o.weight", "mha.linear_o.bias")


def spattention_bwd(cfg, P, G, single_in, pair, d_single, d_pair, prefix: str = "Denoiser.SPAAttnBlock.") -> None:
    """reference AF2_modules.py:421-473; ``d_single``: d out -> d in, ``d_pair`` accumulates."""
    n = _names(prefix, *_SPA)
    B, N = single_in.shape[:2]
    _bwd("spattention", make_dims(cfg, B, N), [single_in, pair], [d_single, d_pair] + [G[k] for k in n], [P[k] for k in n])


def opm_project_bwd(cfg, P, G, single_in, mask, d_a, d_b, d_single, prefix: str = "Denoiser.opm.") -> None:
    """reference AF2_modules.py:519-530; ``d_single`` accumulates."""
    n = _names(prefix, "layer_norm.weight", "layer_norm.bias", "linear_1.weight", "linear_1.bias", "linear_2.weight", "linear_2.bias")
    B, N = mask.shape
    _bwd("opm_project", make_dims(cfg, B, N), [single_in, mask, d_a, d_b], [d_single] + [G[k] for k in n], [P[k] for k in n])


def pair_embed_bwd(cfg, P, G, batch, d_pair, z, mask, t, opm_a, opm_b) -> Tuple[torch.Tensor, torch.Tensor]:
    """Everything that wrote the initial pair tensor (reference model.py:348-361, AF2_modules.py:532-543,
    modules.py:395-397).  Returns (d_opm_a, d_opm_b)."""
    B, N = mask.shape
    d_a, d_b = torch.empty_like(opm_a), torch.empty_like(opm_b)
    outs = [d_a, d_b, G["Denoiser.opm.linear_out.weight"], G["Denoiser.opm.linear_out.bias"], G["embed_dist.1.weight"],
            G["embed_beta.1.weight"], G["embed_bond_feats.embeddings.0.weight"], G["embed_bond_feats.embeddings.1.weight"],
            G["embed_bond_feats.embeddings.2.weight"], G["embed_bond_distance.weight"], G["embed_relpos.weight"]]
    ins = [d_pair, z, mask, t, opm_a, opm_b, batch["atom_mask"], batch["residue_mask"], batch["bond_mask"], batch["bond_feats"],
           batch["bond_distance"], batch["residue_index"], batch["residue_chain_index"]]
    _bwd("pair_embed", make_dims(cfg, B, N), [x.contiguous() for x in ins], outs,
         [P["Denoiser.opm.linear_out.weight"], P["embed_dist.0.center"], P["embed_beta.0.weight"]])
    return d_a, d_b


def single_embed_bwd(cfg, P, G, batch, seq_t, d_single) -> None:
    """reference model.py:342-346 (+ :99-102)."""
    B, N = seq_t.shape[:2]
    outs = [G[f"embed_atom_feats.embeddings.{f}.weight"] for f in range(9)] + [G["embed_residue_type.1.weight"],
                                                                               G["embed_residue_esm.1.weight"]]
    ins = [d_single, batch["atom_feats"], batch["atom_mask"], batch["residue_mask"], seq_t, batch["residue_esm"]]
    _bwd("single_embed", make_dims(cfg, B, N), [x.contiguous() for x in ins], outs, [P["embed_residue_type.1.weight"]])


# --------------------------------------------------------------------------------------------------------------
# the whole network
# --------------------------------------------------------------------------------------------------------------
class DenoiserTape:
    """Checkpoints of one forward evaluation (what the reference's per-block ``checkpoint`` keeps)."""

    __slots__ = ("batch", "z", "seq_t", "mask", "t", "single0", "opm_a", "opm_b", "pair1", "blocks", "single_out", "pair_out", "per_op")


_TAPE_OPS_MAX_BYTES = 32 << 30


def _tape_per_op(cfg, pair: torch.Tensor) -> bool:
    """Keep the input of every residual update (True) or one checkpoint per block and re-run it in the backward pass."""
    import os
    mode = os.environ.get("PRD_TAPE", "")
    if mode in ("ops", "blocks"):
        return mode == "ops"
    return 6 * cfg.num_blocks * pair.numel() * pair.element_size() <= _TAPE_OPS_MAX_BYTES


def _block_forward_taped(cfg, block, single: torch.Tensor, pair: torch.Tensor, mask: torch.Tensor, all_valid: bool):
    """One FoldingBlock in place on (single, pair), op by op in the reference order (modules.py:335-342), returning the input
    of each residual update: (s0, p0) single_attn [+ pair bias], s1 single_fc, s2 outer_linear, p1 pair_mul_outgoing,
    p2 pair_mul_incoming, p3 pair_attn_starting, p4 pair_attn_ending, p5 pair_fc."""
    s0, p0 = single.clone(), pair.clone()
    ops.single_attention(cfg, single, p0, mask, block.single_attn.packed_single(block.attn_bias[1]), single)
    s1 = single.clone()
    ops.single_transition(cfg, single, block.single_fc.packed_single(), single)
    s2 = single.clone()
    block.outer_linear.apply_(cfg, s2, pair)
    p1 = pair.clone()
    block.pair_mul_outgoing.apply_(cfg, pair, mask)
    p2 = pair.clone()
    block.pair_mul_incoming.apply_(cfg, pair, mask)
    p3 = pair.clone()
    block.pair_attn_starting.apply_(cfg, pair, mask, all_valid=all_valid)
    p4 = pair.clone()
    block.pair_attn_ending.apply_(cfg, pair, mask, all_valid=all_valid)
    p5 = pair.clone()
    ops.pair_transition(cfg, pair, block.pair_fc.packed_pair(), pair)
    return s0, p0, s1, s2, p1, p2, p3, p4, p5


def forward_with_checkpoints(model, batch, z, seq_t, mask, t) -> Tuple[torch.Tensor, torch.Tensor, DenoiserTape]:
    """The same kernels as ``ProteinReDiffModel._denoise`` (reference model.py:318-375), keeping the block-boundary
    activations."""
    cfg, w = model.cfg, model._weights()
    den = model.Denoiser
    tape = DenoiserTape()
    tape.batch, tape.z, tape.seq_t, tape.mask, tape.t = batch, z.contiguous(), seq_t.contiguous(), mask.contiguous(), t.contiguous()
    esm_emb, pair_static = model._static_embeddings(batch)
    single = ops.single_embed(cfg, batch["atom_feats"].contiguous(), batch["atom_mask"].contiguous(),
                              batch["residue_mask"].contiguous(), tape.seq_t, esm_emb, w["atom_tabs"], w["w_type"])
    tape.single0 = single.clone()
    tape.opm_a, tape.opm_b = den.opm.project(cfg, single, tape.mask)
    _, (w_o, b_o) = den.opm.packed_weights()
    pair = torch.empty(mask.shape[0], mask.shape[1], mask.shape[1], cfg.pair_dim, dtype=torch.float32, device=z.device)
    ops.pair_embed(cfg, pair_static, tape.z, tape.mask, tape.t, tape.opm_a, tape.opm_b, w["pair_dyn"] + [w_o, b_o], pair,
                   rbf_lut=w["rbf_lut"])
    tape.pair1 = pair.clone()
    den.SPAAttnBlock(single, pair, tape.mask, cfg=cfg, out=single)
    tape.blocks = []
    all_valid = bool(batch.get("_all_valid", False))
    # Reference behaviour = one checkpoint per block and a re-run of the block in the backward pass (modules.py:399-401): that
    # trades a second forward for memory on 16 - 80 GB devices.  With 180 GB the inputs of all eight residual updates of every
    # block are simply kept (6 pair tensors per block: 1.2 GB at B = 2, N = 314) and the re-run disappears -- unless they would
    # not fit (_TAPE_OPS_MAX_BYTES) or PRD_TAPE=blocks asks for the reference scheme.
    tape.per_op = _tape_per_op(cfg, pair)
    for block in den.folding_blocks:
        if tape.per_op:
            tape.blocks.append(_block_forward_taped(cfg, block, single, pair, tape.mask, all_valid))
        else:
            tape.blocks.append((single.clone(), pair.clone()))
            block.forward_(cfg, single, pair, tape.mask, all_valid=all_valid)
    tape.single_out, tape.pair_out = single, pair
    noise_pred = ops.coord_head(cfg, pair, tape.z, tape.mask, w["coord"])
    seq_pred = ops.seq_head(cfg, single, w["seq"])
    return noise_pred, seq_pred, tape


def backward_from_checkpoints(model, tape: DenoiserTape, d_noise_pred: torch.Tensor, d_seq_pred: torch.Tensor,
                              P: Dict[str, torch.Tensor], G: Dict[str, torch.Tensor]) -> None:
    """Accumulates d loss / d parameter into ``G`` for every trainable parameter of the model."""
    cfg, den, mask = model.cfg, model.Denoiser, tape.mask
    d_single = torch.empty_like(tape.single_out)
    d_pair = torch.empty_like(tape.pair_out)
    seq_head_bwd(cfg, P, G, tape.single_out, d_seq_pred.contiguous(), d_single)
    coord_head_bwd(cfg, P, G, tape.pair_out, tape.z, mask, d_noise_pred.contiguous(), d_pair)
    tape.single_out = tape.pair_out = None
    for k in reversed(range(len(den.folding_blocks))):
        block, bp = den.folding_blocks[k], f"Denoiser.folding_blocks.{k}."
        all_valid = bool(tape.batch.get("_all_valid", False))
        if tape.per_op:
            s0, p0, s1, s2, p1, p2, p3, p4, p5 = tape.blocks.pop()
        else:
            # re-run the block once, keeping the input of each residual update (reference order modules.py:335-342)
            s0, p0 = tape.blocks.pop()
            s0, p0, s1, s2, p1, p2, p3, p4, p5 = _block_forward_taped(cfg, block, s0.clone(), p0.clone(), mask, all_valid)
        transition_bwd(cfg, P, G, bp + "pair_fc.", p5, d_pair)
        del p5
        triangle_attention_bwd(cfg, P, G, bp + "pair_attn_ending.attn.", 1, p4, mask, d_pair)
        del p4
        triangle_attention_bwd(cfg, P, G, bp + "pair_attn_starting.attn.", 0, p3, mask, d_pair)
        del p3
        triangle_multiplication_bwd(cfg, P, G, bp + "pair_mul_incoming.", 1, p2, mask, d_pair)
        del p2
        triangle_multiplication_bwd(cfg, P, G, bp + "pair_mul_outgoing.", 0, p1, mask, d_pair)
        del p1
        outer_linear_bwd(cfg, P, G, bp + "outer_linear.", s2, d_pair, d_single)
        transition_bwd(cfg, P, G, bp + "single_fc.", s1, d_single)
        single_attention_bwd(cfg, P, G, bp, s0, p0, mask, d_single, d_pair)
    spattention_bwd(cfg, P, G, tape.single0, tape.pair1, d_single, d_pair)
    d_a, d_b = pair_embed_bwd(cfg, P, G, tape.batch, d_pair, tape.z, mask, tape.t, tape.opm_a, tape.opm_b)
    opm_project_bwd(cfg, P, G, tape.single0, mask, d_a, d_b, d_single)
    single_embed_bwd(cfg, P, G, tape.batch, tape.seq_t, d_single)


def trainable_parameters(model) -> List[Tuple[str, torch.nn.Parameter]]:
    return [(n, p) for n, p in model.named_parameters() if p.requires_grad]


class DenoiserFunction(torch.autograd.Function):
    """(noise_pred, seq_pred) = network(z, seq_t, t | batch) as one autograd node over every trainable parameter."""

    @staticmethod
    def forward(ctx, model, batch, z, seq_t, mask, t, *params):
        with torch.no_grad():
            noise_pred, seq_pred, tape = forward_with_checkpoints(model, batch, z, seq_t, mask, t)
        ctx.model, ctx.tape = model, tape
        return noise_pred, seq_pred

    @staticmethod
    def backward(ctx, d_noise_pred, d_seq_pred):
        model, tape = ctx.model, ctx.tape
        named = trainable_parameters(model)
        P = {n: p.detach().contiguous() for n, p in model.named_parameters()}
        flat = torch.zeros(sum(p.numel() for _, p in named), dtype=torch.float32, device=tape.z.device)
        G, off = {}, 0
        for n, p in named:
            G[n] = flat[off:off + p.numel()].view(p.shape)
            off += p.numel()
        with torch.no_grad():
            backward_from_checkpoints(model, tape, d_noise_pred, d_seq_pred, P, G)
        ctx.tape = None
        return (None,) * 6 + tuple(G[n] for n, _ in named)


class LossFunction(torch.autograd.Function):
    """loss = mean(diff_loss / num_nodes) (reference model.py:499-526, 538-541) with the gradient the loss kernels
    already produce (d loss / d noise_pred, d loss / d seq_pred)."""

    @staticmethod
    def forward(ctx, noise_pred, seq_pred, cfg, noise_z, noise_seq, seq_t1, mask, residue_mask, residue_type, t, sched, detail):
        loss, diff, terms, d_noise, d_seq = ops.diffusion_loss(cfg, noise_pred.contiguous(), seq_pred.contiguous(), noise_z,
                                                               noise_seq, seq_t1, mask, residue_mask, residue_type, t, sched,
                                                               want_grads=True)
        ctx.save_for_backward(d_noise, d_seq)
        if detail is not None:
            detail.update(loss=loss, diff_loss=diff, terms=terms, d_noise_pred=d_noise, d_seq_pred=d_seq)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        d_noise, d_seq = ctx.saved_tensors
        return (g * d_noise, g * d_seq) + (None,) * 10


# --------------------------------------------------------------------------------------------------------------
# training step without the autograd engine (Lightning's "manual optimization" shape): loss + gradients straight into one
# flat fp32 buffer -- the bucket a data-parallel all-reduce sends -- and, optionally, the whole device side of the step
# captured as ONE CUDA graph
# --------------------------------------------------------------------------------------------------------------
class FlatGrads:
    """One contiguous fp32 buffer holding the gradient of every trainable parameter (in ``named_parameters`` order) with
    a per-parameter view; ``attach()`` makes the views the parameters' ``.grad`` so optimisers see them."""

    def __init__(self, model):
        self.named = trainable_parameters(model)
        dev = self.named[0][1].device
        self.flat = torch.zeros(sum(p.numel() for _, p in self.named), dtype=torch.float32, device=dev)
        self.views, off = {}, 0
        for n, p in self.named:
            self.views[n] = self.flat[off:off + p.numel()].view(p.shape)
            off += p.numel()

    def attach(self):
        for n, p in self.named:
            p.grad = self.views[n]
        return self


def _device_step(model, batch, x, mask, t, grads: FlatGrads, noise=None, detail=None):
    """Everything of training_step that runs on the device (reference model.py:490-526, 538-541 + backward): noise draws,
    q(), the network with checkpoints, the loss kernels, the backward kernels.  Fills ``grads`` and returns loss [1]."""
    from .synthetic import NUM_RESIDUE_CLASSES  # noqa: F401
    cfg = model.cfg
    seq, residue_mask = batch["residue_one_hot"].to(torch.float32).contiguous(), batch["residue_mask"].contiguous()
    noise_z = (noise["z"].to(x.device, torch.float32).clone() if noise is not None else torch.randn_like(x)).contiguous()
    noise_seq = (noise["seq"].to(x.device, torch.float32).clone() if noise is not None else torch.randn_like(seq)).contiguous()
    ops.remove_mean(cfg, noise_z, mask)
    ops.remove_mean(cfg, noise_seq, residue_mask)
    z_t, seq_t, seq_t1, _ = model.q(x, seq, t, noise_z, noise_seq, batch)
    noise_pred, seq_pred, tape = forward_with_checkpoints(model, batch, z_t, seq_t, mask, t)
    loss, diff, terms, d_noise, d_seq = ops.diffusion_loss(cfg, noise_pred, seq_pred, noise_z, noise_seq, seq_t1, mask, residue_mask,
                                                           batch["residue_type"].contiguous(), t, model._sched_table(),
                                                           want_grads=True)
    grads.flat.zero_()
    P = {n: p.detach().contiguous() for n, p in model.named_parameters()}
    backward_from_checkpoints(model, tape, d_noise, d_seq, P, grads.views)
    if detail is not None:
        detail.update(loss=loss, diff_loss=diff, terms=terms, noise_pred=noise_pred, seq_pred=seq_pred, t=t)
    return loss


def training_step_manual(model, batch, grads: FlatGrads, batch_idx: int = 0, noise=None, detail=None) -> torch.Tensor:
    """``training_step`` + ``loss.backward()`` in one call without the autograd engine: same host-side draws (prepare_batch's
    CPU randperm, ``t ~ randint``) and the same kernels as the autograd path, gradients written straight into ``grads``
    (no per-parameter accumulate kernels).  Returns the loss (0-d tensor)."""
    if not model.setup_schedule:
        model.run_setup_schedule()
        model.setup_schedule = True
    with torch.no_grad():
        batch = model.prepare_batch(batch, batch_idx)
        x, mask = batch["x"].contiguous(), batch["residue_and_atom_mask"].contiguous()
        t = torch.randint(0, model.num_steps, size=(x.size(0),)).to(x.device)
        return _device_step(model, batch, x, mask, t, grads, noise, detail).reshape(())


class TrainStepGraph:
    """The device side of one training step (noise draws, q(), network forward with checkpoints, loss, full backward into
    the flat gradient buffer) captured as ONE CUDA graph for a fixed (B, N): ~1000 kernel launches replayed without any
    host work in between.  Per step the host still does what the reference does on the host -- prepare_batch's CPU randperm
    and ``t ~ randint`` (model.py:424-468, 534) -- and copies the results into the graph's static input buffers."""

    _KEYS = ("atom_feats", "atom_mask", "residue_mask", "bond_mask", "bond_feats", "bond_distance", "residue_index",
             "residue_chain_index", "residue_type", "residue_esm", "residue_one_hot", "residue_extra_mask",
             "residue_inv_extra_mask", "x", "residue_and_atom_mask")

    def __init__(self, model, example_batch, grads: FlatGrads, inject_noise: bool = False):
        """``inject_noise``: the graph reads its noise draws from static buffers (``step(..., noise=...)`` fills them) instead
        of drawing them from the CUDA generator -- for parity tests."""
        self.model, self.grads = model, grads
        if not model.setup_schedule:
            model.run_setup_schedule()
            model.setup_schedule = True
        dev = grads.flat.device
        with torch.no_grad():
            prepared = model.prepare_batch(dict(example_batch))
            self.static = {k: prepared[k].contiguous().clone() for k in self._KEYS}
            B, N = self.static["atom_mask"].shape
            self.shape = (B, N)
            self.t = torch.zeros(B, dtype=torch.int64, device=dev)
            self.noise = {"z": torch.zeros(B, N, 3, device=dev), "seq": torch.zeros(B, N, 21, device=dev)} if inject_noise else None
            self.prep = torch.cuda.Stream(device=dev)  # host -> device copies and prepare_batch of the NEXT step
            # every workspace (forward and backward ops) sized before capture, on the capture stream
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            from ._packing import repack_in_place
            with torch.cuda.stream(side), repack_in_place():
                self.ws = _reserve_all(model.cfg, B, N, dev)
                for _ in range(2):  # warm-up on the capture stream (lazy kernel attributes, packed weights, allocator pools)
                    model._static_key = None
                    self.loss = _device_step(model, self.static, self.static["x"], self.static["residue_and_atom_mask"], self.t, grads, self.noise)
                torch.cuda.synchronize(dev)
                self.graph = torch.cuda.CUDAGraph()
                lib = _lib.load()
                c0 = int(lib.prd_launch_count())
                with torch.cuda.graph(self.graph, stream=side):
                    # inside the graph: the fp16 weight packs are rebuilt IN PLACE from the current fp32 parameters (an
                    # optimiser step between replays is picked up) and the step-invariant embeddings are recomputed from
                    # the static batch buffers (their content changes every step)
                    model._static_key = None
                    self.loss = _device_step(model, self.static, self.static["x"], self.static["residue_and_atom_mask"], self.t, grads, self.noise)
                model._static_key = None
                #: kernels of libprd_sm100.so inside the captured step (what one replay launches)
                self.launches_per_step = int(lib.prd_launch_count()) - c0
            torch.cuda.current_stream(dev).wait_stream(side)

    def step(self, batch, batch_idx: int = 0, noise=None) -> torch.Tensor:
        """prepare_batch + t on the host path, then one graph replay; returns the loss [1] (a static tensor of the graph).

        ``batch`` may live on the host (pinned): it is copied to the device here.  The copy and prepare_batch run on a side
        stream: prepare_batch reads masks back to the host (the reference's masking module calls ``.item()``), and on the
        main stream that read would wait for the previous step's whole graph -- with the side stream the host prepares
        step k + 1 while the device runs step k.  The copies into the graph's static buffers are ordered after the previous
        replay on the main stream."""
        model = self.model
        dev = self.grads.flat.device
        main = torch.cuda.current_stream(dev)
        with torch.no_grad():
            if any(isinstance(v, torch.Tensor) and v.is_cuda for v in batch.values()):
                self.prep.wait_stream(main)  # device inputs were produced on the caller's stream (no overlap then)
            with torch.cuda.stream(self.prep):
                batch = {k: (v.to(dev, non_blocking=True) if isinstance(v, torch.Tensor) and v.device != dev else v)
                         for k, v in batch.items()}
                prepared = model.prepare_batch(batch, batch_idx)
                t_dev = torch.randint(0, model.num_steps, size=(self.shape[0],)).to(dev, non_blocking=True)
            if tuple(prepared["atom_mask"].shape) != self.shape:
                raise ValueError(f"TrainStepGraph was captured for (B, N) = {self.shape}, got {tuple(prepared['atom_mask'].shape)}")
            main.wait_stream(self.prep)
            for k in self._KEYS:
                self.static[k].copy_(prepared[k], non_blocking=True)
                prepared[k].record_stream(main)  # allocated on the side stream, read on the main one
            self.t.copy_(t_dev, non_blocking=True)
            t_dev.record_stream(main)
            if self.noise is not None:
                if noise is None:
                    raise ValueError("this graph was captured with inject_noise=True: pass noise={'z': ..., 'seq': ...}")
                self.noise["z"].copy_(noise["z"], non_blocking=True)
                self.noise["seq"].copy_(noise["seq"], non_blocking=True)
            self.graph.replay()
        return self.loss


def _reserve_all(cfg, B: int, N: int, device):
    """Workspace of the current stream sized for every forward AND backward op at (B, N)."""
    import ctypes
    lib = _lib.load()
    d = make_dims(cfg, B, N)
    need = max([_lib.workspace_bytes(op, d) for op in _lib.OPS] +
               [int(getattr(lib, f"prd_{op}_bwd_workspace_bytes")(ctypes.byref(d))) for op in _lib.OPS_BWD])
    return _lib.Workspace.reserve(torch.device(device), need)
