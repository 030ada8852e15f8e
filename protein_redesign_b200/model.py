"""B200-native mirror of the reference's ``ProteinReDiff/model.py`` (ProteinReDiffModel).

Same constructor argument (Namespace / Mapping), sub-module and parameter names, and the call
signatures ``forward / sample_step (batch, z, seq_t, mask, t)``, ``sample(batch)``,
``predict_step(batch, batch_idx)``, ``prepare_batch``, ``run_setup_schedule`` (reference
model.py:254,318,378,249,424,172).  The network evaluation is enqueued on hand-written sm_100a
kernels through libprd_sm100 (include/prd_denoiser.h); ``sample`` captures one reverse-diffusion
step as a CUDA graph and replays it ``num_steps`` times with the schedule, the noise and the
step counter resident on the device.

``q``, ``diffusion_loss``, ``validation_step`` and ``training_step`` (reference model.py:471-549, 226-247) run on the loss
kernels (csrc/prd_loss.cu); with autograd enabled ``training_step`` returns a loss whose backward runs on the backward
kernels (autograd.py, csrc/prd_bwd*.cu).  Not in this build: ESM loading (training_mode masking).
"""
from __future__ import annotations

import contextlib
from argparse import ArgumentParser, Namespace
from typing import Dict, Mapping, Optional, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from . import ops
from ._packing import (PackCache, f32, half, invalidate_packed_weights, split_k, split_rows, tensor_version,
                       weights_epoch)
from .modules import AtomEmbedding, BondEmbedding, Denoiser, Linear, RadialBasisProjection, SinusoidalProjection
from .synthetic import NUM_RESIDUE_CLASSES, DenoiserConfig

import os as _os

_USE_RBF_LUT = _os.environ.get("PRD_RBF_LUT", "1") != "0"
_STEP_GRAPH = _os.environ.get("PRD_STEP_GRAPH", "1") != "0"

try:  # the reference derives from LightningModule; keep that when Lightning is installed
    import pytorch_lightning as _pl

    _Base = _pl.LightningModule
except Exception:  # pragma: no cover - Lightning is absent in the build image
    _Base = nn.Module


# --------------------------------------------------------------------------------------------
# host-side helpers restated from the reference (index / schedule path: bit exact)
# --------------------------------------------------------------------------------------------
def get_betas(n_timestep: int, schedule: str) -> torch.Tensor:
    """reference difffusion.py:8-26."""
    if schedule == "linear":
        return torch.linspace(0.0001, 0.02, n_timestep)
    if schedule == "cosine":
        steps = n_timestep + 1
        x = torch.linspace(0, n_timestep, steps)
        ac = torch.cos((x / steps) * torch.pi * 0.5) ** 2
        ac = ac / ac[0]
        return torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    raise ValueError(f"Invalid schedule: {schedule}")


class RandomMaskingModule(nn.Module):
    """reference mask_utils.py:72-112 (CPU ``torch.randperm`` over all valid residues of the batch)."""

    def forward(self, residue_mask, max_p, inverse_mask=False, stochastic=True):
        if stochastic:
            max_p = np.random.rand() * max_p
        ones = residue_mask == 1
        num_ones = int(ones.sum().item())
        rows, cols = torch.where(ones)
        pick = torch.randperm(num_ones)[: int(num_ones * max_p)].to(rows.device)
        self.residue_rand_mask = residue_mask.clone()
        self.residue_rand_mask[rows[pick], cols[pick]] = 0
        self.residue_rand_mask_esm = 1 - residue_mask.detach().clone()
        self.residue_rand_mask_esm[rows[pick], cols[pick]] = 32
        if inverse_mask:
            self.residue_inv_rand_mask = torch.zeros_like(residue_mask)
            self.residue_inv_rand_mask[rows[pick], cols[pick]] = 1
            return self.residue_rand_mask, self.residue_inv_rand_mask, self.residue_rand_mask_esm
        return self.residue_rand_mask, self.residue_rand_mask_esm


class _NullEMA:
    """Stand-in for torch_ema.ExponentialMovingAverage when that package is absent: no shadow weights."""

    def __init__(self, params, decay):
        self.decay = decay

    def to(self, *a, **k):
        return self

    def update(self, *a, **k):
        pass

    def state_dict(self):
        return {}

    def load_state_dict(self, *a, **k):
        pass

    @contextlib.contextmanager
    def average_parameters(self):
        yield


try:
    from torch_ema import ExponentialMovingAverage as _EMA
except Exception:  # pragma: no cover
    _EMA = _NullEMA


class ProteinReDiffModel(_Base):
    def __init__(self, args: Union[Namespace, Mapping, DenoiserConfig]):
        super().__init__()
        if isinstance(args, DenoiserConfig):
            args = args.to_namespace()
        if isinstance(args, Mapping):
            args = Namespace(**args)
        self.cfg = DenoiserConfig.from_args(args)
        self.pair_dim = args.pair_dim
        self.single_dim = args.single_dim
        self.dist_dim = args.dist_dim
        self.time_dim = args.time_dim
        self.max_bond_distance = args.max_bond_distance
        self.max_relpos = args.max_relpos
        self.esm_dim = args.esm_dim
        self.setup_schedule = False
        self.setup_esm = False
        self.mask_prob = args.mask_prob
        self.num_steps = args.num_steps
        self.diffusion_schedule = args.diffusion_schedule
        self.learning_rate = args.learning_rate
        self.warmup_steps = args.warmup_steps
        self.ema_decay = args.ema_decay
        self.n_recycles = args.n_recycles
        self.training_mode = args.training_mode

        self.RandomMaskingBlock = RandomMaskingModule()
        self.Denoiser = Denoiser(args)
        self.embed_atom_feats = AtomEmbedding(self.single_dim)
        self.embed_beta = nn.Sequential(SinusoidalProjection(self.time_dim),
                                        Linear(self.time_dim, self.pair_dim, bias=False, init="normal"))
        self.embed_residue_type = nn.Sequential(
            nn.LayerNorm(NUM_RESIDUE_CLASSES, elementwise_affine=False),
            Linear(NUM_RESIDUE_CLASSES, self.single_dim, bias=False, init="normal"), nn.ReLU())
        self.embed_bond_feats = BondEmbedding(self.pair_dim)
        self.embed_bond_distance = nn.Embedding(self.max_bond_distance + 1, self.pair_dim)
        self.embed_residue_esm = nn.Sequential(nn.LayerNorm(self.esm_dim, elementwise_affine=False),
                                               Linear(self.esm_dim, self.single_dim, bias=False, init="normal"))
        self.embed_relpos = nn.Embedding(self.max_relpos * 2 + 1, self.pair_dim)
        self.embed_dist = nn.Sequential(RadialBasisProjection(self.dist_dim),
                                        Linear(self.dist_dim, self.pair_dim, bias=False, init="normal"))
        self.weight_radial = nn.Sequential(nn.LayerNorm(self.pair_dim, elementwise_affine=False),
                                           Linear(self.pair_dim, self.pair_dim, init="relu"), nn.ReLU(),
                                           Linear(self.pair_dim, 1, bias=False, init="final"))
        self.seq_mlp = nn.Sequential(nn.LayerNorm(self.single_dim, elementwise_affine=False),
                                     Linear(self.single_dim, self.single_dim, init="relu"), nn.ReLU(),
                                     Linear(self.single_dim, NUM_RESIDUE_CLASSES, bias=False, init="final"))
        self.ema = _EMA(self.parameters(), decay=self.ema_decay)
        if hasattr(self, "save_hyperparameters"):
            self.save_hyperparameters(args)
        self._pack = PackCache()
        self._static_key = None
        self._static = None
        self._step_graph = None

    # ---- argparse surface (reference model.py:129-170) ---------------------------------------
    @staticmethod
    def add_argparse_args(parent_parser: ArgumentParser) -> ArgumentParser:
        p = parent_parser.add_argument_group("DiffusionModel")
        p.add_argument("--training_mode", action="store_true")
        p.add_argument("--mask_prob", type=float, default=1.0)
        for name, default in (("esm_dim", 1280), ("time_dim", 256), ("dist_dim", 256), ("single_dim", 512),
                              ("pair_dim", 64), ("head_dim", 16), ("num_heads", 4), ("transition_factor", 4),
                              ("num_blocks", 12), ("max_bond_distance", 7), ("max_relpos", 32), ("num_steps", 64),
                              ("warmup_steps", 1000)):
            p.add_argument(f"--{name}", type=int, default=default)
        p.add_argument("--diffusion_schedule", type=str, default="linear")
        p.add_argument("--learning_rate", type=float, default=4e-4)
        p.add_argument("--ema_decay", type=float, default=0.999)
        q = parent_parser.add_argument_group("IterativeDenoiser")
        q.add_argument("--n_recycles", type=int, default=4)
        return parent_parser

    @property
    def _device(self):
        return next(self.parameters()).device

    # ---- schedule (reference model.py:172-190) -----------------------------------------------
    def run_setup_schedule(self):
        dev = self._device
        self.betas = get_betas(self.num_steps, self.diffusion_schedule).to(dev)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, 0)
        self.alphas_cumprod_prev = torch.cat([torch.ones(1, device=dev), self.alphas_cumprod[:-1]])
        self.one_minus_alphas_cumprod = 1.0 - self.alphas_cumprod
        self.sqrt_betas = torch.sqrt(self.betas)
        self.sqrt_alphas = torch.sqrt(self.alphas)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_one_minus_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)
        # per-step coefficients read by the on-device sampler update (model.py:407-412,419)
        self._coef = torch.stack([1.0 / self.sqrt_alphas,
                                  (1.0 - self.alphas) / self.sqrt_one_minus_alphas_cumprod,
                                  self.sqrt_betas], dim=1).contiguous()

    # ---- packed weights -----------------------------------------------------------------------
    def _weights(self):
        srcs = [self.embed_residue_esm[1].weight, self.embed_residue_type[1].weight, self.embed_beta[0].weight,
                self.embed_beta[1].weight, self.embed_dist[0].center, self.embed_dist[1].weight,
                self.embed_bond_distance.weight, self.embed_relpos.weight, self.weight_radial[1].weight,
                self.weight_radial[1].bias, self.weight_radial[3].weight, self.seq_mlp[1].weight,
                self.seq_mlp[1].bias, self.seq_mlp[3].weight]
        srcs += [e.weight for e in self.embed_atom_feats.embeddings]
        srcs += [e.weight for e in self.embed_bond_feats.embeddings]

        def build():
            if getattr(self, "_rbf_dmax", None) is None and srcs[4].is_cuda:
                self._rbf_dmax = float(srcs[4].max()) + 0.52  # centres are a fixed buffer: read once (host sync)
            w = {
                "esm": [split_k(srcs[0])],
                "w_type": f32(srcs[1]),
                "pair_dyn": [f32(srcs[2]), f32(srcs[3]), split_rows(srcs[5]), f32(srcs[4])],
                # d -> W_dist rbf(d) tabulated once per weight version (PRD_RBF_LUT=0 keeps the per-pair RBF GEMM)
                "rbf_lut": ops.rbf_lut_build(self.cfg, f32(srcs[5]), f32(srcs[4]), self._rbf_dmax) if _USE_RBF_LUT and srcs[5].is_cuda else None,
                "bdist": f32(srcs[6]), "relpos": f32(srcs[7]),
                "coord": [split_rows(srcs[8]), f32(srcs[9]), f32(srcs[10]).reshape(-1).contiguous()],
                "seq": [split_k(srcs[11]), f32(srcs[12]), split_k(srcs[13])],
                "atom_tabs": [f32(t) for t in srcs[14:23]],
                "bond_tabs": [f32(t) for t in srcs[23:26]],
            }
            return w

        return self._pack.get(srcs, build)

    # ---- batch preparation (reference model.py:424-468, inference branch) -----------------------
    def prepare_batch(self, batch, id=None):
        if self.training_mode:
            raise NotImplementedError("training-mode masking needs residue_esm_tokens / ESM (SURVEY N8)")
        atom_mask, residue_mask = batch["atom_mask"], batch["residue_mask"]
        ca = batch["residue_atom_pos"][:, :, 1]
        residue_type = batch["residue_type"]
        batch["residue_one_hot"] = F.one_hot(residue_type, num_classes=NUM_RESIDUE_CLASSES) * 2.0 - 1.0
        pos = atom_mask.unsqueeze(-1) * batch["atom_pos"] + residue_mask.unsqueeze(-1) * ca
        keep, drop, _ = self.RandomMaskingBlock(residue_mask, self.mask_prob, inverse_mask=True, stochastic=False)
        batch["residue_esm"] = batch["residue_esm"] * keep.unsqueeze(-1)
        batch["residue_type_masked"] = (residue_type * keep).long()
        batch["residue_one_hot"] = batch["residue_one_hot"] * keep.unsqueeze(-1)
        batch["residue_extra_mask"] = keep
        batch["residue_inv_extra_mask"] = drop
        batch["x"] = 0.1 * pos
        batch["residue_and_atom_mask"] = atom_mask + residue_mask
        # not in the reference: whether the batch has any padding at all (one more host read next to the masking module's own
        # .item()): a pure performance hint for the triangle-attention core, see ops.triangle_attention
        batch["_all_valid"] = bool((batch["residue_and_atom_mask"] > 0.5).all())
        return batch

    # ---- step-invariant embeddings, cached per batch ---------------------------------------------
    _STATIC_KEYS = ("residue_esm", "bond_feats", "bond_mask", "bond_distance", "residue_index", "residue_chain_index",
                    "atom_mask", "residue_mask")

    def _static_embeddings(self, batch):
        """ESM projection and the bond / relpos pair terms do not depend on the diffusion step.  The cache is keyed on the
        IDENTITY of the batch tensors, which it keeps alive: a raw data_ptr of a tensor nobody holds can be handed out again
        by the caching allocator for the next same-shape batch (and inference tensors carry no version counter), which
        would silently reuse another complex's embeddings."""
        w = self._weights()
        held = self._static_key
        fresh = held is None or held[2] is not w or held[3] != weights_epoch() or any(
            batch[k] is not t or tensor_version(batch[k]) != v for k, t, v in zip(self._STATIC_KEYS, held[0], held[1]))
        if fresh:
            b = {k: batch[k].contiguous() for k in self._STATIC_KEYS}
            esm_emb = ops.esm_embed(self.cfg, b["residue_esm"], w["esm"][0])
            pair_static = ops.pair_embed_static(self.cfg, b, w["bond_tabs"], w["bdist"], w["relpos"])
            self._static = (esm_emb, pair_static)
            self._static_key = (tuple(batch[k] for k in self._STATIC_KEYS),
                                tuple(tensor_version(batch[k]) for k in self._STATIC_KEYS), w, weights_epoch())
        return self._static

    @contextlib.contextmanager
    def _ema_weights(self):
        """``with self.ema.average_parameters()`` (reference model.py:227,250) plus invalidation of every packed weight
        copy: torch_ema swaps through ``param.data.copy_``, which no version counter sees."""
        try:
            with self.ema.average_parameters():
                invalidate_packed_weights()
                yield
        finally:
            invalidate_packed_weights()

    def _denoise(self, batch, z, seq_t, mask, t, bufs=None, sampler_state=None, probe=None):
        """One network evaluation (reference model.py:318-375).  `bufs` lets the sampler reuse storage."""
        cfg, w = self.cfg, self._weights()
        B, N = mask.shape
        rec = probe or (lambda n, x: None)
        esm_emb, pair_static = self._static_embeddings(batch)
        bufs = bufs if bufs is not None else {}
        single = ops.single_embed(cfg, batch["atom_feats"].contiguous(), batch["atom_mask"].contiguous(),
                                  batch["residue_mask"].contiguous(), seq_t.contiguous(), esm_emb, w["atom_tabs"],
                                  w["w_type"], out=bufs.get("single"))
        rec("embed_single", single)
        a, b = self.Denoiser.opm.project(cfg, single, mask, bufs.get("opm_a"), bufs.get("opm_b"))
        _, (w_o, b_o) = self.Denoiser.opm.packed_weights()
        pair = bufs.get("pair")
        if pair is None:
            pair = torch.empty(B, N, N, cfg.pair_dim, dtype=torch.float32, device=z.device)
        ops.pair_embed(cfg, pair_static, z.contiguous(), mask, None if sampler_state is not None else t.contiguous(),
                       a, b, w["pair_dyn"] + [w_o, b_o], pair, sampler_state=sampler_state, rbf_lut=w["rbf_lut"])
        rec("Denoiser.opm", pair)
        single, pair = self.Denoiser.trunk_(single, pair, mask, probe=probe, all_valid=bool(batch.get("_all_valid", False)))
        noise_pred = ops.coord_head(cfg, pair, z.contiguous(), mask, w["coord"], out=bufs.get("noise_pred"))
        seq_pred = ops.seq_head(cfg, single, w["seq"], out=bufs.get("seq_pred"))
        return noise_pred, seq_pred

    def forward(self, batch, z, seq_t, mask, t):
        """reference model.py:254-316.  With autograd enabled the evaluation is one autograd node over every trainable
        parameter (autograd.DenoiserFunction: fused forward kernels + block checkpoints, prd_<op>_bwd kernels backward);
        gradients with respect to z / seq_t are not produced (the reference's training never asks for them)."""
        if torch.is_grad_enabled():
            from .autograd import DenoiserFunction, trainable_parameters
            params = [p for _, p in trainable_parameters(self)]
            if params:
                if z.requires_grad or seq_t.requires_grad:
                    raise NotImplementedError("gradients with respect to z / seq_t are not part of this build")
                return DenoiserFunction.apply(self, batch, z, seq_t, mask.contiguous(), t, *params)
        return self._denoise(batch, z, seq_t, mask.contiguous(), t)

    def sample_step(self, batch, z, seq_t, mask, t):
        """reference model.py:318-375.  Repeated calls on the same prepared batch (what a sampling loop does) replay ONE
        captured CUDA graph of the 119 kernels: the four inputs are copied into the graph's static buffers, the outputs are
        returned as fresh tensors.  ``PRD_STEP_GRAPH=0`` (or autograd / a CPU tensor) takes the eager path."""
        if _STEP_GRAPH and z.is_cuda and not torch.is_grad_enabled():
            return self._sample_step_graph(batch, z, seq_t, mask, t)
        return self._denoise(batch, z, seq_t, mask.contiguous(), t)

    def _sample_step_graph(self, batch, z, seq_t, mask, t):
        w = self._weights()
        keys = self._STATIC_KEYS + ("atom_feats",)
        sg = self._step_graph
        # every parameter's storage and in-place version (the graph baked the packed copies of ALL of them in)
        wsig = (sum(p.data_ptr() for p in self.parameters()), sum(tensor_version(p) for p in self.parameters()))
        stale = sg is None or sg["w"] is not w or sg["wsig"] != wsig or sg["epoch"] != weights_epoch() or tuple(z.shape) != sg["shape"] or any(
            batch[k] is not tns or tensor_version(batch[k]) != v for k, tns, v in zip(keys, sg["tensors"], sg["versions"]))
        if stale:
            self._step_graph = None  # frees the previous graph's pool before the new capture
            dev = z.device
            B, N = mask.shape
            st = {"z": torch.empty(B, N, 3, device=dev), "seq_t": torch.empty(B, N, NUM_RESIDUE_CLASSES, device=dev),
                  "mask": torch.empty(B, N, device=dev), "t": torch.empty(B, dtype=torch.int64, device=dev)}
            for k, v in (("z", z), ("seq_t", seq_t), ("mask", mask), ("t", t)):
                st[k].copy_(v)
            self._static_embeddings(batch)
            self._denoise(batch, st["z"], st["seq_t"], st["mask"], st["t"])  # eager warm-up: packs, lazy kernel attributes
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                ws = ops.reserve_workspace(self.cfg, B, N, dev)
                with torch.cuda.graph(graph, stream=side):
                    out = self._denoise(batch, st["z"], st["seq_t"], st["mask"], st["t"])
            torch.cuda.current_stream(dev).wait_stream(side)
            sg = self._step_graph = {"w": w, "wsig": wsig, "epoch": weights_epoch(), "shape": tuple(z.shape), "static": st, "out": out, "graph": graph,
                                     "ws": ws, "tensors": tuple(batch[k] for k in keys),
                                     "versions": tuple(tensor_version(batch[k]) for k in keys)}
        st = sg["static"]
        st["z"].copy_(z, non_blocking=True)
        st["seq_t"].copy_(seq_t, non_blocking=True)
        st["mask"].copy_(mask, non_blocking=True)
        st["t"].copy_(t, non_blocking=True)
        sg["graph"].replay()
        return sg["out"][0].clone(), sg["out"][1].clone()

    def predict_step(self, batch, batch_idx, noise: Optional[Dict[str, torch.Tensor]] = None):
        """reference model.py:249-252: sample under the EMA weights (``noise``: optional injected draws, see sample)."""
        with self._ema_weights():
            return self.sample(batch, noise=noise)

    # ---- training objective (reference model.py:471-549) ---------------------------
    def _sched_table(self):
        tab = getattr(self, "_sched", None)
        if tab is None or tab.device != self._device or tab.shape[0] != self.num_steps:
            self._sched = tab = torch.stack([self.sqrt_alphas_cumprod, self.sqrt_one_minus_alphas_cumprod], dim=1).contiguous()
        return tab

    def q(self, x, seq, t, noise_z, noise_seq, batch):
        """Forward noising step (reference model.py:471-488) -> (z_t, seq_t, seq_t1, t1)."""
        z_t, seq_t, seq_t1 = ops.diffusion_q(self.cfg, x.contiguous(), seq.contiguous(), t.contiguous(), noise_z.contiguous(),
                                             noise_seq.contiguous(), batch["residue_extra_mask"].contiguous(),
                                             batch["residue_inv_extra_mask"].contiguous(), self._sched_table())
        return z_t, seq_t, seq_t1, (t - 1).clamp(min=0)

    def diffusion_loss(self, batch, x, mask, t, noise: Optional[Dict[str, torch.Tensor]] = None, want_grads: bool = False,
                       detail: Optional[dict] = None):
        """reference model.py:490-526: draws noise_z / noise_seq (or takes raw N(0,1) draws from ``noise`` with keys
        ``z`` [B,N,3] and ``seq`` [B,N,21]), removes their masked means, applies q(), evaluates the network and the three
        loss terms, all on device.  Returns diff_loss [B]; ``detail`` (a dict) receives loss, terms, the network outputs
        and, with ``want_grads``, d loss / d noise_pred and d loss / d seq_pred."""
        seq, residue_mask = batch["residue_one_hot"].to(torch.float32).contiguous(), batch["residue_mask"].contiguous()
        mask, x = mask.contiguous(), x.contiguous()
        noise_z = (noise["z"].to(x.device, torch.float32).clone() if noise is not None else torch.randn_like(x)).contiguous()
        noise_seq = (noise["seq"].to(x.device, torch.float32).clone() if noise is not None else torch.randn_like(seq)).contiguous()
        ops.remove_mean(self.cfg, noise_z, mask)
        ops.remove_mean(self.cfg, noise_seq, residue_mask)
        z_t, seq_t, seq_t1, _ = self.q(x, seq, t, noise_z, noise_seq, batch)
        if torch.is_grad_enabled():
            # training: (noise_pred, seq_pred) and the loss are autograd nodes backed by the backward kernels
            from .autograd import LossFunction
            noise_pred, seq_pred = self.forward(batch, z_t, seq_t, mask, t.contiguous())
            d = detail if detail is not None else {}
            loss = LossFunction.apply(noise_pred, seq_pred, self.cfg, noise_z, noise_seq, seq_t1, mask, residue_mask,
                                      batch["residue_type"].contiguous(), t.contiguous(), self._sched_table(), d)
            d.update(noise_pred=noise_pred, seq_pred=seq_pred, z_t=z_t, seq_t=seq_t, loss_graph=loss)
            self._last_loss = loss
            return d["diff_loss"]
        noise_pred, seq_pred = self._denoise(batch, z_t, seq_t, mask, t.contiguous())
        loss, diff, terms, d_noise, d_seq = ops.diffusion_loss(
            self.cfg, noise_pred, seq_pred, noise_z, noise_seq, seq_t1, mask, residue_mask,
            batch["residue_type"].contiguous(), t.contiguous(), self._sched_table(), want_grads=want_grads)
        if detail is not None:
            detail.update(loss=loss, terms=terms, noise_pred=noise_pred, seq_pred=seq_pred, z_t=z_t, seq_t=seq_t,
                          d_noise_pred=d_noise, d_seq_pred=d_seq)
        return diff

    def _objective(self, batch, batch_idx, noise, detail, want_grads):
        """Shared by training_step / validation_step (reference model.py:528-541 / :226-240)."""
        if not self.setup_schedule:
            self.run_setup_schedule()
            self.setup_schedule = True
        batch = self.prepare_batch(batch, batch_idx)
        x, mask = batch["x"], batch["residue_and_atom_mask"]
        t = torch.randint(0, self.num_steps, size=(x.size(0),)).to(x.device)  # CPU generator, as the reference
        d = detail if detail is not None else {}
        self.diffusion_loss(batch, x, mask, t, noise=noise, want_grads=want_grads, detail=d)
        d["t"] = t
        return d.get("loss_graph", d["loss"].reshape(())), x.size(0)

    def training_step(self, batch, batch_idx, noise: Optional[Dict[str, torch.Tensor]] = None, detail: Optional[dict] = None):
        """reference model.py:528-549: prepare_batch, t ~ randint(0, T), loss = mean(diff_loss / num_nodes).  With autograd
        enabled the returned loss carries a graph whose backward runs on the prd_<op>_bwd kernels (autograd.py), so
        ``loss.backward()`` fills ``.grad`` of every trainable parameter exactly like Lightning's backward does for the
        reference; under ``torch.no_grad()`` only the objective (and, through ``detail``, d loss / d outputs) is evaluated."""
        loss, bs = self._objective(batch, batch_idx, noise, detail, want_grads=detail is not None)
        if hasattr(self, "log"):
            self.log("train_loss", loss, on_step=True, on_epoch=True, sync_dist=True, batch_size=bs)
        return loss

    @torch.no_grad()
    def validation_step(self, batch, batch_idx, noise: Optional[Dict[str, torch.Tensor]] = None, detail: Optional[dict] = None):
        """reference model.py:226-247: the same objective under the EMA weights, logged as ``val_loss`` (the reference
        returns None; the loss is returned here as well)."""
        with self._ema_weights():
            loss, bs = self._objective(batch, batch_idx, noise, detail, want_grads=False)
        if hasattr(self, "log"):
            self.log("val_loss", loss, on_epoch=True, sync_dist=True, batch_size=bs)
        return loss

    # ---- checkpoint / optimiser glue (reference model.py:192-217; host side, not on the hot path) ----------
    def to(self, *args, **kwargs):
        out = torch._C._nn._parse_to(*args, **kwargs)
        self.ema.to(device=out[0], dtype=out[1])
        return super().to(*args, **kwargs)

    def on_save_checkpoint(self, checkpoint):
        checkpoint["ema_state_dict"] = self.ema.state_dict()

    def on_load_checkpoint(self, checkpoint):
        self.ema.load_state_dict(checkpoint["ema_state_dict"])

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=self.learning_rate)
        scheduler = torch.optim.lr_scheduler.LinearLR(optimizer, start_factor=1.0 / self.warmup_steps,
                                                      total_iters=self.warmup_steps - 1)
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "step"}}

    if _Base is nn.Module:  # Lightning provides this otherwise
        @classmethod
        def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, **overrides):
            """Lightning-free reader of a Lightning checkpoint of the reference model (generate.py:103-107):
            ``hyper_parameters`` (+ keyword overrides such as num_steps / mask_prob) -> constructor, ``state_dict`` loaded
            strictly, ``ema_state_dict`` handed to on_load_checkpoint."""
            ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
            hp = ckpt.get("hyper_parameters", {})
            hp = dict(vars(hp)) if isinstance(hp, Namespace) else dict(hp)
            hp.update(overrides)
            model = cls(Namespace(**hp))
            model.load_state_dict(ckpt["state_dict"], strict=strict)
            if "ema_state_dict" in ckpt:
                model.on_load_checkpoint(ckpt)
            return model

    # ---- sampler (reference model.py:377-422) ------------------------------------------------
    @torch.inference_mode()
    def sample(self, batch, noise: Optional[Dict[str, torch.Tensor]] = None, use_cuda_graph: bool = True,
               trace: Optional[list] = None, prepared: bool = False):
        """DDPM ancestral sampling.  `noise` may inject pre-generated draws (keys ``z_T`` [B,N,3],
        ``seq_T`` [B,N,21], ``steps`` [T-1,B,N,3], raw N(0,1), in the reference's draw order); otherwise
        they come from torch's CUDA generator.  ``prepared=True``: ``batch`` already went through prepare_batch (the
        sample-parallel driver masks the FULL batch jointly before sharding it, sampling.py) and must not be masked again."""
        if not self.setup_schedule:
            self.run_setup_schedule()
            self.setup_schedule = True
        if not prepared:
            batch = self.prepare_batch(batch)
        cfg = self.cfg
        x, mask = batch["x"], batch["residue_and_atom_mask"].contiguous()
        residue_mask, seq = batch["residue_mask"].contiguous(), batch["residue_one_hot"]
        keep, drop = batch["residue_extra_mask"], batch["residue_inv_extra_mask"]
        B, N = mask.shape
        dev, T = x.device, self.num_steps
        ops.reserve_workspace(cfg, B, N, dev)

        def draw(key, shape):
            if noise is not None and key in noise:
                return noise[key].to(dev, torch.float32).contiguous().clone()
            return torch.randn(shape, device=dev, dtype=torch.float32)

        z = ops.remove_mean(cfg, draw("z_T", (B, N, 3)), mask)
        seq_noise = ops.remove_mean(cfg, draw("seq_T", (B, N, NUM_RESIDUE_CLASSES)), residue_mask)
        seq_t = (keep.unsqueeze(-1) * seq + drop.unsqueeze(-1) * seq_noise).float().contiguous()
        steps = draw("steps", (max(T - 1, 1), B, N, 3))
        if T > 1:
            ops.remove_mean(cfg, steps.view(-1, N, 3), mask)
        state = torch.tensor([T - 1, 0], dtype=torch.int32, device=dev)
        bufs = {
            "single": torch.empty(B, N, cfg.single_dim, device=dev),
            "pair": torch.empty(B, N, N, cfg.pair_dim, device=dev),
            "opm_a": torch.empty(B, N, cfg.single_dim // 4, device=dev),
            "opm_b": torch.empty(B, N, cfg.single_dim // 4, device=dev),
            "noise_pred": torch.empty(B, N, 3, device=dev),
            "seq_pred": torch.empty(B, N, NUM_RESIDUE_CLASSES, device=dev),
        }
        self._static_embeddings(batch)  # outside the graph: step invariant

        def one_step():
            eps, sp = self._denoise(batch, z, seq_t, mask, None, bufs=bufs, sampler_state=state)
            ops.sampler_update(cfg, eps, sp, steps, self._coef, z, seq_t, state)

        graph = None
        done = 0
        if use_cuda_graph and T > 2 and trace is None:
            one_step()  # eager warm-up step (also settles lazy kernel attributes)
            done = 1
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                graph_ws = ops.reserve_workspace(cfg, B, N, dev)  # the capture stream's own scratch, held with the graph
                with torch.cuda.graph(graph, stream=side):
                    one_step()
            torch.cuda.current_stream(dev).wait_stream(side)
            # capture does not execute: state / z are still those after `done` steps
        for _ in range(done, T):
            if graph is not None:
                graph.replay()
            else:
                one_step()
            if trace is not None:
                trace.append((z.clone(), bufs["seq_pred"].clone(), bufs["noise_pred"].clone()))
        pos = 10.0 * z
        del graph  # before its workspace reference (graph_ws) goes out of scope
        return pos, residue_mask.unsqueeze(-1) * bufs["seq_pred"]
