"""Seeded synthetic complexes and weights for the ProteinReDiff denoiser hot path.

Everything here is deterministic given a seed (torch CPU ``Generator``), so the
same batches / state-dicts can be rebuilt on the GPU box without shipping the
tensors: golden fixtures under ``tests/golden/`` only store outputs plus a
checksum of what this module generated.

Layout contracts restated from the reference (nothing is imported from it):

* batch dict  = output of ``collate_fn``         (ProteinReDiff/data.py:80-142)
* state-dict  = ``ProteinReDiffModel.state_dict`` (ProteinReDiff/model.py:55-126,
  ProteinReDiff/modules.py:35-97,170-404, ProteinReDiff/models/AF2_modules.py:189-545)
* vocab sizes = ``ALLOWABLE_{ATOM,BOND}_FEATURES`` (ProteinReDiff/features.py:31-60)
"""
from __future__ import annotations

import dataclasses
import hashlib
import math
from argparse import Namespace
from typing import Dict, List, Tuple

import torch

# features.py:31-60 -- number of categories per categorical atom / bond feature.
ATOM_VOCAB = (119, 4, 12, 12, 10, 6, 6, 2, 2)
BOND_VOCAB = (5, 6, 2)
NUM_RESIDUE_CLASSES = 21  # protein.py:28-31 (20 residue types) + 1 (pad / X), model.py:90


@dataclasses.dataclass(frozen=True)
class DenoiserConfig:
    """Hyper-parameters read by the reference (model.py:137-170, modules.py:356-363)."""

    single_dim: int = 512
    pair_dim: int = 64
    head_dim: int = 16
    num_heads: int = 4
    transition_factor: int = 4
    num_blocks: int = 4
    esm_dim: int = 1280
    time_dim: int = 256
    dist_dim: int = 256
    max_bond_distance: int = 7
    max_relpos: int = 32
    num_steps: int = 64
    diffusion_schedule: str = "linear"
    mask_prob: float = 0.15
    n_recycles: int = 4
    training_mode: bool = False
    learning_rate: float = 4e-4
    warmup_steps: int = 1000
    ema_decay: float = 0.999

    def to_namespace(self) -> Namespace:
        return Namespace(**dataclasses.asdict(self))

    @staticmethod
    def from_args(args) -> "DenoiserConfig":
        if isinstance(args, DenoiserConfig):
            return args
        if not isinstance(args, dict):
            args = vars(args)
        names = {f.name for f in dataclasses.fields(DenoiserConfig)}
        return DenoiserConfig(**{k: v for k, v in args.items() if k in names})


# Named configurations used by tests / bench (BASELINE.json "configs").
PAPER = DenoiserConfig()  # single 512 / pair 64 / 4 blocks (README.md:148-160)
README = DenoiserConfig(single_dim=256, pair_dim=32)  # README.md:133-141 example dims
TINY = DenoiserConfig(single_dim=64, pair_dim=32, num_blocks=2, esm_dim=48, time_dim=32, dist_dim=32)


def state_dict_spec(cfg: DenoiserConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor in the reference state-dict, in module order.

    kind selects the synthetic initialiser below; it also documents the role:
    ``w``   dense weight [out, in]          ``b``    additive bias
    ``gw``  gate weight                      ``gb``   gate bias (reference init = 1)
    ``fw``  "final" projection weight (reference init = 0)
    ``lnw`` / ``lnb`` LayerNorm affine       ``emb``  embedding table
    ``freq`` / ``center`` fixed buffers (modules.py:77-79, 91-93)
    """
    cs, cz, H, c = cfg.single_dim, cfg.pair_dim, cfg.num_heads, cfg.head_dim
    tf = cfg.transition_factor
    spec: List[Tuple[str, Tuple[int, ...], str]] = []
    add = lambda n, s, k: spec.append((n, tuple(s), k))

    # modules.py:366-375 -> AF2_modules.py:369-420 (SPAttention), :476-501 (OuterProductUpdate)
    p = "Denoiser.SPAAttnBlock."
    add(p + "layer_norm_m.weight", (cs,), "lnw")
    add(p + "layer_norm_m.bias", (cs,), "lnb")
    add(p + "linear_z.0.weight", (cz,), "lnw")
    add(p + "linear_z.0.bias", (cz,), "lnb")
    add(p + "linear_z.1.weight", (H, cz), "w")
    add(p + "mha.linear_q.weight", (H * cs, cs), "w")
    add(p + "mha.linear_k.weight", (H * cs, cs), "w")
    add(p + "mha.linear_v.weight", (H * cs, cs), "w")
    add(p + "mha.linear_o.weight", (cs, H * cs), "fw")
    add(p + "mha.linear_o.bias", (cs,), "b")
    add(p + "mha.linear_g.weight", (H * cs, cs), "gw")
    add(p + "mha.linear_g.bias", (H * cs,), "gb")
    p = "Denoiser.opm."
    ch = cs // 4
    add(p + "layer_norm.weight", (cs,), "lnw")
    add(p + "layer_norm.bias", (cs,), "lnb")
    add(p + "linear_1.weight", (ch, cs), "w")
    add(p + "linear_1.bias", (ch,), "b")
    add(p + "linear_2.weight", (ch, cs), "w")
    add(p + "linear_2.bias", (ch,), "b")
    add(p + "linear_out.weight", (cz, ch), "fw")
    add(p + "linear_out.bias", (cz,), "b")
    # modules.py:290-326 (FoldingBlock)
    for k in range(cfg.num_blocks):
        p = f"Denoiser.folding_blocks.{k}."
        add(p + "attn_bias.1.weight", (H, cz), "w")
        add(p + "attn_bias.1.bias", (H,), "b")
        q = p + "single_attn."
        add(q + "q_proj.weight", (H * c, cs), "w")
        add(q + "k_proj.weight", (H * c, cs), "w")
        add(q + "v_proj.weight", (H * c, cs), "w")
        add(q + "gate_proj.weight", (H * c, cs), "gw")
        add(q + "gate_proj.bias", (H * c,), "gb")
        add(q + "out_proj.weight", (cs, H * c), "fw")
        add(q + "out_proj.bias", (cs,), "b")
        add(p + "single_fc.1.weight", (cs * tf, cs), "w")
        add(p + "single_fc.1.bias", (cs * tf,), "b")
        add(p + "single_fc.3.weight", (cs, cs * tf), "fw")
        add(p + "single_fc.3.bias", (cs,), "b")
        add(p + "outer_linear.linear.weight", (cz, 2 * cs), "fw")
        add(p + "outer_linear.linear.bias", (cz,), "b")
        for mode in ("outgoing", "incoming"):
            q = p + f"pair_mul_{mode}."
            add(q + "ab_proj.weight", (2 * cz, cz), "w")
            add(q + "ab_proj.bias", (2 * cz,), "b")
            add(q + "ab_gate.weight", (2 * cz, cz), "gw")
            add(q + "ab_gate.bias", (2 * cz,), "gb")
            add(q + "out_proj.weight", (cz, cz), "fw")
            add(q + "out_proj.bias", (cz,), "b")
            add(q + "out_gate.weight", (cz, cz), "gw")
            add(q + "out_gate.bias", (cz,), "gb")
        for mode in ("starting", "ending"):
            q = p + f"pair_attn_{mode}.attn."
            add(q + "q_proj.weight", (H * c, cz), "w")
            add(q + "k_proj.weight", (H * c, cz), "w")
            add(q + "v_proj.weight", (H * c, cz), "w")
            add(q + "gate_proj.weight", (H * c, cz), "gw")
            add(q + "gate_proj.bias", (H * c,), "gb")
            add(q + "out_proj.weight", (cz, H * c), "fw")
            add(q + "out_proj.bias", (cz,), "b")
        add(p + "pair_fc.1.weight", (cz * tf, cz), "w")
        add(p + "pair_fc.1.bias", (cz * tf,), "b")
        add(p + "pair_fc.3.weight", (cz, cz * tf), "fw")
        add(p + "pair_fc.3.bias", (cz,), "b")
    # model.py:84-122
    for f, v in enumerate(ATOM_VOCAB):
        add(f"embed_atom_feats.embeddings.{f}.weight", (v, cs), "emb")
    add("embed_beta.0.weight", (cfg.time_dim // 2,), "freq")
    add("embed_beta.1.weight", (cz, cfg.time_dim), "w")
    add("embed_residue_type.1.weight", (cs, NUM_RESIDUE_CLASSES), "w")
    for f, v in enumerate(BOND_VOCAB):
        add(f"embed_bond_feats.embeddings.{f}.weight", (v, cz), "emb")
    add("embed_bond_distance.weight", (cfg.max_bond_distance + 1, cz), "emb")
    add("embed_residue_esm.1.weight", (cs, cfg.esm_dim), "w")
    add("embed_relpos.weight", (2 * cfg.max_relpos + 1, cz), "emb")
    add("embed_dist.0.center", (cfg.dist_dim,), "center")
    add("embed_dist.1.weight", (cz, cfg.dist_dim), "w")
    add("weight_radial.1.weight", (cz, cz), "w")
    add("weight_radial.1.bias", (cz,), "b")
    add("weight_radial.3.weight", (1, cz), "head_r")
    add("seq_mlp.1.weight", (cs, cs), "w")
    add("seq_mlp.1.bias", (cs,), "b")
    add("seq_mlp.3.weight", (NUM_RESIDUE_CLASSES, cs), "head_s")
    return spec


def make_state_dict(cfg: DenoiserConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded, non-degenerate fp32 weights with the reference's names and shapes.

    The reference's own initialiser zeros every output projection, which makes
    the whole network output identically zero, so parity would be vacuous.  Scales here are
    chosen so every residual branch contributes O(0.1-1) to the stream and the two heads
    give rms(noise_pred) ~ 0.3 and rms(logits) ~ 1 (the regime SURVEY §7 found non-chaotic
    for multi-step trajectories).
    """
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, shape, kind in state_dict_spec(cfg):
        if kind == "freq":  # modules.py:91-93
            t = torch.logspace(-4.0, 0.0, shape[0])
        elif kind == "center":  # modules.py:77-79
            t = torch.linspace(0.0, 2.0, shape[0])
        elif kind == "lnw":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "lnb":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "gb":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "emb":
            t = torch.randn(shape, generator=g)
        elif kind == "w":
            t = torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        elif kind == "gw":
            t = 0.5 * torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        elif kind == "fw":
            t = 0.3 * torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        elif kind == "head_r":
            t = 0.02 * torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        elif kind == "head_s":
            t = 1.0 * torch.randn(shape, generator=g) / math.sqrt(shape[-1])
        else:  # pragma: no cover
            raise ValueError(kind)
        sd[name] = t.to(torch.float32).contiguous()
    return sd


def make_batch(
    cfg: DenoiserConfig,
    sizes: List[Tuple[int, int]],
    seed: int = 0,
    n_total: int | None = None,
    two_chains: bool = False,
    with_positions: bool = False,
) -> Dict[str, torch.Tensor]:
    """A synthetic batch in ``collate_fn`` layout (data.py:80-142).

    ``sizes`` holds (num_atoms, num_residues) per batch row; rows are padded with zeros to
    ``n_total`` (default: the largest row).  Token order per row: ligand atoms, residues,
    padding.  ``residue_type`` is already shifted by +1 (data.py:100), 0 = pad.
    """
    g = torch.Generator().manual_seed(seed)
    B = len(sizes)
    N = max(a + r for a, r in sizes)
    if n_total is not None:
        if n_total < N:
            raise ValueError("n_total smaller than the largest complex")
        N = n_total
    f32, i64 = torch.float32, torch.int64
    b = {
        "atom_mask": torch.zeros(B, N, dtype=f32),
        "residue_mask": torch.zeros(B, N, dtype=f32),
        "bond_mask": torch.zeros(B, N, N, dtype=f32),
        "atom_feats": torch.zeros(B, N, len(ATOM_VOCAB), dtype=i64),
        "bond_feats": torch.zeros(B, N, N, len(BOND_VOCAB), dtype=i64),
        "bond_distance": torch.zeros(B, N, N, dtype=i64),
        "residue_type": torch.zeros(B, N, dtype=i64),
        "residue_chain_index": torch.zeros(B, N, dtype=i64),
        "residue_index": torch.zeros(B, N, dtype=i64),
        "atom_pos": torch.zeros(B, N, 3, dtype=f32),
        "residue_atom_pos": torch.zeros(B, N, 37, 3, dtype=f32),
        "residue_atom_mask": torch.zeros(B, N, 37, dtype=f32),
        "residue_esm": torch.zeros(B, N, cfg.esm_dim, dtype=f32),
        "num_atoms": torch.tensor([a for a, _ in sizes], dtype=i64),
        "num_residues": torch.tensor([r for _, r in sizes], dtype=i64),
    }
    for r, (na, nr) in enumerate(sizes):
        b["atom_mask"][r, :na] = 1.0
        b["residue_mask"][r, na : na + nr] = 1.0
        for f, v in enumerate(ATOM_VOCAB):
            b["atom_feats"][r, :na, f] = torch.randint(0, v, (na,), generator=g)
        bm = (torch.rand(na, na, generator=g) < 0.1).float()
        bm = torch.triu(bm, 1)
        bm = bm + bm.T
        b["bond_mask"][r, :na, :na] = bm
        for f, v in enumerate(BOND_VOCAB):
            bf = torch.randint(0, v, (na, na), generator=g)
            bf = torch.triu(bf, 1)
            bf = bf + bf.T
            b["bond_feats"][r, :na, :na, f] = bf * bm.long()
        bd = torch.randint(0, 12, (na, na), generator=g)  # exercises the clamp at 7
        bd = torch.triu(bd, 1)
        b["bond_distance"][r, :na, :na] = bd + bd.T
        b["residue_type"][r, na : na + nr] = torch.randint(1, NUM_RESIDUE_CLASSES, (nr,), generator=g)
        b["residue_index"][r, na : na + nr] = torch.arange(nr)
        if two_chains and nr >= 4:
            cut = nr // 2
            b["residue_chain_index"][r, na + cut : na + nr] = 1
            b["residue_index"][r, na + cut : na + nr] = torch.arange(nr - cut)
        b["residue_esm"][r, na : na + nr] = torch.randn(nr, cfg.esm_dim, generator=g)
        b["residue_atom_mask"][r, na : na + nr, :3] = 1.0
        if with_positions:
            b["atom_pos"][r, :na] = 5.0 * torch.randn(na, 3, generator=g)
            b["residue_atom_pos"][r, na : na + nr, :3] = 10.0 * torch.randn(nr, 3, 3, generator=g)
    return b


def make_complex(cfg: DenoiserConfig, num_atoms: int, num_residues: int, seed: int = 0) -> Dict[str, object]:
    """One featurised complex as ``ligand_to_data`` + ``protein_to_data`` return it (reference data.py:28-77): the
    per-item input of ``collate_fn``.  ``residue_type`` is 0..19 here (collate_fn adds 1); the ``*_mol`` entries are
    placeholders (RDKit objects in the reference)."""
    g = torch.Generator().manual_seed(seed)
    na, nr = num_atoms, num_residues
    bm = torch.triu((torch.rand(na, na, generator=g) < 0.2).float(), 1)
    bm = bm + bm.T
    bf = torch.stack([torch.randint(0, v, (na, na), generator=g) for v in BOND_VOCAB], -1) * bm.long().unsqueeze(-1)
    bd = torch.triu(torch.randint(0, 12, (na, na), generator=g), 1)
    return {
        "ligand_mol": f"ligand-{seed}", "num_atoms": na,
        "atom_feats": torch.stack([torch.randint(0, v, (na,), generator=g) for v in ATOM_VOCAB], -1),
        "atom_pos": torch.randn(na, 3, generator=g), "atom_mask": torch.ones(na),
        "bond_feats": bf, "bond_mask": bm, "bond_distance": bd + bd.T,
        "protein_mol": f"protein-{seed}", "num_residues": nr,
        "residue_type": torch.randint(0, NUM_RESIDUE_CLASSES - 1, (nr,), generator=g), "residue_mask": torch.ones(nr),
        "residue_chain_index": torch.zeros(nr, dtype=torch.int64), "residue_index": torch.arange(nr),
        "residue_atom_pos": torch.randn(nr, 37, 3, generator=g), "residue_atom_mask": (torch.rand(nr, 37, generator=g) < 0.5).float(),
        "residue_esm": torch.randn(nr, cfg.esm_dim, generator=g),
    }


def make_step_inputs(batch: Dict[str, torch.Tensor], num_steps: int, seed: int = 0):
    """Seeded (z, seq_t, mask, t) for one denoiser step.

    z is mean-removed unit noise in nm (model.py:399), seq_t unit noise, t uniform in [0, T).
    """
    g = torch.Generator().manual_seed(seed + 7919)
    mask = batch["atom_mask"] + batch["residue_mask"]
    B, N = mask.shape
    z = torch.randn(B, N, 3, generator=g)
    m3 = mask.unsqueeze(-1)
    z = z - m3 * (m3 * z).sum(1, keepdim=True) / m3.sum(1, keepdim=True)
    seq_t = torch.randn(B, N, NUM_RESIDUE_CLASSES, generator=g)
    t = torch.randint(0, num_steps, (B,), generator=g)
    return z, seq_t, mask, t


def checksum(tensors) -> str:
    """sha256 over the raw bytes of a tensor / list / dict of tensors (golden-fixture guard)."""
    h = hashlib.sha256()
    if isinstance(tensors, torch.Tensor):
        tensors = [tensors]
    if isinstance(tensors, dict):
        tensors = [tensors[k] for k in sorted(tensors) if isinstance(tensors[k], torch.Tensor)]
    for t in tensors:
        h.update(t.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()[:16]
