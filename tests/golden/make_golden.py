"""Generate golden vectors from the UNMODIFIED reference (run in the authoring container only).

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

The reference (/root/reference, read-only, pure Python) cannot travel to the GPU box and has
import-time dependencies that are absent here (pytorch_lightning, torch_ema, rdkit, Bio).  We
inject minimal ``sys.modules`` stubs for those, import
``ProteinReDiff.model.ProteinReDiffModel`` as is, load the seeded synthetic state-dict from
``protein_redesign_b200.synthetic`` with ``strict=True`` (which also pins the state-dict
names/shapes), run it on seeded synthetic batches and store the outputs.

Nothing of the reference is copied: fixtures hold only numeric outputs plus checksums of the
seeded inputs, which the tests regenerate locally.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from protein_redesign_b200 import synthetic as syn  # noqa: E402

REFERENCE_ROOT = "/root/reference"
ONLY_NEW = "--only-new" in sys.argv  # add missing fixtures without rewriting the existing ones


def _skip(kind, tag):
    return ONLY_NEW and os.path.exists(os.path.join(HERE, f"{kind}_{tag}.npz"))


def install_stubs() -> None:
    pl = types.ModuleType("pytorch_lightning")

    class LightningModule(torch.nn.Module):
        @property
        def device(self):
            return next(self.parameters()).device

        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

    pl.LightningModule = LightningModule
    pl.LightningDataModule = type("LightningDataModule", (), {})  # data.py:172 derives from it (never instantiated here)
    sys.modules["pytorch_lightning"] = pl

    te = types.ModuleType("torch_ema")

    class ExponentialMovingAverage:
        def __init__(self, params, decay):
            pass

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

    te.ExponentialMovingAverage = ExponentialMovingAverage
    sys.modules["torch_ema"] = te

    rdkit = types.ModuleType("rdkit")
    chem = types.ModuleType("rdkit.Chem")
    for n in ("Atom", "Bond", "Mol"):
        setattr(chem, n, type(n, (), {}))
    rdkit.Chem = chem
    sys.modules["rdkit"] = rdkit
    sys.modules["rdkit.Chem"] = chem
    bio = types.ModuleType("Bio")
    pdb = types.ModuleType("Bio.PDB")
    pdbp = types.ModuleType("Bio.PDB.PDBParser")
    pdbp.PDBParser = type("PDBParser", (), {})
    sys.modules["Bio"] = bio
    sys.modules["Bio.PDB"] = pdb
    sys.modules["Bio.PDB.PDBParser"] = pdbp


def load_reference_model(cfg: syn.DenoiserConfig, seed: int):
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from ProteinReDiff.model import ProteinReDiffModel  # type: ignore

    model = ProteinReDiffModel(cfg.to_namespace())
    sd = syn.make_state_dict(cfg, seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model, sd


def clone_batch(b):
    return {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}


def ref_prepare(model, batch, torch_seed):
    torch.manual_seed(torch_seed)
    return model.prepare_batch(clone_batch(batch))


def capture_modules(model, names):
    """Forward hooks that record the output of the named submodules."""
    store = {}
    handles = []
    mods = dict(model.named_modules())
    for n in names:
        def hook(mod, inp, out, n=n):
            if isinstance(out, tuple):  # Denoiser returns (single, pair, cache)
                for i, o in enumerate(out):
                    if isinstance(o, torch.Tensor):
                        store[f"{n}:{i}"] = o.detach().clone()
            else:
                store[n] = out.detach().clone()
        handles.append(mods[n].register_forward_hook(hook))
    return store, handles


def gen_step_case(tag, cfg, sizes, seed, n_total=None, two_chains=False, probes=False, out=None):
    if _skip("step", tag):
        return
    model, sd = load_reference_model(cfg, seed)
    batch = syn.make_batch(cfg, sizes, seed=seed, n_total=n_total, two_chains=two_chains)
    pb = ref_prepare(model, batch, torch_seed=seed)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    rec = {
        "weights_checksum": syn.checksum(sd),
        "batch_checksum": syn.checksum(batch),
        "inputs_checksum": syn.checksum([z, seq_t, mask, t]),
        "keep_mask": pb["residue_extra_mask"].numpy(),
        "drop_mask": pb["residue_inv_extra_mask"].numpy(),
        "residue_one_hot": pb["residue_one_hot"].numpy(),
        "residue_type_masked": pb["residue_type_masked"].numpy(),
        "x": pb["x"].numpy(),
    }
    names = []
    if probes:
        names = ["Denoiser.opm", "Denoiser.SPAAttnBlock"]
        for k in range(cfg.num_blocks):
            p = f"Denoiser.folding_blocks.{k}."
            names += [p + s for s in ("attn_bias", "single_attn", "single_fc", "outer_linear", "pair_mul_outgoing",
                                      "pair_mul_incoming", "pair_attn_starting", "pair_attn_ending", "pair_fc")]
        names += ["Denoiser", "weight_radial", "embed_dist", "embed_beta", "embed_residue_type", "embed_residue_esm"]
    store, handles = capture_modules(model, names)
    with torch.inference_mode():
        noise_pred, seq_pred = model.sample_step(pb, z, seq_t, mask, t)
        fwd_noise, fwd_seq = model.forward(pb, z, seq_t, mask, t)
    for h in handles:
        h.remove()
    assert torch.equal(noise_pred, fwd_noise) and torch.equal(seq_pred, fwd_seq)
    rec["noise_pred"] = noise_pred.numpy()
    rec["seq_pred"] = seq_pred.numpy()
    for n, v in store.items():
        rec["probe:" + n] = v.numpy()
    path = os.path.join(HERE, f"step_{tag}.npz")
    np.savez_compressed(path, **rec)
    print(f"{tag}: noise rms {noise_pred.square().mean().sqrt():.4f} logits rms {seq_pred.square().mean().sqrt():.4f}"
          f" -> {os.path.relpath(path, ROOT)} ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_sample_case(tag, cfg, sizes, seed, n_total=None):
    """Full sampler with injected noise: torch.randn_like is replaced by draws from a seeded
    generator, in the reference's own draw order."""
    if _skip("sample", tag):
        return
    model, sd = load_reference_model(cfg, seed)
    batch = syn.make_batch(cfg, sizes, seed=seed, n_total=n_total)
    g = torch.Generator().manual_seed(seed + 31337)
    draws = []

    def fake_randn_like(x, **kw):
        n = torch.randn(x.shape, generator=g, dtype=x.dtype)
        draws.append(n)
        return n

    orig = torch.randn_like
    torch.randn_like = fake_randn_like
    try:
        torch.manual_seed(seed)
        pos, logits = model.sample(clone_batch(batch))
    finally:
        torch.randn_like = orig
    rec = {
        "weights_checksum": syn.checksum(sd),
        "batch_checksum": syn.checksum(batch),
        "pos": pos.numpy(),
        "logits": logits.numpy(),
        "num_draws": len(draws),
    }
    # schedule tables pinned as well (model.py:172-190)
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_betas", "sqrt_alphas", "sqrt_one_minus_alphas_cumprod",
              "sqrt_alphas_cumprod"):
        rec["sched:" + k] = getattr(model, k).numpy()
    path = os.path.join(HERE, f"sample_{tag}.npz")
    np.savez_compressed(path, **rec)
    print(f"{tag}: pos rms {pos.square().mean().sqrt():.3f} A -> {os.path.relpath(path, ROOT)}")


def gen_loss_case(tag, cfg, sizes, seed, n_total=None, two_chains=False):
    """training_step (model.py:528-549, training_mode=False: SURVEY N8) with injected randn_like draws; stores the
    loss, diff_loss, t, the q() outputs, the network outputs and d loss / d (noise_pred, seq_pred) from autograd."""
    if _skip("loss", tag):
        return
    model, sd = load_reference_model(cfg, seed)
    batch = syn.make_batch(cfg, sizes, seed=seed, n_total=n_total, two_chains=two_chains, with_positions=True)
    g = torch.Generator().manual_seed(seed + 4242)
    cap = {}

    def fake_randn_like(x, **kw):
        return torch.randn(x.shape, generator=g, dtype=x.dtype)

    fwd = model.forward

    def spy_forward(b, z, seq_t, mask, t):
        noise_pred, seq_pred = fwd(b, z, seq_t, mask, t)
        noise_pred.retain_grad()
        seq_pred.retain_grad()
        cap.update(z_t=z.detach().clone(), seq_t=seq_t.detach().clone(), t=t.clone(), noise_pred=noise_pred,
                   seq_pred=seq_pred)
        return noise_pred, seq_pred

    diff = {}
    dl = model.diffusion_loss

    def spy_loss(*a, **k):
        out = dl(*a, **k)
        diff["diff_loss"] = out.detach().clone()
        return out

    model.forward = spy_forward
    model.diffusion_loss = spy_loss
    orig = torch.randn_like
    torch.randn_like = fake_randn_like
    try:
        torch.manual_seed(seed)
        loss = model.training_step(clone_batch(batch), 0)
        loss.backward()
    finally:
        torch.randn_like = orig
    # parameter gradients of the reference's backward (through per-block checkpointing, modules.py:399): per parameter the
    # L2 norm and the projection on a seeded random direction (2 numbers instead of the full tensor)
    gp = torch.Generator().manual_seed(seed + 777)
    names, norms, projs = [], [], []
    for n, prm in model.named_parameters():
        if prm.grad is None:
            continue
        d = torch.randn(prm.shape, generator=gp)
        names.append(n)
        norms.append(float(prm.grad.norm()))
        projs.append(float((prm.grad * d).sum()))
    rec = {
        "weights_checksum": syn.checksum(sd),
        "batch_checksum": syn.checksum(batch),
        "grad_names": np.array(names),
        "grad_norms": np.array(norms, dtype=np.float64),
        "grad_projs": np.array(projs, dtype=np.float64),
        "loss": loss.detach().numpy(),
        "diff_loss": diff["diff_loss"].numpy(),
        "t": cap["t"].numpy(),
        "z_t": cap["z_t"].numpy(),
        "seq_t": cap["seq_t"].numpy(),
        "noise_pred": cap["noise_pred"].detach().numpy(),
        "seq_pred": cap["seq_pred"].detach().numpy(),
        "d_noise_pred": cap["noise_pred"].grad.numpy(),
        "d_seq_pred": cap["seq_pred"].grad.numpy(),
    }
    path = os.path.join(HERE, f"loss_{tag}.npz")
    np.savez_compressed(path, **rec)
    print(f"{tag}: loss {float(loss):.5f} t {cap['t'].tolist()} -> {os.path.relpath(path, ROOT)}"
          f" ({os.path.getsize(path) / 1024:.0f} KiB)")


def gen_collate_case():
    """collate_fn (data.py:80-142) on three synthetic complexes of different sizes: every tensor of the batch."""
    if ONLY_NEW and os.path.exists(os.path.join(HERE, "collate.npz")):
        return
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    from ProteinReDiff.data import collate_fn  # type: ignore

    items = [syn.make_complex(syn.TINY, na, nr, seed=60 + i) for i, (na, nr) in enumerate([(5, 9), (3, 14), (7, 4)])]
    batch = collate_fn(items)
    out = {k: v.numpy() for k, v in batch.items() if isinstance(v, torch.Tensor)}
    out["mol_lists"] = np.array([",".join(batch["ligand_mol"]), ",".join(batch["protein_mol"])])
    np.savez_compressed(os.path.join(HERE, "collate.npz"), **out)
    print("collate:", {k: v.shape for k, v in out.items()})


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    gen_collate_case()
    # (1) tiny dims, every module probed, ragged batch with padding + two chains.
    gen_step_case("tiny_probes", syn.TINY, [(5, 14), (3, 9)], seed=1, n_total=22, two_chains=True, probes=True)
    # (2) README dims (BASELINE config 1 shape family), one ragged batch.
    gen_step_case("readme_n40", syn.README, [(8, 32), (6, 27)], seed=2)
    # (3) paper dims (north-star dims) at a size the CPU finishes in seconds; includes padding.
    gen_step_case("paper_n72", syn.PAPER, [(12, 60), (9, 50)], seed=3)
    gen_step_case("paper_n128", syn.PAPER, [(16, 112)], seed=4)
    # (4) sampler trajectories with injected noise (cosine + linear schedules).
    gen_sample_case("tiny_T8", syn.DenoiserConfig(**{**syn.TINY.__dict__, "num_steps": 8, "mask_prob": 0.3}),
                    [(5, 14), (3, 9)], seed=5)
    gen_sample_case("tiny_T6_cos", syn.DenoiserConfig(**{**syn.TINY.__dict__, "num_steps": 6, "mask_prob": 1.0,
                                                         "diffusion_schedule": "cosine"}), [(4, 12)], seed=6)
    # north_star: a 50-step fixed-noise trajectory
    gen_sample_case("tiny_T50", syn.DenoiserConfig(**{**syn.TINY.__dict__, "num_steps": 50, "mask_prob": 0.3}),
                    [(5, 14), (3, 9)], seed=8)
    # (5) training objective (a18): q(), the three loss terms and the gradient with respect to the network outputs
    gen_loss_case("tiny", syn.DenoiserConfig(**{**syn.TINY.__dict__, "mask_prob": 0.15}), [(5, 14), (3, 9)], seed=9,
                  n_total=22, two_chains=True)
    gen_loss_case("readme_n40", syn.DenoiserConfig(**{**syn.README.__dict__, "mask_prob": 0.15, "num_steps": 2000}),
                  [(8, 32), (6, 27)], seed=10)


if __name__ == "__main__":
    main()
