"""CPU-only checks: the C-ABI library loads and exports every symbol include/prd_denoiser.h declares,
the Python mirror loads reference-shaped state-dicts strictly, host logic (batch preparation,
schedule) is bit exact against the oracle / goldens, and there is no CPU fallback."""
import dataclasses
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from oracle import denoiser_ref as ref
from protein_redesign_b200 import _lib, ops
from protein_redesign_b200 import synthetic as syn
from protein_redesign_b200.model import ProteinReDiffModel
from protein_redesign_b200.modules import Denoiser, TriangleMultiplication


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "prd_denoiser.h")).read()
    syms = set(re.findall(r"\b(prd_[a-z0-9_]+)\s*\(", text))
    syms = {s for s in syms if "##" not in s}
    for op in re.findall(r"^PRD_DECLARE_OP\((\w+)\)", text, flags=re.M):
        syms.add(f"prd_{op}_fwd")
        syms.add(f"prd_{op}_workspace_bytes")
    return syms, re.findall(r"^PRD_DECLARE_OP\((\w+)\)", text, flags=re.M)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    syms, header_ops = _declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in sorted(syms) if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(header_ops) == sorted(_lib.OPS)
    assert lib.prd_version() == 2


def test_workspace_queries_run_without_gpu():
    d = ops.make_dims(syn.PAPER, 8, 512)
    sizes = {op: _lib.workspace_bytes(op, d) for op in _lib.OPS}
    assert all(v >= 256 for v in sizes.values())
    # triangle attention holds q,k,g,og [R,64] + vt planes in fp16
    R = 8 * 512 * 512
    assert sizes["triangle_attention"] >= 4 * R * 64 * 2
    d2 = ops.make_dims(syn.PAPER, 1, 1024)
    assert _lib.workspace_bytes("triangle_multiplication", d2) > 0


@pytest.mark.parametrize("cfg", [syn.PAPER, syn.README])
def test_state_dict_contract(cfg):
    model = ProteinReDiffModel(cfg)
    sd = syn.make_state_dict(cfg, 0)
    assert sorted(model.state_dict().keys()) == sorted(sd.keys())
    model.load_state_dict(sd, strict=True)
    assert {k: tuple(v.shape) for k, v in model.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
    # Denoiser accepts a Mapping like the reference (modules.py:353)
    Denoiser(dataclasses.asdict(cfg))


def test_invalid_modes_raise_like_the_reference():
    with pytest.raises(ValueError):
        TriangleMultiplication(64, "sideways")
    from protein_redesign_b200.modules import Linear, TriangleAttention
    with pytest.raises(ValueError):
        TriangleAttention(64, 16, 4, "middle")
    with pytest.raises(ValueError):
        Linear(4, 4, init="nope")


def test_no_cpu_fallback():
    model = ProteinReDiffModel(syn.README)
    model.load_state_dict(syn.make_state_dict(syn.README, 0))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.symmetrize(syn.README, torch.zeros(1, 8, 8, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.Denoiser.folding_blocks[0].pair_fc.packed_pair()  # CPU parameters cannot be packed


def test_prepare_batch_matches_oracle_bit_exact():
    cfg = dataclasses.replace(syn.README, mask_prob=0.4)
    model = ProteinReDiffModel(cfg)
    batch = syn.make_batch(cfg, [(8, 32), (6, 27)], seed=2, with_positions=True)
    torch.manual_seed(123)
    want = ref.prepare_batch(batch, cfg.mask_prob)
    torch.manual_seed(123)
    got = model.prepare_batch({k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
    for k in ("residue_extra_mask", "residue_inv_extra_mask", "residue_one_hot", "residue_type_masked", "x",
              "residue_and_atom_mask", "residue_esm"):
        assert torch.equal(got[k], want[k]), k
    assert int(got["residue_inv_extra_mask"].sum()) == int(int(batch["residue_mask"].sum()) * 0.4)


@pytest.mark.parametrize("tag,over", [("tiny_T8", dict(num_steps=8)),
                                      ("tiny_T6_cos", dict(num_steps=6, diffusion_schedule="cosine"))])
def test_schedule_tables_match_reference_golden(tag, over):
    gold = load_golden(f"sample_{tag}.npz")
    model = ProteinReDiffModel(dataclasses.replace(syn.TINY, **over))
    model.run_setup_schedule()
    for k in ("betas", "alphas", "alphas_cumprod", "sqrt_betas", "sqrt_alphas", "sqrt_one_minus_alphas_cumprod"):
        assert np.array_equal(getattr(model, k).numpy(), gold["sched:" + k]), k
    coef = model._coef.numpy()
    assert np.array_equal(coef[:, 2], gold["sched:sqrt_betas"])


def test_checkpoint_round_trip_and_optimizer_glue(tmp_path):
    """Lightning-free load_from_checkpoint (generate.py:103-107): hyper-parameters + overrides -> constructor, strict
    state-dict, EMA hook; configure_optimizers mirrors reference model.py:203-213."""
    import torch
    from protein_redesign_b200 import synthetic as syn
    from protein_redesign_b200.model import ProteinReDiffModel
    cfg = syn.TINY
    m = ProteinReDiffModel(cfg)
    m.load_state_dict(syn.make_state_dict(cfg, 3), strict=True)
    ckpt = {"state_dict": m.state_dict(), "hyper_parameters": vars(cfg.to_namespace())}
    m.on_save_checkpoint(ckpt)
    assert "ema_state_dict" in ckpt
    path = tmp_path / "model.ckpt"
    torch.save(ckpt, path)
    m2 = ProteinReDiffModel.load_from_checkpoint(str(path), num_steps=20, mask_prob=0.5)
    assert (m2.num_steps, m2.mask_prob) == (20, 0.5)
    for (k, a), b in zip(m.state_dict().items(), m2.state_dict().values()):
        assert torch.equal(a, b), k
    opt = m2.configure_optimizers()
    assert isinstance(opt["optimizer"], torch.optim.Adam) and opt["lr_scheduler"]["interval"] == "step"
    assert abs(opt["optimizer"].param_groups[0]["lr"] - m2.learning_rate / m2.warmup_steps) < 1e-12


def test_training_step_has_no_cpu_fallback():
    """No CPU fallback: with or without autograd, training_step on CPU tensors fails loudly in the CUDA-only ops."""
    import pytest
    import torch
    from protein_redesign_b200 import synthetic as syn
    from protein_redesign_b200.model import ProteinReDiffModel
    m = ProteinReDiffModel(syn.TINY)
    batch = syn.make_batch(syn.TINY, [(3, 9)], seed=0, with_positions=True)
    with pytest.raises(RuntimeError):
        m.training_step(dict(batch), 0)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m.training_step(dict(batch), 0)


def test_collate_fn_matches_reference_golden():
    """predict.collate_fn (reference data.py:80-142) bit-exactly against the unmodified reference (tests/golden/collate.npz,
    generated by make_golden.py)."""
    from protein_redesign_b200.predict import InferenceDataset, RepeatDataset, collate_fn
    gold = load_golden("collate.npz")
    items = [syn.make_complex(syn.TINY, na, nr, seed=60 + i) for i, (na, nr) in enumerate([(5, 9), (3, 14), (7, 4)])]
    batch = collate_fn(items)
    tensors = {k: v for k, v in batch.items() if isinstance(v, torch.Tensor)}
    assert sorted(tensors) == sorted(k for k in gold if k != "mol_lists")
    for k, v in tensors.items():
        assert v.dtype == torch.from_numpy(gold[k]).dtype, k
        assert np.array_equal(v.numpy(), gold[k]), k
    assert ",".join(batch["ligand_mol"]) == str(gold["mol_lists"][0]) and ",".join(batch["protein_mol"]) == str(gold["mol_lists"][1])
    # token order: ligand atoms, residues, padding; residue types shifted by +1
    assert int(batch["residue_type"].min()) == 0 and int(batch["residue_type"].max()) <= 20
    assert len(RepeatDataset(items[0], 5)) == 5 and RepeatDataset(items[0], 5)[3] is items[0]
    assert len(InferenceDataset(items, 0)) == 3 and InferenceDataset(items, 0)[1] is items[1]


def test_training_tape_policy(monkeypatch):
    """The forward keeps the input of every residual update when that fits (a few GB at training sizes) and falls back to
    the reference's one-checkpoint-per-block scheme otherwise; PRD_TAPE overrides either way."""
    import torch
    from protein_redesign_b200 import autograd as ag
    from protein_redesign_b200 import synthetic as syn
    monkeypatch.delenv("PRD_TAPE", raising=False)
    small = torch.empty(2, 314, 314, 64, device="meta")   # 50 MB: 24 copies = 1.2 GB
    huge = torch.empty(8, 2048, 2048, 64, device="meta")  # 8.6 GB: 24 copies do not fit the 32 GB tape budget
    assert ag._tape_per_op(syn.PAPER, small) is True
    assert ag._tape_per_op(syn.PAPER, huge) is False
    monkeypatch.setenv("PRD_TAPE", "blocks")
    assert ag._tape_per_op(syn.PAPER, small) is False
    monkeypatch.setenv("PRD_TAPE", "ops")
    assert ag._tape_per_op(syn.PAPER, huge) is True
