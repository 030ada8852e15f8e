"""Pins the CPU oracle (oracle/denoiser_ref.py) against outputs of the unmodified reference
(tests/golden/*.npz, produced by tests/golden/make_golden.py).  CPU only."""
import dataclasses

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import denoiser_ref as ref
from protein_redesign_b200 import synthetic as syn

STEP_CASES = {
    "tiny_probes": dict(cfg=syn.TINY, sizes=[(5, 14), (3, 9)], seed=1, n_total=22, two_chains=True),
    "readme_n40": dict(cfg=syn.README, sizes=[(8, 32), (6, 27)], seed=2),
    "paper_n72": dict(cfg=syn.PAPER, sizes=[(12, 60), (9, 50)], seed=3),
    "paper_n128": dict(cfg=syn.PAPER, sizes=[(16, 112)], seed=4),
}


def _build(case):
    cfg, seed = case["cfg"], case["seed"]
    sd = syn.make_state_dict(cfg, seed)
    batch = syn.make_batch(cfg, case["sizes"], seed=seed, n_total=case.get("n_total"),
                           two_chains=case.get("two_chains", False))
    return cfg, seed, sd, batch


@pytest.mark.parametrize("tag", list(STEP_CASES))
def test_step_matches_reference(tag):
    cfg, seed, sd, batch = _build(STEP_CASES[tag])
    gold = load_golden(f"step_{tag}.npz")
    # the fixtures only hold outputs: make sure we regenerated the very same inputs
    assert syn.checksum(sd) == str(gold["weights_checksum"])
    assert syn.checksum(batch) == str(gold["batch_checksum"])
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    assert syn.checksum([z, seq_t, mask, t]) == str(gold["inputs_checksum"])

    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    # index / mask path: bit exact
    assert np.array_equal(pb["residue_extra_mask"].numpy(), gold["keep_mask"])
    assert np.array_equal(pb["residue_inv_extra_mask"].numpy(), gold["drop_mask"])
    assert np.array_equal(pb["residue_one_hot"].numpy(), gold["residue_one_hot"])
    assert np.array_equal(pb["residue_type_masked"].numpy(), gold["residue_type_masked"])
    assert np.array_equal(pb["x"].numpy(), gold["x"])

    probes = {}
    with torch.inference_mode():
        noise, seq = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t, probe=lambda n, v: probes.__setitem__(n, v))
    # fp32 restatement vs fp32 reference: same math, op order may differ slightly
    assert rel_l2(noise, torch.from_numpy(gold["noise_pred"])) < 2e-5
    assert rel_l2(seq, torch.from_numpy(gold["seq_pred"])) < 2e-5
    # padded tokens: noise_pred exactly zero
    assert float((noise * (1 - mask).unsqueeze(-1)).abs().max()) == 0.0


def test_module_probes_match_reference():
    """Every FoldingBlock sub-update, OPM, SPA and the trunk output vs forward-hook captures."""
    cfg, seed, sd, batch = _build(STEP_CASES["tiny_probes"])
    gold = load_golden("step_tiny_probes.npz")
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    m2 = mask.unsqueeze(-1) * mask.unsqueeze(-2)
    H = cfg.num_heads
    with torch.inference_mode():
        single = ref.embed_single(sd, pb, seq_t)
        pair = ref.embed_pair_static(sd, pb, cfg.max_bond_distance, cfg.max_relpos) \
            + ref.embed_pair_dynamic(sd, z, t, mask, cfg.num_steps)
        g = lambda n: torch.from_numpy(gold["probe:" + n])
        opm = ref.outer_product_update(sd, single, mask)
        assert rel_l2(opm, g("Denoiser.opm")) < 1e-5
        pair = pair + m2.unsqueeze(-1) * opm
        single = ref.single_pair_attention(sd, single, pair, H)
        assert rel_l2(single, g("Denoiser.SPAAttnBlock")) < 1e-5
        for k in range(cfg.num_blocks):
            p = f"Denoiser.folding_blocks.{k}."
            bias = ref.attn_bias_from_pair(sd, p, pair)
            assert rel_l2(bias, g(p + "attn_bias")) < 1e-5
            d = ref.gated_attention(sd, p + "single_attn.", single, mask, H, bias)
            assert rel_l2(d, g(p + "single_attn")) < 1e-5
            single = single + d
            d = ref.transition(sd, p + "single_fc.", single)
            assert rel_l2(d, g(p + "single_fc")) < 1e-5
            single = single + d
            d = ref.outer_linear(sd, p + "outer_linear.", single)
            assert rel_l2(d, g(p + "outer_linear")) < 1e-5
            pair = pair + d
            for name, fn in (("pair_mul_outgoing", lambda x: ref.triangle_multiplication(sd, p + "pair_mul_outgoing.", x, m2, "outgoing")),
                             ("pair_mul_incoming", lambda x: ref.triangle_multiplication(sd, p + "pair_mul_incoming.", x, m2, "incoming")),
                             ("pair_attn_starting", lambda x: ref.triangle_attention(sd, p + "pair_attn_starting.", x, m2, H, "starting")),
                             ("pair_attn_ending", lambda x: ref.triangle_attention(sd, p + "pair_attn_ending.", x, m2, H, "ending")),
                             ("pair_fc", lambda x: ref.transition(sd, p + "pair_fc.", x))):
                d = fn(pair)
                assert rel_l2(d, g(p + name)) < 1e-5, name
                pair = pair + d
        pair = 0.5 * (pair + pair.transpose(1, 2))
        assert rel_l2(single, g("Denoiser:0")) < 1e-5
        assert rel_l2(pair, g("Denoiser:1")) < 1e-5
        w = torch.nn.functional.linear(torch.relu(torch.nn.functional.linear(
            torch.nn.functional.layer_norm(pair, pair.shape[-1:]), sd["weight_radial.1.weight"],
            sd["weight_radial.1.bias"])), sd["weight_radial.3.weight"])
        assert rel_l2(w, g("weight_radial")) < 1e-5


@pytest.mark.parametrize("tag,cfg_over,sizes,seed", [
    ("tiny_T8", dict(num_steps=8, mask_prob=0.3), [(5, 14), (3, 9)], 5),
    ("tiny_T6_cos", dict(num_steps=6, mask_prob=1.0, diffusion_schedule="cosine"), [(4, 12)], 6),
    ("tiny_T50", dict(num_steps=50, mask_prob=0.3), [(5, 14), (3, 9)], 8),  # north_star: 50-step fixed-noise trajectory
])
def test_sampler_matches_reference(tag, cfg_over, sizes, seed):
    cfg = dataclasses.replace(syn.TINY, **cfg_over)
    gold = load_golden(f"sample_{tag}.npz")
    sd = syn.make_state_dict(cfg, seed)
    batch = syn.make_batch(cfg, sizes, seed=seed)
    assert syn.checksum(sd) == str(gold["weights_checksum"])
    assert syn.checksum(batch) == str(gold["batch_checksum"])
    tab = ref.schedule_tables(cfg.num_steps, cfg.diffusion_schedule)
    for k, v in tab.items():
        assert np.array_equal(v.numpy(), gold["sched:" + k]), k  # schedule: bit exact
    g = torch.Generator().manual_seed(seed + 31337)
    n_draws = [0]

    def randn_like(x):
        n_draws[0] += 1
        return torch.randn(x.shape, generator=g, dtype=x.dtype)

    torch.manual_seed(seed)
    with torch.inference_mode():
        pos, logits = ref.sample(sd, cfg, batch, randn_like=randn_like)
    assert n_draws[0] == int(gold["num_draws"])
    assert rel_l2(pos, torch.from_numpy(gold["pos"])) < 1e-4
    assert rel_l2(logits, torch.from_numpy(gold["logits"])) < 1e-4
    rmsd = float((pos - torch.from_numpy(gold["pos"])).square().sum(-1).mean().sqrt())
    assert rmsd < 1e-3, rmsd  # Angstrom


LOSS_CASES = {
    "tiny": (dataclasses.replace(syn.TINY, mask_prob=0.15), [(5, 14), (3, 9)], 9, dict(n_total=22, two_chains=True)),
    "readme_n40": (dataclasses.replace(syn.README, mask_prob=0.15, num_steps=2000), [(8, 32), (6, 27)], 10, {}),
}


@pytest.mark.parametrize("tag", list(LOSS_CASES))
def test_training_loss_matches_reference(tag):
    """a18: prepare_batch + randint + q() + the three loss terms vs the reference's training_step, and the gradient of
    the loss with respect to the network outputs vs the reference's autograd."""
    cfg, sizes, seed, kw = LOSS_CASES[tag]
    gold = load_golden(f"loss_{tag}.npz")
    sd = syn.make_state_dict(cfg, seed)
    batch = syn.make_batch(cfg, sizes, seed=seed, with_positions=True, **kw)
    assert syn.checksum(sd) == str(gold["weights_checksum"])
    assert syn.checksum(batch) == str(gold["batch_checksum"])
    g = torch.Generator().manual_seed(seed + 4242)
    outs = []
    torch.manual_seed(seed)
    loss, diff, t = ref.training_loss(sd, cfg, batch, randn_like=lambda x: torch.randn(x.shape, generator=g, dtype=x.dtype),
                                      outputs=outs)
    assert np.array_equal(t.numpy(), gold["t"])  # index path: bit exact
    assert rel_l2(outs[0], torch.from_numpy(gold["noise_pred"])) < 2e-5
    assert rel_l2(outs[1], torch.from_numpy(gold["seq_pred"])) < 2e-5
    assert rel_l2(diff, torch.from_numpy(gold["diff_loss"])) < 1e-5
    assert abs(float(loss.detach()) - float(gold["loss"])) < 1e-5 * abs(float(gold["loss"]))
    loss.backward()
    assert rel_l2(outs[0].grad, torch.from_numpy(gold["d_noise_pred"])) < 1e-5
    assert rel_l2(outs[1].grad, torch.from_numpy(gold["d_seq_pred"])) < 1e-5


def test_oracle_invariants():
    """Properties the reference satisfies (SURVEY §4 probe): E(3) equivariance of noise_pred,
    invariance of seq_pred, masked zero mean."""
    cfg, seed, sd, batch = _build(STEP_CASES["tiny_probes"])
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    g = torch.Generator().manual_seed(0)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    shift = torch.randn(1, 1, 3, generator=g)
    with torch.inference_mode():
        n0, s0 = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
        n1, s1 = ref.denoiser_step(sd, cfg, pb, z @ q + shift, seq_t, mask, t)
    assert rel_l2(n1, n0 @ q) < 1e-4
    assert rel_l2(s1, s0) < 1e-4
    mean = (mask.unsqueeze(-1) * n0).sum(1) / mask.sum(1, keepdim=True)
    assert float(mean.abs().max()) < 1e-6


@pytest.mark.parametrize("tag", list(LOSS_CASES))
def test_backward_matches_reference(tag):
    """Ground truth for the backward pass (built next): autograd through the oracle reproduces the
    parameter gradients of the reference's training_step (through its per-block checkpointing) -- per parameter the L2
    norm and the projection on a seeded random direction."""
    cfg, sizes, seed, kw = LOSS_CASES[tag]
    gold = load_golden(f"loss_{tag}.npz")
    sd = {k: (v.clone().requires_grad_() if v.is_floating_point() else v) for k, v in syn.make_state_dict(cfg, seed).items()}
    batch = syn.make_batch(cfg, sizes, seed=seed, with_positions=True, **kw)
    g = torch.Generator().manual_seed(seed + 4242)
    torch.manual_seed(seed)
    loss, _, _ = ref.training_loss(sd, cfg, batch, randn_like=lambda x: torch.randn(x.shape, generator=g, dtype=x.dtype))
    loss.backward()
    gp = torch.Generator().manual_seed(seed + 777)
    names = [str(n) for n in gold["grad_names"]]
    assert len(names) >= 100
    worst = 0.0
    for n, want_norm, want_proj in zip(names, gold["grad_norms"], gold["grad_projs"]):
        grad = sd[n].grad
        assert grad is not None, n
        d = torch.randn(grad.shape, generator=gp)
        scale = max(float(want_norm), 1e-12)
        worst = max(worst, abs(float(grad.norm()) - float(want_norm)) / scale,
                    abs(float((grad * d).sum()) - float(want_proj)) / (scale * float(d.norm())))
    assert worst < 1e-3, worst
