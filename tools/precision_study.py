"""Which op's fp16 operand rounding dominates the end-to-end error?  (CPU study, documents the
tolerance / operand-splitting choices in DESIGN.md.)

Emulates "fp16 operands, fp32 accumulate" inside the CPU oracle per call site (by rounding the
operands of F.linear / matmul / einsum to fp16 while a given oracle function is on the stack) and
reports the relative L2 error of noise_pred / seq_pred against the unmodified fp32 oracle.
"""
import contextlib
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import denoiser_ref as ref
from protein_redesign_b200 import synthetic as syn

h = lambda t: t.half().float()
ACTIVE = set()
STACK = []
orig_linear, orig_matmul, orig_einsum = F.linear, torch.matmul, torch.einsum


def on():
    return any(s in ACTIVE for s in STACK) or "all" in ACTIVE


def linear(x, w, b=None):
    if on() and w.shape[0] > 4:   # the c_z -> H bias projections run in fp32 SIMT in the product
        return orig_linear(h(x), h(w), b)
    return orig_linear(x, w, b)


def matmul(a, b):
    return orig_matmul(h(a), h(b)) if on() else orig_matmul(a, b)


def einsum(eq, *ops):
    return orig_einsum(eq, *[h(o) for o in ops]) if on() else orig_einsum(eq, *ops)


def wrap(name):
    f = getattr(ref, name)
    def g(*a, **k):
        STACK.append(name)
        try:
            return f(*a, **k)
        finally:
            STACK.pop()
    setattr(ref, name, g)


SITES = ["embed_single", "embed_pair_dynamic", "outer_product_update", "single_pair_attention", "gated_attention",
         "transition", "outer_linear", "triangle_multiplication", "coord_head", "seq_head"]


def main():
    cfg, sizes, seed = syn.PAPER, [(16, 112)], 4
    sd = syn.make_state_dict(cfg, seed)
    batch = syn.make_batch(cfg, sizes, seed=seed)
    z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
    torch.manual_seed(seed)
    pb = ref.prepare_batch(batch, cfg.mask_prob)
    with torch.inference_mode():
        n0, s0 = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
    ref.F.linear = linear
    ref.torch.matmul = matmul
    ref.torch.einsum = einsum
    for s in SITES:
        wrap(s)
    rel = lambda a, b: float((a - b).norm() / b.norm())
    for active in [["all"]] + [[s] for s in SITES] + [["embed_single", "embed_pair_dynamic", "outer_product_update"]]:
        ACTIVE.clear(); ACTIVE.update(active)
        with torch.inference_mode():
            n1, s1 = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
        print(f"{'+'.join(active):60s} noise {rel(n1, n0):.2e}  seq {rel(s1, s0):.2e}")


if __name__ == "__main__":
    main()
