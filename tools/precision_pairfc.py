"""Does pair_fc need BOTH weight halves (hi + lo fp16) on both of its layers?  CPU emulation of the kernel's arithmetic
(fp16 A operand after LayerNorm, fp16 hidden tile after ReLU, fp32 accumulate) inside the fp32 oracle, everything else exact:
reports the step error (rel. L2 of noise_pred / seq_pred vs the fp32 oracle) contributed by pair_fc alone for the four
combinations.  Dropping a lo half would free 32 KB of shared memory in pair_transition_ws (DESIGN §0 row f) and a quarter of
its UMMAs.  Usage: python tools/precision_pairfc.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import denoiser_ref as ref
from protein_redesign_b200 import synthetic as syn

h = lambda t: t.half().float()
MODE = {"w1_split": True, "w2_split": True, "on": False}
orig_transition = ref.transition


def transition(sd, prefix, x):
    if not MODE["on"] or "pair_fc" not in prefix:
        return orig_transition(sd, prefix, x)
    w1, b1, w2, b2 = sd[prefix + "1.weight"], sd[prefix + "1.bias"], sd[prefix + "3.weight"], sd[prefix + "3.bias"]
    a = h(F.layer_norm(x, x.shape[-1:]))
    w1e = w1 if MODE["w1_split"] else h(w1)   # hi + lo reproduces the fp32 weight to 2^-22
    hid = h(F.relu(F.linear(a, w1e, b1)))
    w2e = w2 if MODE["w2_split"] else h(w2)
    return F.linear(hid, w2e, b2)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref.transition = transition
    rel = lambda a, b: float((a - b).norm() / b.norm())
    for cfg, sizes, seed in ((syn.PAPER, [(16, 112)], 4), (syn.PAPER, [(12, 60), (9, 50)], 3)):
        sd = syn.make_state_dict(cfg, seed)
        batch = syn.make_batch(cfg, sizes, seed=seed)
        z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, seed)
        torch.manual_seed(seed)
        pb = ref.prepare_batch(batch, cfg.mask_prob)
        with torch.inference_mode():
            MODE["on"] = False
            n0, s0 = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
            MODE["on"] = True
            for w1s in (True, False):
                for w2s in (True, False):
                    MODE["w1_split"], MODE["w2_split"] = w1s, w2s
                    n1, s1 = ref.denoiser_step(sd, cfg, pb, z, seq_t, mask, t)
                    print(f"N={mask.shape[1]} B={mask.shape[0]} W1 {'hi+lo' if w1s else 'hi   '} W2 {'hi+lo' if w2s else 'hi   '}: "
                          f"noise {rel(n1, n0):.2e} seq {rel(s1, s0):.2e}")


if __name__ == "__main__":
    main()
