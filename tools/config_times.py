"""Step time of the BASELINE.json configurations other than the headline one (CUDA events, eager launches, 5 iterations
after 2 warm-ups): config 1 (README dims, B=3, N=140), config 2 (paper dims, B=1, N=300), config 5 (paper dims, B=1, N=1024).
Usage (GPU box): python tools/config_times.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from protein_redesign_b200 import synthetic as syn  # noqa: E402
from protein_redesign_b200.model import ProteinReDiffModel  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = {}
    for tag, cfg, sizes in (("config1_readme_B3_N140", syn.README, [(30, 110)] * 3), ("config2_paper_B1_N300", syn.PAPER, [(30, 270)]),
                            ("config5_paper_B1_N1024", syn.PAPER, [(1, 1023)]), ("config3_paper_B8_N512", syn.PAPER, [(32, 480)] * 8)):
        m = ProteinReDiffModel(cfg)
        m.load_state_dict(syn.make_state_dict(cfg, 0), strict=True)
        m = m.to(dev).eval()
        batch = syn.make_batch(cfg, sizes, seed=0)
        z, seq_t, mask, t = syn.make_step_inputs(batch, cfg.num_steps, 0)
        torch.manual_seed(0)
        db = m.prepare_batch({k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in batch.items()})
        z, seq_t, mask, t = z.to(dev), seq_t.to(dev), mask.to(dev), t.to(dev)
        with torch.inference_mode():
            for _ in range(2):
                n, s = m.sample_step(db, z, seq_t, mask, t)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                n, s = m.sample_step(db, z, seq_t, mask, t)
            e1.record()
            torch.cuda.synchronize()
        out[tag] = {"ms_per_step": round(e0.elapsed_time(e1) / 5, 3), "finite": bool(torch.isfinite(n).all() and torch.isfinite(s).all())}
        del m, db
        torch.cuda.empty_cache()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
