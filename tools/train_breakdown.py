"""Per-op GPU time of one training step (CUDA events around every C-ABI call; serialised, so slightly pessimistic).
Usage (GPU box): python tools/train_breakdown.py [--sizes 0]"""
import argparse
import collections
import dataclasses
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from protein_redesign_b200 import _lib, synthetic as syn  # noqa: E402
from protein_redesign_b200 import autograd as ag  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, default=0)
    ap.add_argument("--kernels", action="store_true", help="per-kernel table of one step (torch.profiler / CUPTI) instead")
    ap.add_argument("--op-kernels", default=None, help="ordered kernel list of the first backward call of this op, e.g. triangle_multiplication")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    cfg = dataclasses.replace(syn.PAPER, mask_prob=0.15, num_steps=2000)
    model = bench._make_model(cfg, dev, train=True)
    host = syn.make_batch(cfg, bench.TRAIN_SIZES[a.sizes], seed=700, with_positions=True)
    grads = ag.FlatGrads(model).attach()
    to_dev = lambda: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in host.items()}
    for _ in range(2):
        ag.training_step_manual(model, to_dev(), grads)
    torch.cuda.synchronize()
    if a.op_kernels:
        from torch.profiler import ProfilerActivity, profile
        orig_bwd = _lib.call_bwd
        seen = []

        def traced(op, *args, **kw):
            if op != a.op_kernels or seen:
                return orig_bwd(op, *args, **kw)
            seen.append(1)
            torch.cuda.synchronize()
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                orig_bwd(op, *args, **kw)
                torch.cuda.synchronize()
            evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA), key=lambda e: e.time_range.start)
            print(f"kernels of one {op} backward call: {len(evs)} launches, {sum(e.device_time for e in evs) / 1e3:.3f} ms")
            for e in evs:
                print(f"  {e.device_time:8.1f} us  {e.name[:120]}")

        _lib.call_bwd = traced
        ag._lib = _lib
        ag.training_step_manual(model, to_dev(), grads)
        torch.cuda.synchronize()
        return
    if a.kernels:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            ag.training_step_manual(model, to_dev(), grads)
            torch.cuda.synchronize()
        rows = collections.defaultdict(lambda: [0, 0.0])
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                rows[ev.name][0] += 1
                rows[ev.name][1] += ev.device_time / 1e3
        total = sum(v[1] for v in rows.values())
        print(f"kernels of one training step: {sum(v[0] for v in rows.values())} launches, {total:.1f} ms of GPU time")
        for k, (n, ms) in sorted(rows.items(), key=lambda kv: -kv[1][1])[:40]:
            print(f"{k[:110]:110s} x{n:4d} {ms:8.2f} ms {100 * ms / total:5.1f} % {ms / n:7.3f}")
        return
    acc = collections.defaultdict(lambda: [0, 0.0])
    orig_call, orig_bwd = _lib.call, _lib.call_bwd

    def timed(kind, fn):
        def wrapper(op, *args, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn(op, *args, **kw)
            e1.record()
            e1.synchronize()
            acc[f"{kind}:{op}"][0] += 1
            acc[f"{kind}:{op}"][1] += e0.elapsed_time(e1)
        return wrapper

    _lib.call, _lib.call_bwd = timed("fwd", orig_call), timed("bwd", orig_bwd)
    ag._lib = _lib
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    ag.training_step_manual(model, to_dev(), grads)
    t1.record()
    torch.cuda.synchronize()
    total = sum(v[1] for v in acc.values())
    print(f"step (serialised, with per-op sync): {t0.elapsed_time(t1):.1f} ms; sum of op times {total:.1f} ms")
    for k, (n, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:24]:
        print(f"{k:36s} x{n:3d} {ms:8.2f} ms  {100 * ms / total:5.1f} %  {ms / n:7.3f} ms each")


if __name__ == "__main__":
    main()
