"""Time individual kernels at the north-star size (B=8, N=512, paper dims) without running a full bench:
runs one triangle-attention / triangle-multiplication op to fill the workspace, then times the named
kernels alone through prd_profile_kernel (CUDA events on the launch stream).
Usage: python tools/kprof.py [triattn_flash trimul_gemm pair_bias ...] [--B 8] [--N 512] [--iters 10]"""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from protein_redesign_b200 import _lib, ops  # noqa: E402
from protein_redesign_b200 import synthetic as syn  # noqa: E402
from protein_redesign_b200.model import ProteinReDiffModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernels", nargs="*", default=["triattn_flash"])
    ap.add_argument("--B", type=int, default=8)
    ap.add_argument("--N", type=int, default=512)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--pad", type=int, default=0)
    a = ap.parse_args()
    cfg, dev = syn.PAPER, torch.device("cuda:0")
    m = ProteinReDiffModel(cfg)
    m.load_state_dict(syn.make_state_dict(cfg, 0), strict=True)
    m = m.to(dev).eval()
    g = torch.Generator().manual_seed(0)
    pair = (torch.randn(a.B, a.N, a.N, cfg.pair_dim, generator=g) * 1.5 + 0.3).to(dev)
    mask = torch.ones(a.B, a.N, device=dev)
    if a.pad:
        mask[:, a.N - a.pad:] = 0
    blk = m.Denoiser.folding_blocks[0]
    lib = _lib.load()
    d = ops.make_dims(cfg, a.B, a.N)
    ops.reserve_workspace(cfg, a.B, a.N, dev)
    ws = _lib.Workspace.reserve(dev, max(_lib.workspace_bytes(op, d) for op in _lib.OPS))
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    lib.prd_profile_kernel.restype = ctypes.c_int
    res = {}
    single = torch.randn(a.B, a.N, cfg.single_dim, generator=g).to(dev)
    op_calls = {
        "op_triattn_start": lambda: blk.pair_attn_starting.apply_(cfg, pair, mask),
        "op_triattn_end": lambda: blk.pair_attn_ending.apply_(cfg, pair, mask),
        "op_trimul_out": lambda: blk.pair_mul_outgoing.apply_(cfg, pair, mask),
        "op_trimul_in": lambda: blk.pair_mul_incoming.apply_(cfg, pair, mask),
        "op_pair_fc": lambda: ops.pair_transition(cfg, pair, blk.pair_fc.packed_pair(), pair),
        "op_outer_linear": lambda: blk.outer_linear.apply_(cfg, single, pair),
        "op_single_attn": lambda: ops.single_attention(cfg, single, pair, mask, blk.single_attn.packed_single(blk.attn_bias[1]), single),
        "op_block": lambda: blk.forward_(cfg, single, pair, mask),
        "op_single_fc": lambda: ops.single_transition(cfg, single, blk.single_fc.packed_single(), single),
        "op_spattention": lambda: ops.spattention(cfg, single, pair, m.Denoiser.SPAAttnBlock.packed_weights(), out=single),
    }
    # per-step pair embedding (RBF + time + OPM) with / without the interpolated distance table
    zc = torch.randn(a.B, a.N, 3, generator=g).to(dev)
    tt = torch.full((a.B,), 17, dtype=torch.int64, device=dev)
    w = m._weights()
    oa, ob = m.Denoiser.opm.project(cfg, single, mask)
    _, (w_o, b_o) = m.Denoiser.opm.packed_weights()
    pe_out = torch.empty_like(pair)
    op_calls["op_pair_embed_lut"] = lambda: ops.pair_embed(cfg, pair, zc, mask, tt, oa, ob, w["pair_dyn"] + [w_o, b_o], pe_out,
                                                           rbf_lut=w["rbf_lut"])
    op_calls["op_pair_embed_gemm"] = lambda: ops.pair_embed(cfg, pair, zc, mask, tt, oa, ob, w["pair_dyn"] + [w_o, b_o], pe_out)
    for name in list(a.kernels):
        if name == "ops":
            a.kernels.remove("ops")
            a.kernels += [k for k in op_calls if k not in a.kernels]
    for name in a.kernels:
        if name in op_calls:  # whole module-level op (several kernels), CUDA events on the current stream
            f = op_calls[name]
            f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.iters):
                f()
            e1.record()
            torch.cuda.synchronize()
            res[name] = round(e0.elapsed_time(e1) / a.iters, 4)
            continue
        if name == "triattn_flash":
            blk.pair_attn_starting.apply_(cfg, pair, mask)
            aux = mask
        elif name == "trimul_gemm":
            blk.pair_mul_outgoing.apply_(cfg, pair, mask)
            aux = None
        else:
            aux = pair
        torch.cuda.synchronize()
        ms = ctypes.c_float(0.0)
        rc = lib.prd_profile_kernel(name.encode(), ctypes.byref(d), ctypes.c_void_p(ws.data_ptr()),
                                    ctypes.c_size_t(ws.numel()), ctypes.c_void_p(aux.data_ptr() if aux is not None else 0),
                                    a.iters, ctypes.byref(ms), stream)
        if rc != 0:
            raise RuntimeError(_lib.last_error())
        res[name] = round(ms.value, 4)
    print(json.dumps({"B": a.B, "N": a.N, "ms": res, "env": {k: v for k, v in os.environ.items() if k.startswith("PRD_")}}))


if __name__ == "__main__":
    main()
