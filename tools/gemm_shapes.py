"""Times of the tf32 GEMM shapes of the backward pass (B = 2, N = 314 by default) with their HBM floors.
Usage (GPU box): python tools/gemm_shapes.py [--B 2 --N 314]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protein_redesign_b200 import _lib  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--B", type=int, default=2)
    ap.add_argument("--N", type=int, default=314)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    R = a.B * a.N * a.N
    Np = (a.N + 3) // 4 * 4
    print(f"R = {R} rows, HBM peak ~6454 GB/s")
    shapes = [("pre = p Wcat^T (tri-mul)", R, 320, 64, {}), ("qkvg = x Wcat^T (tri-attn)", R, 256, 64, {}),
              ("o = xn Wo^T", R, 64, 64, {}), ("dp = dpre WcatT (K 320)", R, 64, 320, {}),
              ("dxh = dqkvg WcatT (K 256)", R, 64, 256, {}), ("h = x W1^T (pair_fc, N 256)", R, 256, 64, {}),
              ("dh = dy W2T, ReLU gate", R, 256, 64, {"mul": True})]
    for name, M, N, K, kw in shapes:
        A = torch.randn(M, K, device=dev)
        Bm = torch.randn(N, K, device=dev)
        C = torch.empty(M, N, device=dev)
        mul = torch.randn(M, N, device=dev) if kw.get("mul") else None
        ms = timeit(lambda: _lib.gemm_f16(A, Bm, C, mul=mul, mul_step=mul is not None, round_tf32=True))
        byts = 4 * (M * K + M * N * (2 if mul is not None else 1))
        print(f"{name:36s} M={M} N={N:3d} K={K:3d}  {ms:7.3f} ms  {byts / ms / 1e6:7.0f} GB/s  floor {byts / 6454e6:6.3f} ms")
    # weight-gradient reductions dW[n, k] = sum_r dY[r, n] X[r, k] (tensor-core kernel, bias gradient fused)
    import ctypes
    lib = _lib.load()
    lib.prd_dw_acc.restype = ctypes.c_int
    lib.prd_dw_acc.argtypes = [ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_int,
                               ctypes.c_int, ctypes.c_void_p, ctypes.c_longlong, ctypes.c_void_p, ctypes.c_float, ctypes.c_int, ctypes.c_void_p]
    for name, Nout, K in (("dW 64 x 64", 64, 64), ("dW 128 x 64", 128, 64), ("dW 256 x 64 (q k v g)", 256, 64), ("dW 64 x 256 (pair_fc W2)", 64, 256),
                          ("dW 256 x 64 (pair_fc W1)", 256, 64)):
        dY = torch.randn(R, Nout, device=dev)
        X = torch.randn(R, K, device=dev)
        dW = torch.zeros(Nout, K, device=dev)
        db = torch.zeros(Nout, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        ms = timeit(lambda: lib.prd_dw_acc(dY.data_ptr(), Nout, X.data_ptr(), K, R, Nout, K, dW.data_ptr(), K, db.data_ptr(), 1.0, 2, st))
        byts = 4 * R * (Nout + K)
        print(f"{name:36s} R={R}  {ms:7.3f} ms  {byts / ms / 1e6:7.0f} GB/s  floor {byts / 6454e6:6.3f} ms")
    # what bounds the wide-output shapes: pure-write bandwidth, fp16 output (half the bytes), an L2-resident problem
    big = torch.empty(R, 256, device=dev)
    ms = timeit(lambda: big.fill_(1.0))
    print(f"{'torch fill_ of [R, 256] fp32':36s} {ms:7.3f} ms  {big.numel() * 4 / ms / 1e6:7.0f} GB/s (write only)")
    src = torch.randn(R, 256, device=dev)
    ms = timeit(lambda: big.copy_(src))
    print(f"{'torch copy_ of [R, 256] fp32':36s} {ms:7.3f} ms  {2 * big.numel() * 4 / ms / 1e6:7.0f} GB/s (read + write)")
    A16 = torch.randn(R, 64, device=dev).half()
    B16 = torch.randn(256, 64, device=dev).half()
    for dt in (torch.float32, torch.float16):
        C = torch.empty(R, 256, device=dev, dtype=dt)
        ms = timeit(lambda: _lib.gemm_f16(A16, B16, C))
        print(f"{'f16 operands, C ' + str(dt)[6:]:36s} M={R} N=256 K= 64  {ms:7.3f} ms  {(A16.numel() * 2 + C.numel() * C.element_size()) / ms / 1e6:7.0f} GB/s")
    for M in (16384, 65536):
        A = torch.randn(M, 64, device=dev)
        Bm = torch.randn(256, 64, device=dev)
        C = torch.empty(M, 256, device=dev)
        ms = timeit(lambda: _lib.gemm_f16(A, Bm, C, round_tf32=True), n=50)
        print(f"{'tf32, small M (L2 resident)':36s} M={M} N=256 K= 64  {ms:7.3f} ms  {4 * (M * 64 + M * 256) / ms / 1e6:7.0f} GB/s  tiles/CTA {M / 128 * 2 / 148:.1f}")
    # plane GEMM: B * 64 planes of [N, N] x [N, N] (operands [N, Np] with the K axis padded to Np)
    nb = a.B * 64
    for n, dt in ((Np, torch.float32), (384, torch.float32), (256, torch.float32), (128, torch.float32), (a.N // 8 * 8, torch.float16), (384, torch.float16)):
        npad = (n + 3) // 4 * 4
        A = torch.randn(nb, n, npad, device=dev).to(dt)
        Bm = torch.randn(nb, n, npad, device=dev).to(dt)
        C = torch.empty(nb, n, n, device=dev)   # the wrapper writes C compactly: ldc = n
        ms = timeit(lambda: _lib.gemm_f16(A, Bm, C, round_tf32=dt == torch.float32))
        byts = (2 * A.element_size() + 4) * nb * n * npad
        print(f"{'plane GEMM ' + str(dt)[6:]:36s} nb={nb} N={n} K={npad}  {ms:7.3f} ms  {byts / ms / 1e6:7.0f} GB/s  floor {byts / 6454e6:6.3f} ms  "
              f"{2 * nb * n ** 3 / ms / 1e9:6.1f} TFLOP/s")


if __name__ == "__main__":
    main()
