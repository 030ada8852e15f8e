"""One tf32 GEMM shape a few times (for ncu): python tools/gemm_one.py [M N K]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from protein_redesign_b200 import _lib  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (197192, 256, 64)
dev = torch.device("cuda", 0)
A, B, C = torch.randn(M, K, device=dev), torch.randn(N, K, device=dev), torch.empty(M, N, device=dev)
for _ in range(3):
    _lib.gemm_f16(A, B, C, round_tf32=True)
torch.cuda.synchronize()
