"""The tower GEMM against cuBLAS (torch.matmul) on the tower's own shapes, back to back in one process.

    python tools/cublas_probe.py [W=1024]        # QKV (N = 3W), fc1-shaped without activation (N = 4W), out-proj (N = W), K = W; fc2 (K = 4W)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from leaf_b200 import synth                                  # noqa: E402
from leaf_b200.tower import LeafTextTower                    # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
M = 131072
eng = LeafTextTower.random(synth.TowerCfg("ab", W, 1, W // 64, W), seed=0).leaf_engine
g = torch.Generator(device="cuda").manual_seed(0)
for name, N, K in (("qkv", 3 * W, W), ("fc1 (no act)", 4 * W, W), ("out", W, W), ("fc2 (no residual)", W, 4 * W)):
    A = (torch.randn((M, K), generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    Wt = (torch.randn((N, K), generator=g, device="cuda") * 0.03).to(torch.bfloat16)
    b = torch.randn((N,), generator=g, device="cuda")
    C = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    fns = (("ours  ", lambda: eng.gemm(A, Wt, b, 0, 0, C)), ("cublas", lambda: torch.matmul(A, Wt.T, out=C)))
    for _ in range(3):
        for _, fn in fns:
            fn()
    res = []
    for rep in range(2):
        for label, fn in fns:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            res.append(f"{label} {ms * 1e3:7.1f} us {2 * M * N * K / ms / 1e9:6.0f} TFLOP/s")
    print(f"W={W} {name:18s} M={M} N={N} K={K}:  " + "  |  ".join(res))
    del A, Wt, C
