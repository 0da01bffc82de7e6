"""Fused MLP (svit_mlp_fused) against the two-GEMM path at the stage-1 / stage-2 shapes of the batch-64 forward.
usage: python tools/mlp_bench.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for Cn, tokens in ((96, 25153), (192, 6337)):
    M, Hd = B * tokens, 4 * Cn
    x = torch.randn(M, Cn, device="cuda").bfloat16()
    res = torch.randn(M, Cn, device="cuda").bfloat16()
    w1 = (torch.randn(Hd, Cn, device="cuda") * Cn ** -0.5)
    w2 = (torch.randn(Cn, Hd, device="cuda") * Hd ** -0.5)
    b1, b2 = torch.randn(Hd, device="cuda"), torch.randn(Cn, device="cuda")
    g, bt = torch.ones(Cn, device="cuda"), torch.zeros(Cn, device="cuda")
    variants = {
        "fused, LN prologue, residual = x": ((lambda: ops.mlp_fused(x, w1, b1, w2, b2, x, ln=(g, bt, 1e-6))) if Cn == 96 else None, 2.0),
        "fused, separate residual        ": (lambda: ops.mlp_fused(x, w1, b1, w2, b2, res), 3.0),
        "LayerNorm + 2 GEMMs             ": (lambda: ops.mlp(ops.layer_norm(x, g, bt, 1e-6), w1, b1, w2, b2, x), 5.0 + 2.0 * Hd / Cn),
    }
    for name, (fn, passes) in variants.items():
        if fn is None:
            continue
        ops._MLP_FUSED["enabled"] = not name.startswith("LayerNorm")
        with torch.no_grad():
            for _ in range(3): y = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): y = fn()
            e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        fl = 4.0 * M * Cn * Hd
        by = passes * M * Cn * 2
        print(f"C {Cn} M {M} {name}: {ms*1e3:8.1f} us  {fl/ms/1e9:6.0f} TF/s  {by/ms/1e6:6.0f} GB/s of its own traffic")
    ops._MLP_FUSED["enabled"] = True
