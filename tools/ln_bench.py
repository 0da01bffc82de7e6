"""LayerNorm forward (bf16) at the model's four widths: us per launch and GB/s (read + write)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops
for rows, C in ((104512, 384), (405568, 192), (1609792, 96), (29248, 768)):
    x = torch.randn(rows, C, device="cuda").bfloat16()
    g = torch.randn(C, device="cuda"); b = torch.randn(C, device="cuda")
    with torch.no_grad():
        for _ in range(3): ops.layer_norm(x, g, b)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.layer_norm(x, g, b); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"rows {rows} C {C}: {ms*1e3:.1f} us  {4.0*rows*C/ms/1e6:.0f} GB/s")
