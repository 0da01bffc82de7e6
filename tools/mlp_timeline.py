"""Per-role timeline of CTA 0 of the fused-MLP kernel (diagnostic; -DSVIT_TIMELINE build, svit_debug_mlp_timeline hook).
usage: SVIT_LIB=svit_b200/libsvit_sm100_tl.so python tools/mlp_timeline.py [M] [ln|res]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops, _lib

M = int(sys.argv[1]) if len(sys.argv) > 1 else 64 * 25153
mode = sys.argv[2] if len(sys.argv) > 2 else "ln"
Cn, Hd = 96, 384
x = torch.randn(M, Cn, device="cuda").bfloat16()
res = torch.randn(M, Cn, device="cuda").bfloat16()
w1 = torch.randn(Hd, Cn, device="cuda") * Cn ** -0.5
w2 = torch.randn(Cn, Hd, device="cuda") * Hd ** -0.5
b1, b2 = torch.randn(Hd, device="cuda"), torch.randn(Cn, device="cuda")
g, bt = torch.ones(Cn, device="cuda"), torch.zeros(Cn, device="cuda")
run = (lambda: ops.mlp_fused(x, w1, b1, w2, b2, x, ln=(g, bt, 1e-6))) if mode == "ln" else (lambda: ops.mlp_fused(x, w1, b1, w2, b2, res))
with torch.no_grad():
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    print(f"M {M} {mode}: {e0.elapsed_time(e1) / 5 * 1e3:.1f} us")
    buf = torch.zeros(3 * 4096 * 2, dtype=torch.int64, device="cuda")
    hook = _lib.lib().svit_debug_mlp_timeline
    hook.argtypes = [ctypes.c_void_p]
    hook(buf.data_ptr()); run(); torch.cuda.synchronize(); hook(None)
b = buf.cpu().reshape(3, 4096, 2)
t0 = min(int(b[r, 0, 1]) for r in range(3) if int(b[r, 0, 1]) > 0)
names = {0: "prod", 1: "mma", 2: "epi"}
ev = []
for r in range(3):
    for i in range(4096):
        tag, t = int(b[r, i, 0]), int(b[r, i, 1])
        if t == 0: break
        ev.append((t - t0, names[r], tag))
ev.sort()
for t, nm, tag in ev[:260]: print(f"{t:9d} {nm:5s} {tag}")
