"""Per-role timeline of CTA 0 of the TMA-epilogue GEMM (diagnostic; uses the svit_debug_gemm_timeline hook).
usage: python tools/gemm_timeline.py M N K [gelu|res|none] [first last]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes
import torch
from svit_b200 import ops, _lib

M, N, K = (int(x) for x in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "none"
A = torch.randn(M, K, device="cuda").bfloat16()
W = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
res = torch.randn(M, N, device="cuda").bfloat16() if mode == "res" else None
kw = dict(bias=bias)
if mode == "gelu": kw["act"] = 1
if mode == "res": kw.update(residual=res, ldr=N)
for _ in range(3): ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, impl=ops.IMPL_TC, **kw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, impl=ops.IMPL_TC, **kw)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{M}x{N}x{K} {mode}: {ms*1e3:.1f} us  {2*M*N*K/ms/1e9:.0f} TF/s")
buf = torch.zeros(3 * 4096 * 2, dtype=torch.int64, device="cuda")
hook = _lib.lib().svit_debug_gemm_timeline
hook.argtypes = [ctypes.c_void_p]
hook(buf.data_ptr())
ops.gemm(A, W, out, M, N, K, K, K, N, 0, 1, impl=ops.IMPL_TC, **kw)
torch.cuda.synchronize()
hook(None)
b = buf.cpu().reshape(3, 4096, 2)
if not any(int(b[r, 0, 1]) > 0 for r in range(3)):
    print("no timeline events: build with SVIT_NVCC_EXTRA=-DSVIT_TIMELINE"); sys.exit(0)
t0 = min(int(b[r, 0, 1]) for r in range(3) if int(b[r, 0, 1]) > 0)
names = {0: "prod", 1: "mma", 2: "epi"}
ev = []
for r in range(3):
    for i in range(4096):
        tag, t = int(b[r, i, 0]), int(b[r, i, 1])
        if t == 0: break
        ev.append((t - t0, names[r], tag))
ev.sort()
lo, hi = (int(x) for x in sys.argv[5:7]) if len(sys.argv) > 6 else (0, 160)
for t, nm, tag in ev[lo:hi]:
    print(f"{t:9d} {nm:5s} {tag}")
print("last event at", ev[-1][0], "cycles;", len(ev), "events")
