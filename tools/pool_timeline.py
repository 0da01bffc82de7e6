"""Timeline of thread 0 of CTA 0 of the channel-pair pool kernel (needs a -DSVIT_TIMELINE build)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import _lib
from svit_b200.ops import _call, _stream, tap_fractions, BF16
B, h, T, H, W, O, s = 64, 4, 8, 14, 14, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1
N = 1 + T * H * W + O
Ho = (H - 1) // s + 1
No = 1 + T * Ho * Ho + O
w = torch.randn(96, 27, device="cuda"); g = torch.ones(96, device="cuda"); b_ = torch.zeros(96, device="cuda")
frac = tap_fractions(s, "cuda")
out = torch.empty(B, h, No, 96, device="cuda", dtype=torch.bfloat16)
packed = torch.randn(B, N, 3 * h * 96, device="cuda").bfloat16()
run = lambda: _call("svit_pool_ln_fwd", packed.data_ptr(), N * 3 * h * 96, 3 * h * 96, 96, w.data_ptr(), frac.data_ptr(), g.data_ptr(),
                    b_.data_ptr(), out.data_ptr(), B, h, T, H, W, O, s, 1e-6, BF16, _stream())
for _ in range(3): run()
buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
hook = _lib.lib().svit_debug_pool_timeline
hook.argtypes = [ctypes.c_void_p]
hook(buf.data_ptr()); run(); torch.cuda.synchronize(); hook(None)
b = buf.cpu().reshape(1024, 2)
t0 = int(b[0, 1])
for i in range(1024):
    if int(b[i, 1]) == 0: break
    print(f"{int(b[i,1]) - t0:8d} {int(b[i,0])}")
