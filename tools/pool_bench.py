"""Pool kernel timing vs input layout (diagnostic): packed qkv slices (token stride 3*h*96) vs contiguous tokens."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops, _lib
from svit_b200.ops import _call, _stream, tap_fractions, BF16

def bench(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

B, h, T, H, W, O = 64, 4, 8, 14, 14, 64
for (h, H, W) in ((4, 14, 14), (1, 56, 56), (8, 7, 7)):
  for s in (1, 2):
    N = 1 + T * H * W + O
    Ho = (H - 1) // s + 1
    No = 1 + T * Ho * Ho + O
    w = torch.randn(96, 27, device="cuda"); g = torch.ones(96, device="cuda"); b_ = torch.zeros(96, device="cuda")
    frac = tap_fractions(s, "cuda")
    out = torch.empty(B, h, No, 96, device="cuda", dtype=torch.bfloat16)
    packed = torch.randn(B, N, 3 * h * 96, device="cuda").bfloat16()
    contig = torch.randn(B, h, N, 96, device="cuda").bfloat16()
    def run_packed():
        _call("svit_pool_ln_fwd", packed.data_ptr(), N * 3 * h * 96, 3 * h * 96, 96, w.data_ptr(), frac.data_ptr(), g.data_ptr(),
              b_.data_ptr(), out.data_ptr(), B, h, T, H, W, O, s, 1e-6, BF16, _stream())
    def run_contig():  # batch stride h*N*96, token stride 96, head stride N*96
        _call("svit_pool_ln_fwd", contig.data_ptr(), h * N * 96, 96, N * 96, w.data_ptr(), frac.data_ptr(), g.data_ptr(),
              b_.data_ptr(), out.data_ptr(), B, h, T, H, W, O, s, 1e-6, BF16, _stream())
    by = (B * h * N * 96 + B * h * No * 96) * 2
    tp, tc = bench(run_packed), bench(run_contig)
    print(f"h{h} {H}x{W} s{s}: packed {tp*1e3:7.1f} us {by/tp/1e6:6.0f} GB/s | contiguous {tc*1e3:7.1f} us {by/tc/1e6:6.0f} GB/s")
# plain strided copies for reference
x = torch.randn(64, 1633, 3, 4, 96, device="cuda").bfloat16()
t1 = bench(lambda: x[:, :, 0, 0].contiguous()); t2 = bench(lambda: x[:, :, 0].contiguous()); t3 = bench(lambda: x.clone())
n1 = 64 * 1633 * 96 * 2 * 2
print(f"torch copy 192B slices: {n1/t1/1e6:.0f} GB/s; 768B slices: {4*n1/t2/1e6:.0f} GB/s; contiguous: {12*n1/t3/1e6:.0f} GB/s")
