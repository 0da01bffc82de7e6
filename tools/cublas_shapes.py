"""cuBLAS (torch.matmul / F.linear, bf16) on the stage-3 GEMM shapes next to svit_gemm: what the library reaches on the same box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops
shapes = [(104512, 1536, 384), (104512, 384, 1536), (104512, 1152, 384), (104512, 384, 384), (405568, 576, 192), (29248, 3072, 768)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, n=8):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * 0.05)
    b = torch.randn(N, device="cuda")
    wb, bb = w.bfloat16(), b.bfloat16()
    with torch.no_grad():
        t_lib = timeit(lambda: torch.nn.functional.linear(x, wb, bb))
        t_lib_nb = timeit(lambda: torch.matmul(x, wb.t()))
        t_own = timeit(lambda: ops.linear(x, w, b))
    fl = 2.0 * M * N * K
    print(f"[{M}x{N}x{K}] cuBLAS linear {t_lib*1e3:7.1f} us {fl/t_lib/1e9:6.0f} TF/s | matmul {t_lib_nb*1e3:7.1f} us {fl/t_lib_nb/1e9:6.0f} TF/s | svit_gemm {t_own*1e3:7.1f} us {fl/t_own/1e9:6.0f} TF/s")
