"""Patch-embed timing: space-to-depth layout kernel + implicit GEMM (SVIT_PE_MODE selects the operand scheme) vs im2col + GEMM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = torch.Generator().manual_seed(0)
w = (torch.randn(96, 3, 3, 7, 7, generator=g) * 0.1).cuda(); b = torch.zeros(96).cuda()
cls, qs, pt = torch.zeros(1, 1, 96).cuda(), torch.zeros(1, 4, 96).cuda(), torch.zeros(1, 16, 96).cuda()
clip = torch.randn(B, 3, 16, 224, 224, generator=g).bfloat16().cuda()
frames = torch.randint(0, 256, (B, 16, 224, 224, 3), generator=g, dtype=torch.uint8).cuda()
K, S, P = (3, 7, 7), (2, 4, 4), (1, 3, 3)
def bench(f, it=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it
with torch.no_grad():
    ops.profile_start()
    for _ in range(3): ops.patch_embed_tokens(clip, w, b, cls, qs, pt, K, S, P, torch.bfloat16)
    for _ in range(3): ops.patch_embed_tokens(frames, w, b, cls, qs, pt, K, S, P, torch.bfloat16, mean=[0.45] * 3, std=[0.225] * 3)
    pr = ops.profile_stop(3)["detail"]
    for k, v in pr.items(): print(f"  {v['ms_per_step']/max(v['calls_per_step'],1e-9)*1e3:8.1f} us  x{v['calls_per_step']:.0f} {k}")
    t_new = bench(lambda: ops.patch_embed_tokens(clip, w, b, cls, qs, pt, K, S, P, torch.bfloat16))
    t_u8 = bench(lambda: ops.patch_embed_tokens(frames, w, b, cls, qs, pt, K, S, P, torch.bfloat16, mean=[0.45] * 3, std=[0.225] * 3))
    ops._state["implicit_patch_embed"] = False
    t_old = bench(lambda: ops.patch_embed_tokens(clip, w, b, cls, qs, pt, K, S, P, torch.bfloat16))
print(f"mode {os.environ.get('SVIT_PE_MODE', '0')}: implicit (bf16 clip) {t_new*1e3:.0f} us, implicit (uint8 frames) {t_u8*1e3:.0f} us, im2col + GEMM {t_old*1e3:.0f} us")
