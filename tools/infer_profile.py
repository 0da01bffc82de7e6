"""Per-kernel breakdown of one inference forward (batch 64 by default) via the ops profile hooks."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200 import ops
from svit_b200.config import ssv2_cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cfg = ssv2_cfg(); torch.manual_seed(0)
model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).cuda().eval()
clip = torch.randn(B, 3, 16, 224, 224).bfloat16().cuda()
with torch.no_grad():
    for _ in range(2): model([clip])
    torch.cuda.synchronize()
    ops.profile_start(); model([clip]); prof = ops.profile_stop(1)
rows = sorted(prof["detail"].items(), key=lambda kv: -kv[1]["ms_per_step"])
print("total kernel ms", sum(v["ms_per_step"] for _, v in rows))
for k, v in sorted(prof["families"].items(), key=lambda kv: -kv[1]["ms_per_step"]):
    print(f"{v['ms_per_step']:9.3f} x{v['calls_per_step']:<4.0f} {k}")
print()
for k, v in rows[:45]: print(f"{v['ms_per_step']:9.3f} x{v['calls_per_step']:<4.0f} {k}")
