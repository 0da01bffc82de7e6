"""Per-kernel breakdown of one training step (fwd + bwd), batch 8, via the ops profile hooks."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200 import ops
from svit_b200.config import ssv2_cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = ssv2_cfg(); torch.manual_seed(0)
model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).cuda().train()
clip = torch.randn(B, 3, 16, 224, 224).bfloat16().cuda()
labels = torch.randint(0, 174, (B,)).cuda()
def step():
    for p in model.parameters(): p.grad = None
    preds, extra = model([clip])
    torch.nn.functional.cross_entropy(extra["logits"].float(), labels).backward()
for _ in range(2): step()
torch.cuda.synchronize()
ops.profile_start(); step(); prof = ops.profile_stop(1)
rows = sorted(prof["detail"].items(), key=lambda kv: -kv[1]["ms_per_step"])
tot = sum(v["ms_per_step"] for _, v in rows)
print("total kernel ms", tot)
fam = sorted(prof["families"].items(), key=lambda kv: -kv[1]["ms_per_step"])
for k, v in fam: print(f"{v['ms_per_step']:9.3f} x{v['calls_per_step']:<4.0f} {k}")
print()
for k, v in rows[:25]: print(f"{v['ms_per_step']:9.3f} x{v['calls_per_step']:<4.0f} {k}")
