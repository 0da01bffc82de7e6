"""One training step (fwd + bwd, batch 8) for an ncu launch list: warm-up outside the profiled range."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
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
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
