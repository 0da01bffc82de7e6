"""cProfile of the host side of one training step (where do the ~50 us per C-ABI launch go?)."""
import sys, os, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200.config import ssv2_cfg
cfg = ssv2_cfg(); torch.manual_seed(0)
model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).cuda().train()
clip = torch.randn(8, 3, 16, 224, 224).bfloat16().cuda()
labels = torch.randint(0, 174, (8,)).cuda()
def step():
    for p in model.parameters(): p.grad = None
    preds, extra = model([clip])
    torch.nn.functional.cross_entropy(extra["logits"].float(), labels).backward()
for _ in range(3): step()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
