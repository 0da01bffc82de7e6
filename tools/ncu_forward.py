"""Three eager batch-64 inference forwards of the ssv2.yaml model (the ncu target: `ncu -k regex:... python tools/ncu_forward.py`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200.config import ssv2_cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(0)
m = svit_b200.SViT(ssv2_cfg(), compute_dtype=torch.bfloat16).cuda().eval()
x = torch.randn(B, 3, 16, 224, 224).bfloat16().cuda()
with torch.no_grad():
    for _ in range(n):
        m([x])
torch.cuda.synchronize()
print("ok")
