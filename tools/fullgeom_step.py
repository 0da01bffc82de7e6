"""One full-geometry (configs/ssv2.yaml, 16x224^2) bf16 forward + backward at batch B (default 2): the sanitizer target."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200.config import ssv2_cfg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = ssv2_cfg(); torch.manual_seed(0)
model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).cuda().train()
clip = torch.randn(B, 3, 16, 224, 224).bfloat16().cuda()
labels = torch.randint(0, 174, (B,)).cuda()
preds, extra = model([clip])
loss = torch.nn.functional.cross_entropy(extra["logits"].float(), labels)
loss.backward()
torch.cuda.synchronize()
print("full-geometry fwd+bwd ok, loss", float(loss), "grad norm", float(torch.sqrt(sum(p.grad.float().square().sum() for p in model.parameters() if p.grad is not None))))
