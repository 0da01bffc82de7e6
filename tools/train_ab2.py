import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200 import ops
from svit_b200.config import ssv2_cfg
from svit_b200.optim import construct_optimizer
B = 8
cfg = ssv2_cfg(); torch.manual_seed(0)
model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).cuda().train()
clip = torch.randn(B, 3, 16, 224, 224).bfloat16().cuda()
labels = torch.randint(0, 174, (B,)).cuda()
opt = construct_optimizer(model, cfg)
T = {}
def tic(k, t0): T[k] = T.get(k, 0.0) + (time.perf_counter() - t0) * 1e3
def step():
    t = time.perf_counter()
    for p in model.parameters(): p.grad = None
    tic("zero", t); t = time.perf_counter()
    preds, extra = model([clip])
    loss = torch.nn.functional.cross_entropy(extra["logits"].float(), labels)
    tic("fwd", t); t = time.perf_counter()
    loss.backward()
    tic("bwd", t); t = time.perf_counter()
    opt.step(max_norm=1.0)
    tic("opt", t)
for _ in range(3): step()
torch.cuda.synchronize(); T.clear()
for _ in range(5): step()
torch.cuda.synchronize()
print({k: round(v / 5, 2) for k, v in T.items()})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
