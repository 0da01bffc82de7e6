"""Training-step timing A/B: eager without optimizer, eager with the fused AdamW, CUDA-graph replay (1 GPU, B = 8)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, svit_b200
from svit_b200 import ops
from svit_b200.config import ssv2_cfg
from svit_b200.optim import construct_optimizer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = ssv2_cfg(); torch.manual_seed(0)
model = svit_b200.SViT(cfg, compute_dtype=torch.bfloat16).cuda().train()
clip = torch.randn(B, 3, 16, 224, 224).bfloat16().cuda()
labels = torch.randint(0, 174, (B,)).cuda()
opt = construct_optimizer(model, cfg)
def step(use_opt):
    for p in model.parameters(): p.grad = None
    preds, extra = model([clip])
    loss = torch.nn.functional.cross_entropy(extra["logits"].float(), labels)
    loss.backward()
    if use_opt: opt.step(max_norm=1.0)
    return loss
def timeit(f, n=8, w=3):
    for _ in range(w): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): f()
    e1.record(); t1 = time.perf_counter(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n
print("eager, no optimizer   : gpu %.2f ms  host-enqueue %.2f ms" % timeit(lambda: step(False)))
print("eager, fused AdamW    : gpu %.2f ms  host-enqueue %.2f ms" % timeit(lambda: step(True)))
ops.profile_start(); step(True); prof = ops.profile_stop(1)
rows = sorted(prof["detail"].items(), key=lambda kv: -kv[1]["ms_per_step"])
print("profiled step with optimizer: total kernel ms", sum(v["ms_per_step"] for _, v in rows))
for k, v in rows[:8]: print(f"{v['ms_per_step']:9.3f} x{v['calls_per_step']:<4.0f} {k}")
g = svit_b200.GraphedTrainStep(model, opt, clip, labels, max_norm=1.0)
print("graph replay          : gpu %.2f ms  host-enqueue %.2f ms" % timeit(lambda: g(clip, labels)))
