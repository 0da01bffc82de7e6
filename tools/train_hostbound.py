"""Is the training step host-bound?  Host enqueue time per step (no synchronisation) vs GPU time per step."""
import sys, os, time
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
for _ in range(3): step()
torch.cuda.synchronize()
N = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(N): step()
t1 = time.perf_counter(); e1.record()
torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue {1e3*(t1-t0)/N:.2f} ms/step; gpu {e0.elapsed_time(e1)/N:.2f} ms/step; wall incl. drain {1e3*(t2-t0)/N:.2f} ms/step")
# forward only / backward only split of the host time
t0 = time.perf_counter()
for _ in range(N):
    with torch.no_grad(): model([clip])
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host enqueue of a no-grad forward {1e3*(t1-t0)/N:.2f} ms")
# true GPU time of one step: park the GPU behind a spin kernel while the host enqueues the whole step
torch.cuda.synchronize()
torch.cuda._sleep(int(60e-3 * 1.9e9))      # ~60 ms of GPU spinning
e0.record(); step(); e1.record()
torch.cuda.synchronize()
print(f"gpu time of one pre-enqueued step {e0.elapsed_time(e1):.2f} ms")
