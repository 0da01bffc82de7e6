"""Pure-read / pure-write / copy HBM bandwidth on this GPU (diagnostic for the roofline denominators)."""
import torch
n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
def t(f, iters=10):
    for _ in range(3): f()
    best = 1e9
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
w = t(lambda: a.fill_(1.0))
c = t(lambda: b.copy_(a))
r = t(lambda: a.sum())
print(f"write {2*n/w/1e6:.0f} GB/s  copy {4*n/c/1e6:.0f} GB/s  read(sum) {2*n/r/1e6:.0f} GB/s")
