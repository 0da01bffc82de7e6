"""Per-role timeline of CTA (0,0) of the tcgen05 attention kernel (diagnostic; svit_debug_attn_timeline hook).
usage: python tools/attn_timeline.py B h qT qH qW kT kH kW O"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from svit_b200 import ops, _lib, msa
B, h, qt, qh, qw, kt, kh, kw, O = (int(x) for x in sys.argv[1:10])
q_thw, k_thw = (qt, qh, qw), (kt, kh, kw)
gen = torch.Generator().manual_seed(0)
Nq, Nk = 1 + qt * qh * qw + O, 1 + kt * kh * kw + O
q = torch.randn(B, h, Nq, 96, generator=gen).bfloat16().cuda()
k = torch.randn(B, h, Nk, 96, generator=gen).bfloat16().cuda()
v = torch.randn(B, h, Nk, 96, generator=gen).bfloat16().cuda()
pairs = ((qh, kh), (qw, kw), (qt, kt))
rels = [(0.2 * torch.randn(2 * max(a, b) - 1, 96, generator=gen)).cuda() for a, b in pairs]
R = [msa.gathered_rel_pos(r, a, b) for r, (a, b) in zip(rels, pairs)]
tabs = [r.bfloat16() for r in rels]
tc_tables = (torch.cat(tabs).contiguous(), [t.shape[0] for t in tabs], msa._index32_on(q.device, qh, kh),
             msa._index32_on(q.device, qw, kw), msa._index32_on(q.device, qt, kt), msa.key_column_codes(k_thw, O, q.device),
             msa.key_select_table(k_thw, O, q.device))
run = lambda: ops.attention(q, k, v, R[0], R[1], R[2], q_thw, k_thw, O, 96 ** -0.5, tc_tables)
with torch.no_grad():
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 4.0 * B * h * Nq * Nk * 96
    print(f"B{B} h{h} Nq{Nq} Nk{Nk}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TF/s")
    buf = torch.zeros(3 * 4096 * 2, dtype=torch.int64, device="cuda")
    hook = _lib.lib().svit_debug_attn_timeline
    hook.argtypes = [ctypes.c_void_p]
    hook(buf.data_ptr()); run(); torch.cuda.synchronize(); hook(None)
b = buf.cpu().reshape(3, 4096, 2)
t0 = min(int(b[r, 0, 1]) for r in range(3) if int(b[r, 0, 1]) > 0)
names = {0: "prod", 1: "mma", 2: "smx"}
ev = []
for r in range(3):
    for i in range(4096):
        tag, t = int(b[r, i, 0]), int(b[r, i, 1])
        if t == 0: break
        ev.append((t - t0, names[r], tag))
ev.sort()
for t, nm, tag in ev[:220]: print(f"{t:9d} {nm:5s} {tag}")
