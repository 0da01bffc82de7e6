"""Library attention (torch SDPA: cuDNN / flash backends, bf16, head_dim 96, NO rel-pos bias, no residual) on the model's
pooled-attention shapes next to svit_attn_fwd (which includes the decomposed rel-pos bias and the residual pooling add)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.nn.attention import sdpa_kernel, SDPBackend
from svit_b200 import ops, msa
shapes = [(64, 4, (8, 14, 14), (8, 7, 7)), (64, 1, (8, 56, 56), (8, 7, 7)), (64, 2, (8, 28, 28), (8, 14, 14)), (64, 4, (8, 14, 14), (8, 14, 14))]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, n=6):
    for _ in range(2): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
O = 64
gen = torch.Generator().manual_seed(0)
for B, h, q_thw, k_thw in shapes:
    Nq, Nk = 1 + q_thw[0] * q_thw[1] * q_thw[2] + O, 1 + k_thw[0] * k_thw[1] * k_thw[2] + O
    q = torch.randn(B, h, Nq, 96, generator=gen).bfloat16().cuda()
    k = torch.randn(B, h, Nk, 96, generator=gen).bfloat16().cuda()
    v = torch.randn(B, h, Nk, 96, generator=gen).bfloat16().cuda()
    fl = 4.0 * B * h * Nq * Nk * 96
    line = f"B{B} h{h} Nq{Nq} Nk{Nk}:"
    with torch.no_grad():
        for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION)):
            try:
                with sdpa_kernel(be):
                    t = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v))
                line += f" {name} {t*1e3:7.1f} us {fl/t/1e9:5.0f} TF/s |"
            except Exception as e:
                line += f" {name} unavailable ({type(e).__name__}) |"
        pairs = ((q_thw[1], k_thw[1]), (q_thw[2], k_thw[2]), (q_thw[0], k_thw[0]))
        rels = [(0.2 * torch.randn(2 * max(a, b) - 1, 96, generator=gen)).cuda() for a, b in pairs]
        R = [msa.gathered_rel_pos(r, a, b) for r, (a, b) in zip(rels, pairs)]
        tabs = [r.bfloat16() for r in rels]
        tc_tables = (torch.cat(tabs).contiguous(), [t.shape[0] for t in tabs], msa._index32_on(q.device, *pairs[0]),
                     msa._index32_on(q.device, *pairs[1]), msa._index32_on(q.device, *pairs[2]),
                     msa.key_column_codes(k_thw, O, q.device), msa.key_select_table(k_thw, O, q.device))
        t = timeit(lambda: ops.attention(q, k, v, R[0], R[1], R[2], q_thw, k_thw, O, 96 ** -0.5, tc_tables))
        line += f" svit_attn_fwd (with bias + residual) {t*1e3:7.1f} us {fl/t/1e9:5.0f} TF/s"
    print(line)
