"""configs[4]: pooled-attention micro-bench per MViTv2 stage shape (stride/kernel from ssv2.yaml) at 32x312^2 clips.
One MultiScaleAttention per distinct stage shape, batch B (default 8), bf16; prints the kernel time of the pooled
attention core and of the three attention_pool calls with their roofline fractions (MEASURED_PEAKS.json)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from functools import partial
import svit_b200
from svit_b200 import ops
from svit_b200.config import ssv2_cfg, block_specs

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
frames, crop = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (32, 312)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}
cfg = ssv2_cfg(); cfg.DATA.NUM_FRAMES = frames; cfg.DATA.TRAIN_CROP_SIZE = crop
specs = block_specs(cfg)[0]
ps = cfg.MVIT.PATCH_STRIDE
thw = [frames // ps[0], crop // ps[1], crop // ps[2]]
O = frames * cfg.SVIT.O
LN = partial(torch.nn.LayerNorm, eps=1e-6)
seen = set()
rows = []
torch.manual_seed(0)
for i, sp in enumerate(specs):
    sq = sp["stride_q"][1] if sp["stride_q"] else 1
    key = (sp["dim"], sp["dim_out"], tuple(thw), sq, tuple(sp["stride_kv"]))
    nthw = [thw[0], (thw[1] - 1) // sq + 1, (thw[2] - 1) // sq + 1]
    if key not in seen:
        seen.add(key)
        m = svit_b200.MultiScaleAttention(sp["dim"], sp["dim_out"], sp["input_size"], num_heads=sp["num_heads"], qkv_bias=True,
                                          kernel_q=sp["kernel_q"], kernel_kv=sp["kernel_kv"], stride_q=sp["stride_q"],
                                          stride_kv=sp["stride_kv"], norm_layer=LN, rel_pos_spatial=True,
                                          rel_pos_temporal=True, residual_pooling=True).cuda()
        N = 1 + thw[0] * thw[1] * thw[2] + O
        x = torch.randn(B, N, sp["dim"], device="cuda").bfloat16()
        with torch.no_grad():
            for _ in range(2): m(x, thw)
            torch.cuda.synchronize()
            ops.profile_start()
            for _ in range(3): m(x, thw)
            prof = ops.profile_stop(3)
        att = [(k, v) for k, v in prof["detail"].items() if k.startswith("svit_attn_fwd")][0]
        pool_ms = sum(v["ms_per_step"] for k, v in prof["detail"].items() if k.startswith("svit_pool_ln_fwd"))
        h = sp["num_heads"]
        skv = sp["stride_kv"][1]
        kthw = [thw[0], (thw[1] - 1) // skv + 1, (thw[2] - 1) // skv + 1]
        Nq = 1 + nthw[0] * nthw[1] * nthw[2] + O
        Nk = 1 + kthw[0] * kthw[1] * kthw[2] + O
        fl = 4.0 * B * h * Nq * Nk * 96 + 2.0 * B * h * (Nq - 1 - O) * 96 * sum(kthw)
        pool_bytes = 2 * 96 * B * h * (3 * N + Nq + 2 * Nk)
        tf = fl / (att[1]["ms_per_step"] / 1e3) / 1e12
        gb = pool_bytes / (pool_ms / 1e3) / 1e9
        rows.append(dict(block=i, h=h, Nq=Nq, Nk=Nk, attn_us=att[1]["ms_per_step"] * 1e3, attn_tflops=tf,
                         attn_frac_of_sustained_peak=tf / peaks["bf16_tflops_sustained"], pool_us=pool_ms * 1e3,
                         pool_gbs=gb, pool_frac_of_hbm=gb / peaks["hbm_gbs"]))
        print(f"blk{i:2d} h{h} Nq{Nq:6d} Nk{Nk:5d}: attention {att[1]['ms_per_step']*1e3:8.1f} us {tf:6.0f} TF/s "
              f"({100*tf/peaks['bf16_tflops_sustained']:4.1f}% of sustained bf16 peak) | attention_pool x3 {pool_ms*1e3:8.1f} us "
              f"{gb:6.0f} GB/s ({100*gb/peaks['hbm_gbs']:4.1f}% of HBM)")
    thw = nthw
print(json.dumps({"config": f"pooled-attention micro-bench, {frames}x{crop}^2, B={B}", "rows": rows}))
