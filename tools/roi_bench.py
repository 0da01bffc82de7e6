"""BASELINE.json configs[3]: object-token stress sweep -- per-frame RoIAlign (7x7, aligned, adaptive sampling) + max
over bins, T = 16 frames x K boxes/frame, on the token-major feature grid of every MViTv2 stage.
Prints one JSON line per case: time, boxes/s, achieved GB/s against the lower-bound traffic of SURVEY 8(d):
elem * C * (min(feature map, sum of box footprints) + boxes)  (the kernel writes the max over the 49 bins, one row per box).
usage: python tools/roi_bench.py [out.json]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from svit_b200 import ops

def bench(f, it=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it

def boxes_xyxy(B, T, K, rng, size=224.0):  # datasets/doh_frames.py:479-491 recipe, then cxcywh -> xyxy pixels
    c = rng.uniform(0, 1, (B, T, K, 2))
    wh = rng.uniform(0, 1, (B, T, K, 2)) * 2 * np.minimum(c, 1 - c)
    return torch.tensor(np.concatenate([c - wh / 2, c + wh / 2], -1) * size, dtype=torch.float32)

peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
rows = []
rng = np.random.RandomState(1234)
for (C, Tf, Hf) in ((96, 8, 56), (192, 8, 28), (384, 8, 14), (768, 8, 7)):
    for B in (1, 8, 64):
        for K in (1, 2, 4, 8, 16):
            N = 1 + Tf * Hf * Hf + 64
            x = torch.randn(B, N, C, device="cuda").bfloat16()
            bx = boxes_xyxy(B, 16, K, rng).cuda()
            ms = bench(lambda: ops.roi_tokens(x, (Tf, Hf, Hf), bx, 2, Hf / 224.0, 7))
            nb = B * 16 * K
            s = bx.cpu().numpy() * (Hf / 224.0) - 0.5
            fw = np.clip(np.ceil(s[..., 2]) + 1, 0, Hf) - np.clip(np.floor(s[..., 0]), 0, Hf)
            fh = np.clip(np.ceil(s[..., 3]) + 1, 0, Hf) - np.clip(np.floor(s[..., 1]), 0, Hf)
            foot = float(np.sum(np.maximum(fw, 1) * np.maximum(fh, 1)))
            touched = min(B * Tf * Hf * Hf, foot)
            by = 2 * C * (touched + nb)
            rows.append({"C": C, "grid": [Tf, Hf, Hf], "B": B, "K": K, "boxes": nb, "us": round(ms * 1e3, 2),
                         "boxes_per_s": round(nb / ms * 1e3), "lower_bound_bytes": int(by), "GB_s": round(by / ms / 1e6, 1)})
            print(json.dumps(rows[-1]))
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=0)
