"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (plain torch on the host, fp32 or fp64) of the SViT hot path of
eladb3/SViT: the MViTv2 pooled-attention block with object tokens.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file; nothing under svit_b200/ does.  It is the checker, never the thing shipped.

Parity status: PINNED.  Every function below is checked in tests/test_oracle_golden.py
against fixtures produced by executing the unmodified reference modules in the build
container (tests/golden/make_golden.py, which imports /root/reference through
oracle/ref_loader.py).  The one exception is RoIAlign (roi_align / roi_object_tokens):
the reference ships only the call site (video_model_builder.py:385-392, 472-491) and its
head_helper.py is absent, so that arithmetic restates torchvision.ops.roi_align
(aligned=True, sampling_ratio=0) and is pinned against torchvision 0.26 -- "parity
unpinned" with respect to the reference itself.

All model parameters are passed as a flat ``dict[str, Tensor]`` that uses the reference's
state_dict key names (e.g. ``blocks.3.attn.pool_q.weight``), so the same dict can be loaded
into the reference, into svit_b200, and here.

Each function cites the reference lines it restates (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

LN_EPS = 1e-6  # video_model_builder.py:68-69
HEAD_DIM = 96


# --------------------------------------------------------------------------------------
# A4: integer index tables of the decomposed relative position bias
# --------------------------------------------------------------------------------------
def rel_pos_index_table(q_n: int, k_n: int) -> torch.Tensor:
    """dist[i, j] (int64) exactly as attention.py:100-106 / 156-163 computes it: an fp32
    expression truncated by .long().  Must stay an fp32 torch expression to be bit-exact
    when the q/k ratio is not an integer (312^2 clips)."""
    q_ratio = max(k_n / q_n, 1.0)
    k_ratio = max(q_n / k_n, 1.0)
    dist = torch.arange(q_n)[:, None] * q_ratio - torch.arange(k_n)[None, :] * k_ratio
    dist += (k_n - 1) * k_ratio
    return dist.long()


def interp_rel_pos(rel_pos: torch.Tensor, d: int) -> torch.Tensor:
    """A2 -- attention.py:68-81: linear interpolation of the table to ``d`` rows."""
    ori = rel_pos.shape[0]
    if ori == d:
        return rel_pos
    new = F.interpolate(rel_pos.reshape(1, ori, -1).permute(0, 2, 1), size=d, mode="linear")
    return new.reshape(-1, d).permute(1, 0)


def rel_pos_tables(rel_pos: torch.Tensor, q_n: int, k_n: int) -> torch.Tensor:
    """R[a, b, :] = interp(rel_pos, 2*max(q,k)-1)[dist[a, b]]  -> [q_n, k_n, 96]."""
    tab = interp_rel_pos(rel_pos, int(2 * max(q_n, k_n) - 1))
    return tab[rel_pos_index_table(q_n, k_n)]


# --------------------------------------------------------------------------------------
# A1: attention_pool
# --------------------------------------------------------------------------------------
def pooled_thw(thw: Sequence[int], stride: Sequence[int], kernel=(3, 3, 3)) -> List[int]:
    """Conv/MaxPool output size, floor((n + 2p - k)/s) + 1 with p = k // 2."""
    return [(n + 2 * (k // 2) - k) // s + 1 for n, s, k in zip(thw, stride, kernel)]


def conv_obj_scale(w: torch.Tensor, stride: Sequence[int]) -> torch.Tensor:
    """Per-channel scale an object token receives from the pooling conv (attention.py:45-53):
    the token is broadcast to a kT x kH x kW cube, convolved with padding, and the outputs are
    averaged.  Equivalent to mean over output positions of the sum of in-bounds taps."""
    C = w.shape[0]
    k = w.shape[-3:]
    ones = torch.ones(1, C, *k, dtype=w.dtype, device=w.device)
    out = F.conv3d(ones, w, stride=tuple(stride), padding=tuple(x // 2 for x in k), groups=C)
    return out.mean(dim=(-1, -2, -3)).reshape(C)


def pool_tokens(z, conv_w, stride, gamma, beta, thw):
    """attention.py:13-65 for a Conv3d pool followed by LayerNorm.
    z [B, h, 1 + T*H*W + O, d] -> ([B, h, 1 + T'H'W' + O, d], [T',H',W'])."""
    B, h, N, d = z.shape
    T, H, W = thw
    L = T * H * W
    O = N - 1 - L
    assert O > 0  # attention.py:32
    cls, patch, obj = z[:, :, :1], z[:, :, 1:1 + L], z[:, :, 1 + L:]
    x = patch.reshape(B * h, T, H, W, d).permute(0, 4, 1, 2, 3)
    k = conv_w.shape[-3:]
    x = F.conv3d(x, conv_w, stride=tuple(stride), padding=tuple(i // 2 for i in k), groups=d)
    thw2 = list(x.shape[2:])
    x = x.reshape(B, h, d, -1).transpose(2, 3)
    obj = obj * conv_obj_scale(conv_w, stride)
    out = torch.cat([cls, x, obj], dim=2)
    out = F.layer_norm(out, (d,), gamma, beta, LN_EPS)
    return out, thw2


def skip_pool_tokens(x, stride, thw):
    """attention.py:562-564 with MaxPool3d(kernel=[s+1 if s>1 else s], stride, pad=k//2)
    (attention.py:503-505, 549-555): patch tokens max-pooled, cls and object tokens copied."""
    B, N, C = x.shape
    T, H, W = thw
    L = T * H * W
    kernel = [s + 1 if s > 1 else s for s in stride]
    cls, patch, obj = x[:, :1], x[:, 1:1 + L], x[:, 1 + L:]
    p = patch.reshape(B, T, H, W, C).permute(0, 4, 1, 2, 3)
    p = F.max_pool3d(p, kernel, tuple(stride), [k // 2 for k in kernel])
    thw2 = list(p.shape[2:])
    p = p.reshape(B, C, -1).transpose(1, 2)
    return torch.cat([cls, p, obj], dim=1), thw2


# --------------------------------------------------------------------------------------
# A3/A5/A7: pooled attention with decomposed relative position bias
# --------------------------------------------------------------------------------------
def attention_core(q, k, v, q_thw, k_thw, rel_h, rel_w, rel_t, scale=None):
    """attention.py:429-459.  q [B,h,Nq,d], k/v [B,h,Nk,d] are the pooled+normed tensors.
    S = (q*scale) k^T; patch x patch block += q.(Rh[i,i'] + Rw[j,j'] + Rt[t,t']) with the
    UN-scaled q; softmax over all keys; o = P v; o[1:] += q[1:]."""
    B, h, Nq, d = q.shape
    scale = d ** -0.5 if scale is None else scale  # attention.py:217-218
    qt, qh, qw = q_thw
    kt, kh, kw = k_thw
    Lq, Lk = qt * qh * qw, kt * kh * kw
    S = (q * scale) @ k.transpose(-2, -1)
    Rh = rel_pos_tables(rel_h, qh, kh)  # [qh, kh, d]
    Rw = rel_pos_tables(rel_w, qw, kw)
    Rt = rel_pos_tables(rel_t, qt, kt)
    qp = q[:, :, 1:1 + Lq].reshape(B, h, qt, qh, qw, d)
    eh = torch.einsum("bythwc,hkc->bythwk", qp, Rh)
    ew = torch.einsum("bythwc,wkc->bythwk", qp, Rw)
    et = torch.einsum("bythwc,tkc->bythwk", qp, Rt)
    bias = (et[..., :, None, None] + eh[..., None, :, None] + ew[..., None, None, :])
    S[:, :, 1:1 + Lq, 1:1 + Lk] += bias.reshape(B, h, Lq, Lk)
    P = S.softmax(dim=-1)
    o = P @ v
    o[:, :, 1:] += q[:, :, 1:]
    return o


def msa_forward(x, thw, p: Dict[str, torch.Tensor], prefix: str, num_heads: int,
                stride_q, stride_kv):
    """MultiScaleAttention.forward (attention.py:331-466) for the ssv2.yaml subset
    (mode=conv, pool_first=False, separate_qkv=False, cls on, rel-pos on, residual pooling)."""
    B, N, _ = x.shape
    g = lambda n: p[prefix + n]
    qkv = F.linear(x, g("qkv.weight"), g("qkv.bias")).reshape(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4)
    q, q_thw = pool_tokens(qkv[0], g("pool_q.weight"), stride_q, g("norm_q.weight"), g("norm_q.bias"), thw)
    k, k_thw = pool_tokens(qkv[1], g("pool_k.weight"), stride_kv, g("norm_k.weight"), g("norm_k.bias"), thw)
    v, _ = pool_tokens(qkv[2], g("pool_v.weight"), stride_kv, g("norm_v.weight"), g("norm_v.bias"), thw)
    o = attention_core(q, k, v, q_thw, k_thw, g("rel_pos_h"), g("rel_pos_w"), g("rel_pos_t"))
    o = o.transpose(1, 2).reshape(B, -1, num_heads * o.shape[-1])
    return F.linear(o, g("proj.weight"), g("proj.bias")), q_thw


def mlp_forward(x, p, prefix):
    """common.py:27-34: fc1 -> exact-erf GELU -> fc2."""
    hdn = F.gelu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))
    return F.linear(hdn, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])


def block_forward(x, thw, p, prefix, spec, drop_masks=None):
    """MultiScaleBlock.forward (attention.py:557-571), dim_mul_in_att=True.
    ``drop_masks`` = (m_attn[B], m_mlp[B]) already divided by keep-prob, or None (eval)."""
    C = x.shape[-1]
    xn = F.layer_norm(x, (C,), p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], LN_EPS)
    xb, thw2 = msa_forward(xn, thw, p, prefix + "attn.", spec["num_heads"], spec["stride_q"], spec["stride_kv"])
    if spec["dim"] != spec["dim_out"]:
        x = F.linear(xn, p[prefix + "proj.weight"], p[prefix + "proj.bias"])
    xres, _ = skip_pool_tokens(x, spec["stride_q"], thw)
    if drop_masks is not None:
        xb = xb * drop_masks[0][:, None, None]
    x = xres + xb
    D = x.shape[-1]
    xn2 = F.layer_norm(x, (D,), p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], LN_EPS)
    xm = mlp_forward(xn2, p, prefix + "mlp.")
    if drop_masks is not None:
        xm = xm * drop_masks[1][:, None, None]
    return x + xm, thw2


# --------------------------------------------------------------------------------------
# A10 / R1 / R2: stem, object tokens, split + head
# --------------------------------------------------------------------------------------
def patch_embed(x, w, b, stride=(2, 4, 4), padding=(1, 3, 3)):
    """stem_helper.py:309-320: conv3d then flatten(2).transpose(1,2)."""
    y = F.conv3d(x, w, b, stride=tuple(stride), padding=tuple(padding))
    return y.flatten(2).transpose(1, 2), list(y.shape[2:])


def object_tokens(object_queries, pos_embed_temporal, B, Tx):
    """video_model_builder.py:354-363.  [B, Tx*O, C]; token (t, o) = query[o] + pos_t[t]
    (frame mode Tx == 1: no temporal term)."""
    O, C = object_queries.shape[1:]
    xo = object_queries.unsqueeze(1).expand(B, Tx, O, C)
    if Tx > 1:
        xo = xo + pos_embed_temporal.unsqueeze(2).expand(B, Tx, O, C)
    return xo.flatten(1, 2)


def object_token_index(T, H, W, t, o, O=4):
    """INT: sequence index of object token (frame t, slot o): 1 + T'H'W' + t*O + o."""
    return 1 + T * H * W + t * O + o


def head_forward(x_cls_obj, p, Tx, training=False, prefix="head."):
    """SViTHead.forward (video_model_builder.py:507-546), dropout off."""
    B = x_cls_obj.shape[0]
    x, xobj = x_cls_obj[:, 0], x_cls_obj[:, 1:]
    xobj = xobj.reshape(B, Tx, -1, xobj.shape[-1])
    logits = F.linear(x, p[prefix + "projection.weight"], p[prefix + "projection.bias"])
    out = logits if training else logits.softmax(dim=1)
    boxes = F.linear(xobj, p[prefix + "boxes_mlp.0.weight"], p[prefix + "boxes_mlp.0.bias"]).sigmoid()
    bce = F.linear(xobj, p[prefix + "boxes_bce_mlp.weight"], p[prefix + "boxes_bce_mlp.bias"])
    contact = F.linear(xobj[:, :, :2], p[prefix + "contact_mlp.weight"], p[prefix + "contact_mlp.bias"])
    if not training:
        bce = bce.sigmoid()
        contact = contact.softmax(dim=-1)
    extra = {"obj_desc": xobj, "pred_bboxes": torch.cat([bce, boxes], dim=-1),
             "pred_contact_state": contact, "logits": logits}
    return out, extra


def svit_forward(clip, p: Dict[str, torch.Tensor], specs, cfg, training=False, return_tokens=False):
    """SViT.forward (video_model_builder.py:315-398), abs-pos off, detection off.
    ``specs`` = svit_b200.config.block_specs(cfg)[0] (pure geometry)."""
    x = clip
    if x.dim() == 4:
        x = x.unsqueeze(2)
    Tx = x.shape[2]
    ps = cfg.MVIT.PATCH_STRIDE
    x, (Tp, H, W) = patch_embed(x, p["patch_embed.proj.weight"], p["patch_embed.proj.bias"],
                                ps, cfg.MVIT.PATCH_PADDING)
    T = cfg.DATA.NUM_FRAMES // ps[0] if Tx > 1 else Tx
    B = x.shape[0]
    xo = object_tokens(p["object_queries"], p["pos_embed_temporal"], B, Tx)
    O_tot = xo.shape[1]
    x = torch.cat([p["cls_token"].expand(B, -1, -1), x, xo], dim=1)
    thw = [T, H, W]
    for i, spec in enumerate(specs):
        x, thw = block_forward(x, thw, p, f"blocks.{i}.", spec)
    C = x.shape[-1]
    x = F.layer_norm(x, (C,), p["norm.weight"], p["norm.bias"], LN_EPS)
    cls, obj, patch = x[:, :1], x[:, -O_tot:], x[:, 1:-O_tot]
    out, extra = head_forward(torch.cat([cls, obj], dim=1), p, Tx, training)
    if return_tokens:
        extra["patch_tokens"] = patch
        extra["thw"] = thw
    return out, extra


# --------------------------------------------------------------------------------------
# R3: RoIAlign object tokens (reference arithmetic absent -> torchvision semantics)
# --------------------------------------------------------------------------------------
def roi_align(feat: torch.Tensor, rois: torch.Tensor, out_size: int, spatial_scale: float,
              sampling_ratio: int = 0, aligned: bool = True) -> torch.Tensor:
    """Restates torchvision.ops.roi_align (the op detectron2.layers.ROIAlign wraps).
    feat [N,C,H,W]; rois [K,5] = (batch_idx, x1,y1,x2,y2) in input pixels -> [K,C,P,P].
    Pure-python loops over boxes/bins: small cases only."""
    N, C, H, W = feat.shape
    K = rois.shape[0]
    P = out_size
    out = torch.zeros(K, C, P, P, dtype=feat.dtype)
    off = 0.5 if aligned else 0.0
    for r in range(K):
        b = int(rois[r, 0])
        x1, y1, x2, y2 = [float(v) * spatial_scale - off for v in rois[r, 1:]]
        rw, rh = x2 - x1, y2 - y1
        if not aligned:
            rw, rh = max(rw, 1.0), max(rh, 1.0)
        bw, bh = rw / P, rh / P
        gh = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rh / P))
        gw = sampling_ratio if sampling_ratio > 0 else int(math.ceil(rw / P))
        cnt = max(gh * gw, 1)
        for ph in range(P):
            for pw in range(P):
                acc = torch.zeros(C, dtype=feat.dtype)
                for iy in range(gh):
                    y = y1 + ph * bh + (iy + 0.5) * bh / gh
                    for ix in range(gw):
                        x = x1 + pw * bw + (ix + 0.5) * bw / gw
                        acc += _bilinear(feat[b], y, x, H, W)
                out[r, :, ph, pw] = acc / cnt
    return out


def _bilinear(fm, y, x, H, W):
    if y < -1.0 or y > H or x < -1.0 or x > W:
        return torch.zeros(fm.shape[0], dtype=fm.dtype)
    y = max(y, 0.0)
    x = max(x, 0.0)
    yl, xl = int(y), int(x)
    if yl >= H - 1:
        yh = yl = H - 1
        y = float(yl)
    else:
        yh = yl + 1
    if xl >= W - 1:
        xh = xl = W - 1
        x = float(xl)
    else:
        xh = xl + 1
    ly, lx = y - yl, x - xl
    hy, hx = 1.0 - ly, 1.0 - lx
    return hy * hx * fm[:, yl, xl] + hy * lx * fm[:, yl, xh] + ly * hx * fm[:, yh, xl] + ly * lx * fm[:, yh, xh]


def frame_to_slice(t: int, Tx: int, patch_stride_t: int) -> int:
    """INT: input frame t -> temporal slice of the patch grid (t // patch_stride_t; identity in frame mode)."""
    return t if Tx == 1 else t // patch_stride_t


def roi_object_tokens(feat, boxes, patch_stride_t=2, spatial_scale=1.0 / 16, out_size=7):
    """Per-frame box-conditioned object tokens (north_star extension of
    video_model_builder.py:385-392, 472-491).  feat [B,C,T',H',W']; boxes [B,Tx,K,4] xyxy pixels.
    One RoIAlign(out_size, aligned) per (b,t,k) on slice t//patch_stride_t, then max over the
    P x P bins -> tokens [B, Tx*K, C] in (t, k) order; also returns the int (b, slice) table."""
    B, C, Tp, H, W = feat.shape
    _, Tx, K, _ = boxes.shape
    toks = torch.zeros(B, Tx * K, C, dtype=feat.dtype)
    assign = torch.zeros(B, Tx * K, 2, dtype=torch.int64)
    for b in range(B):
        for t in range(Tx):
            s = frame_to_slice(t, Tx, patch_stride_t) if Tp > 1 else 0
            rois = torch.cat([torch.zeros(K, 1, dtype=boxes.dtype), boxes[b, t]], dim=1)
            r = roi_align(feat[b:b + 1, :, s], rois, out_size, spatial_scale, 0, True)
            toks[b, t * K:(t + 1) * K] = r.amax(dim=(-1, -2))
            assign[b, t * K:(t + 1) * K, 0] = b
            assign[b, t * K:(t + 1) * K, 1] = s
    return toks, assign


# --------------------------------------------------------------------------------------
# R4: box -> slot assignment (integer / ordering semantics, bit-exact)
# --------------------------------------------------------------------------------------
def assign_slots(labels: Sequence[Tuple[str, Sequence[float]]], num_boxes: int = 4) -> torch.Tensor:
    """ssv2_frames.py:503-517: in annotation order, 'hand' -> slots 0,1, anything else ->
    slots 2,3; at most two per category, extras dropped.  Returns [1, num_boxes, 4] xyxy."""
    out = torch.zeros((1, num_boxes, 4), dtype=torch.float32)
    inds = {"hand": 0, "obj": 0}
    offs = {"hand": 0, "obj": 2}
    for cat, box in labels:
        c = "hand" if cat == "hand" else "obj"
        if inds[c] > 1:
            continue
        out[0, inds[c] + offs[c]] = torch.tensor(list(box), dtype=torch.float32)
        inds[c] += 1
    return out


def match_haog(haog: torch.Tensor):
    """box_ops.py:140-194.  Cost = L2 distance between the first two coordinates (top-left
    corners for xyxy input: line 165 discards the cxcywh conversion of line 159); all-zero
    boxes cost 1e8; if the crossed pairing is cheaper the order becomes (0, 2, 3, 1) --
    the reference's own variable mix-up, reproduced verbatim.  Contact state per pair:
    -1 if 1e8, 3 if dist < 0.1, else 0."""
    HIGH = 1e8
    squeeze = haog.ndim == 3
    if squeeze:
        assert haog.size(0) == 1
        haog = haog[0]
    xy = haog[:, :2]
    cost = torch.cdist(xy[None, :2], xy[None, 2:], p=2)[0]
    obj_zero = torch.all(haog[2:] == 0, dim=-1)
    hand_zero = torch.all(haog[:2] == 0, dim=-1)
    cost[:, obj_zero] = HIGH
    cost[:, hand_zero] = HIGH  # indexes columns, as the reference does
    if cost[0, 1] + cost[1, 0] < cost[0, 0] + cost[1, 1]:
        haog = torch.stack((haog[0], haog[2], haog[3], haog[1]), dim=0)
        d = [cost[0, 1], cost[1, 0]]
    else:
        d = [cost[0, 0], cost[1, 1]]
    state = [(-1 if x == HIGH else (3 if x < 0.1 else 0)) for x in d]
    if squeeze:
        haog = haog[None]
    return haog, torch.tensor(state, dtype=torch.int64)


def zero_empty_boxes(boxes: torch.Tensor, mode="cxcywh", eps=0.05) -> torch.Tensor:
    """box_ops.py:116-130: zero every box whose w or h <= eps."""
    shp = boxes.shape
    b = boxes.reshape(-1, 4).clone()
    wh = b[:, 2:] if mode == "cxcywh" else b[:, 2:] - b[:, :2]
    b[torch.any(wh <= eps, dim=-1)] = 0
    return b.reshape(shp)


def xyxy_to_cxcywh(b):
    """box_ops.py:32-36."""
    x0, y0, x1, y1 = b.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], dim=-1)


def normalise_boxes(boxes_xyxy: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """ssv2_frames.py:347-353: /crop, clip to [0,1], xyxy->cxcywh, zero_empty_boxes."""
    b = boxes_xyxy.clone()
    b[..., [0, 2]] = b[..., [0, 2]] / w
    b[..., [1, 3]] = b[..., [1, 3]] / h
    return zero_empty_boxes(xyxy_to_cxcywh(b.clamp(0, 1)))


def gen_random_boxes(T: int, O: int, rng) -> torch.Tensor:
    """Synthetic box recipe of doh_frames.py:479-491 (cxcywh in [0,1]); ``rng`` is a
    numpy RandomState so the draw order matches np.random.rand(T, O, 4)."""
    import numpy as np

    out = rng.rand(T, O, 4)
    cxcy, wh = out[:, :, :2], out[:, :, 2:]
    dmax = np.min(np.stack([cxcy, 1 - cxcy], axis=0), axis=0) * 2
    return torch.from_numpy(np.concatenate([cxcy, wh * dmax], axis=2))
